"""Frame-window sharding over 2 and 3 ranks on CPU (gloo): every rank scans only the slots living on its frames,
the fp64 accumulator block is summed with one all-reduce, and the (replicated) small-parameter update is
then bit-identical on all ranks and equal to the single-shard result.  The per-rank work is re-enacted by
the CPU emulation harness; the GPU path runs the same plan/reduce code with NCCL (test_gpu_parity.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.helpers import (TRAIN_ALL, cpu_problem, emul_eval, emul_lib, go, golden_case, _lib, adam_schedule)
from align3r_b200.plan import partition_frames, frame_costs
import ctypes as C
import numpy as np


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fx, scene, prob = golden_case(name)
        edges = np.asarray(prob.edges, dtype=np.int32)
        wins = partition_frames(frame_costs(edges, np.asarray([(prob.H, prob.W)] * prob.n_imgs)), world)
        bufs = cpu_problem(prob, fx["init_params"], frame_range=wins[rank])
        lib = emul_lib()
        st = bufs.struct(TRAIN_ALL)
        mode = _lib.M_FLOW_ACTIVE | _lib.M_APPLY_ADAM
        losses = []
        for it in range(3):
            adam = adam_schedule([0.01], start_step=it)
            lib.emul_prepare(C.byref(st))
            lib.emul_accumulate(C.byref(st), mode)
            dist.all_reduce(bufs.acc, op=dist.ReduceOp.SUM)          # the only exchange step of an iteration
            lib.emul_update(C.byref(st), mode, adam)
            losses.append(float(bufs.scal[_lib.S_LOSS]))
        small = {k: bufs.params[k].clone() for k in ("pw_poses", "im_poses", "im_focals")}
        gathered = [None] * world
        dist.all_gather_object(gathered, (wins[rank], losses, small, bufs.params["im_depthmaps"][:, :prob.P].clone()))
        if rank == 0:
            torch.save(gathered, out)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("name", ["small_flow"])
def test_sharded_iterations_equal_single_shard(name, world, tmp_path):
    """2 ranks (one boundary) and 3 ranks (the middle window has slots of boundary edges on both sides)."""
    emul_lib()                                   # build once before forking
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(world, _free_port(), name, out), nprocs=world, join=True)
    res = torch.load(out, weights_only=False)

    fx, scene, prob = golden_case(name)
    bufs = cpu_problem(prob, fx["init_params"])
    ref_losses = []
    for it in range(3):
        l, _, _ = emul_eval(bufs, True, apply=True, lr=0.01, step=it)
        ref_losses.append(l)

    wins = [r[0] for r in res]
    assert wins[0][0] == 0 and wins[-1][1] == prob.n_imgs and all(wins[r][1] == wins[r + 1][0] for r in range(world - 1))
    assert all(b > a for a, b in wins)
    l0, s0 = res[0][1], res[0][2]
    for _, lr_, sr, _ in res[1:]:
        assert lr_ == l0                                             # replicated state stays bit-identical
        for k in s0:
            assert torch.equal(s0[k], sr[k]), k
    for k in s0:
        assert torch.allclose(s0[k], bufs.params[k], rtol=0, atol=2e-6), k
    assert np.allclose(l0, ref_losses, rtol=1e-9, atol=0)
    depth = torch.cat([r[3] for r in res])
    assert torch.allclose(depth, bufs.params["im_depthmaps"][:, :prob.P], rtol=0, atol=2e-6)
