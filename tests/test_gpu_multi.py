"""Frame-window sharding over real GPUs with NCCL (needs >= 2 visible GPUs; skipped on a 1-GPU box, where
test_gpu_parity.py::test_rank_emulated_frame_window_sharding and tests/test_dist_gloo.py cover the same logic)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

from tests.helpers import golden_case, rel_err

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("exchange", ["peer", "nccl"])
@pytest.mark.parametrize("name", ["small_flow"])
def test_sharded_alignment_matches_reference(cuda_device, name, exchange, tmp_path):
    """exchange='peer': the accumulator block crosses GPUs inside our own kernel over CUDA-IPC peer memory
    (a3r_ga_run_peers); 'nccl': torch.distributed all-reduce between the scan and the update."""
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    out = str(tmp_path / "res.pt")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dist_gpu_worker.py"), name, out]
    subprocess.run(cmd, check=True, timeout=600, cwd=ROOT, env=dict(os.environ, A3R_EXCHANGE=exchange))
    res = torch.load(out, weights_only=False)
    assert res["exchange"] == exchange
    fx, scene, prob = golden_case(name)
    parts = sorted(res["parts"], key=lambda t: t[0])
    assert parts[0][0] == 0 and parts[-1][1] == prob.n_imgs and all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
    assert all(p[3] for p in parts), "replicated parameters diverged between ranks"
    assert len({p[4] for p in parts}) == 1, "ranks disagree on the loss"
    ref = fx["free_run"]
    assert parts[0][5] == ref["flow_dropped"]
    depth = torch.cat([p[2] for p in parts])
    # same end-state tolerances as the single-GPU canonical call (test_free_running_alignment)
    assert abs(parts[0][4] - ref["loss"]) <= 2e-3 * abs(ref["loss"])
    assert float((depth - ref["params"]["im_depthmaps"]).abs().median()) < 2e-3
    assert float((res["im_poses"] - ref["im_poses_4x4"]).abs().max()) < 1e-2
    assert rel_err(res["focals"], ref["focals"]) < 5e-3


def test_sharded_self_estimated_masks(cuda_device, tmp_path):
    """use_self_mask=True on frame windows: every rank estimates the dynamic masks of its own frames from the slots living on
    them (no exchange beyond the focals / confidences the initialisation already shares); together they are the reference's."""
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    out = str(tmp_path / "res.pt")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dist_gpu_worker.py"), "small_self_mask_a", out]
    subprocess.run(cmd, check=True, timeout=600, cwd=ROOT, env=dict(os.environ))
    res = torch.load(out, weights_only=False)
    fx, scene, prob = golden_case("small_self_mask_a")
    parts = sorted(res["parts"], key=lambda t: t[0])
    masks = torch.cat([p[6] for p in parts])
    assert masks.shape == fx["dynamic_masks"].shape
    assert int((masks != fx["dynamic_masks"]).sum()) <= 1e-3 * masks.numel()
    assert all(p[3] for p in parts) and len({p[4] for p in parts}) == 1
    assert abs(parts[0][4] - fx["free_run"]["loss"]) <= 1e-2 * abs(fx["free_run"]["loss"])


def test_sharded_readouts_equal_the_single_device_ones(cuda_device, tmp_path):
    """clean_pointcloud (reference base_opt.py:587-623, the step right after the loop in tool/pose_test.py:202-216) and the
    flow validity maps (OccMask, goem_opt.py:575-618) on frame windows: bit-equal to the single-device product path at the
    same parameters (which tests/test_gpu_parity.py pins on the reference's fixtures)."""
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    from align3r_b200.cloud_opt_flow import GlobalAlignerMode, global_aligner
    from align3r_b200.synth import load_params, near_solution_params
    out = str(tmp_path / "res.pt")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dist_gpu_worker.py"), "small_flow", out, "readout"]
    subprocess.run(cmd, check=True, timeout=600, cwd=ROOT, env=dict(os.environ))
    res = torch.load(out, weights_only=False)
    fx, scene, prob = golden_case("small_flow")
    net = global_aligner(scene.output, cuda_device, mode=GlobalAlignerMode.PointCloudOptimizer, flow=(scene.flow_ij, scene.flow_ji),
                         **dict(fx["kw"]))
    load_params(net, *near_solution_params(scene.gt, scene.edges, seed=0))
    before = torch.stack([c.detach().cpu().clone() for c in net.im_conf])
    net.clean_pointcloud()
    conf = torch.stack([c.detach().cpu() for c in net.im_conf])
    assert int((conf != before).sum()) > 0, "the clean-up changed nothing: the test scene does not exercise it"
    parts = sorted(res["parts"], key=lambda t: t[0])
    assert torch.equal(torch.cat([p[2] for p in parts]), conf)
    vi, vj = net.flow_valid_mask_i.cpu(), net.flow_valid_mask_j.cpu()
    seen_i, seen_j = set(), set()
    for _, _, _, rows_i, part_i, rows_j, part_j in parts:
        assert torch.equal(part_i, vi[rows_i]) and torch.equal(part_j, vj[rows_j])
        seen_i |= set(rows_i)
        seen_j |= set(rows_j)
    assert seen_i == set(range(prob.n_edges)) and seen_j == set(range(prob.n_edges))


def test_sharded_known_poses_initialisation(cuda_device, tmp_path):
    """init='known_poses' (reference init_im_poses.py:44-85) on frame windows: the rank owning a pair's first frame solves it,
    pred_j planes of boundary pairs are fetched from the neighbour - same result as the real reference's fixture."""
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    from tests.helpers import go
    out = str(tmp_path / "res.pt")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dist_gpu_worker.py"), "small_known_poses", out, "known_poses"]
    subprocess.run(cmd, check=True, timeout=600, cwd=ROOT, env=dict(os.environ))
    res = torch.load(out, weights_only=False)
    fx, scene, prob = golden_case("small_known_poses")
    ref = fx["known_poses"]["params"]
    parts = sorted(res["parts"], key=lambda t: t[0])
    prob.norm_pw_scale = False
    T_ref = go.pw_pose_matrices(prob, ref["pw_poses"])
    for p in parts:                                            # replicated on every rank
        assert float((p[3] - T_ref).abs().max()) < 2e-3
        assert float((p[4] - fx["known_c2w"]).abs().max()) < 1e-5
    depth = torch.cat([p[2] for p in parts])
    assert float((depth - ref["im_depthmaps"]).abs().median()) < 1e-5
