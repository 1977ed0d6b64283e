"""Diagnostic (not a test): the accumulate scan of workload C2 alone - a cheap target for timing and ncu captures.

    python tools/scan_only.py [--reps 20] [--mode flow|align|both] [--workload c2]
    ncu --set full --clock-control none --import-source on -k regex:k_accumulate --launch-skip 3 -c 1 \
        -o gpurun_out/prof python tools/scan_only.py --reps 2 --mode flow
"""
import argparse, ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from align3r_b200 import _lib
from align3r_b200.cloud_opt_flow import global_aligner
from align3r_b200.synth import CANONICAL_KW, CONFIGS, load_params, make_edge_list, make_scene, make_scene_device, near_solution_params

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--mode", default="both")
ap.add_argument("--workload", default="c2")
ap.add_argument("--world", type=int, default=1, help="time the scan of ONE frame-window shard of `world` (rank-emulated on this GPU)")
ap.add_argument("--rank", type=int, default=0)
ap.add_argument("--folded", action="store_true", help="the scan as the loop runs it: the previous iteration's depth Adam step in the tile prologue")
ap.add_argument("--recipe", action="store_true", help="device-drawn inputs (only this shard's edges) + parameters by recipe "
                "(synth.near_solution_params) instead of host inputs + MST initialisation: the 1000-frame workload")
a = ap.parse_args()
n, H, W, graph = CONFIGS[a.workload]
if a.recipe:
    import numpy as np
    from align3r_b200.plan import frame_costs, partition_frames
    all_edges = make_edge_list(n, graph)
    rows, shard = None, None
    if a.world > 1:
        lo, hi = partition_frames(frame_costs(np.asarray(all_edges, dtype=np.int32), np.asarray([(H, W)] * n, dtype=np.int32)), a.world)[a.rank]
        rows = [e for e, (i, j) in enumerate(all_edges) if lo <= i < hi or lo <= j < hi]
        shard = dict(world=a.world, rank=a.rank, group=None)
    sc = make_scene_device(n, H, W, graph, seed=0, device="cuda", edges=None if rows is None else [all_edges[e] for e in rows])
    out = sc.output
    if rows is not None:
        out = dict(out, view1=dict(out["view1"], idx=[i for i, _ in all_edges]), view2=dict(out["view2"], idx=[j for _, j in all_edges]))
    net = global_aligner(out, "cuda", flow=(sc.flow_ij, sc.flow_ji), shard=shard, edge_rows=rows, **dict(CANONICAL_KW))
    net.verbose = False
    load_params(net, *near_solution_params(sc.gt, all_edges, seed=0))
    net._join_uploads()
    a.world = 1          # built: skip the host-input shard path below
else:
    sc = make_scene(n, H, W, graph, seed=0, with_flow=True)
if not a.recipe:
    net = global_aligner(sc.output, "cuda", flow=(sc.flow_ij, sc.flow_ji), **dict(CANONICAL_KW))
    net.verbose = False
    net.compute_global_alignment(init="mst", niter=0)
if a.world > 1:        # one shard of a `world`-way frame-window partition, with the full problem's initial parameters
    full = net
    net = global_aligner(sc.output, "cuda", flow=(sc.flow_ij, sc.flow_ji), shard=dict(world=a.world, rank=a.rank, group=None),
                         **dict(CANONICAL_KW))
    net.verbose = False
    with torch.no_grad():
        for k in ("pw_poses", "im_poses", "im_focals"):
            getattr(net, k).data.copy_(getattr(full, k).data)
        net.im_depthmaps.data.copy_(full.im_depthmaps.data[torch.as_tensor(net._plan.lframe_id.tolist(), device="cuda")])
    net._join_uploads() if hasattr(net, "_join_uploads") else None
    del full
st = net._bufs.struct(net._trainable())
stream = net._stream()
ps, nl, P = net._plan.pixel_slots(), net._plan.n_local_frames, H * W
from align3r_b200.problem import adam_schedule
prev = adam_schedule([0.0], start_step=10, first_iter=-1)       # lr 0: repeated depth steps leave the planes alone
extra = 24 if a.folded else 0
for label, mode, nbytes in (("flow", _lib.M_FLOW_ACTIVE, ps * 24 + nl * P * (5 + extra)), ("align", 0, ps * 16 + nl * P * (4 + extra))):
    if a.mode not in ("both", label):
        continue

    def launch():
        if a.folded:
            return net._lib.a3r_ga_accumulate_folded(C.byref(st), mode, prev, mode, stream)
        return net._lib.a3r_ga_accumulate(C.byref(st), mode, stream)
    _lib.check(net._lib.a3r_ga_prepare(C.byref(st), mode, stream))
    for _ in range(3):
        _lib.check(launch())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        _lib.check(launch())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    print(f"scan[{label}] {ms:.4f} ms/launch  {nbytes / ms / 1e6:.1f} GB/s algorithmic  ({nbytes / ms / 1e6 / 6549.4:.3f} of 6549.4 GB/s measured peak)", flush=True)
