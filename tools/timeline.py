"""Diagnostic: the launch timeline of the hot loop as the GPU saw it (a3r_debug_timeline: CTA 0 of every scan / update launch
stamps its entry, the return of its programmatic-launch wait and its exit with %globaltimer).

    python tools/timeline.py [--workload c2] [--iters 12] [--frames N --height H --width W]
"""
import argparse, ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from align3r_b200.cloud_opt_flow import global_aligner
from align3r_b200.cloud_opt_flow.optimizer import global_alignment_loop
from align3r_b200.synth import CANONICAL_KW, CONFIGS, make_scene

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c2")
ap.add_argument("--iters", type=int, default=12)
ap.add_argument("--frames", type=int, default=0)
ap.add_argument("--height", type=int, default=0)
ap.add_argument("--width", type=int, default=0)
a = ap.parse_args()
n, H, W, graph = CONFIGS[a.workload]
n, H, W = a.frames or n, a.height or H, a.width or W
sc = make_scene(n, H, W, graph, seed=0, with_flow=True, workers=os.cpu_count() or 1)
net = global_aligner(sc.output, "cuda", flow=(sc.flow_ij, sc.flow_ji), **dict(CANONICAL_KW, num_total_iter=a.iters))
net.verbose = False
net.compute_global_alignment(init="mst", niter=0)
global_alignment_loop(net, lr=0.01, niter=a.iters, schedule="linear")          # warm
lib = net._lib
lib.a3r_debug_k2_trace(1, None)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
global_alignment_loop(net, lr=0.01, niter=a.iters, schedule="linear")
e1.record()
torch.cuda.synchronize()
buf = (C.c_ulonglong * 256)()
cnt = C.c_int()
lib.a3r_debug_timeline(buf, C.byref(cnt))
lib.a3r_debug_k2_trace(0, None)
rows = [[buf[4 * k + j] for j in range(4)] for k in range(cnt.value)][-2 * a.iters:]
t0 = rows[0][1]
print(f"{n} frames {H}x{W}: {a.iters} iterations, {1e3 * e0.elapsed_time(e1) / a.iters:.1f} us / iteration (CUDA events)")
print("kernel   entry us   wait returned   exit   | in kernel after the wait | from the previous kernel's exit to this wait's return")
prev_exit = None
for kid, t_in, t_wait, t_out in rows:
    name = {1: "scan  ", 2: "update"}.get(int(kid), "?")
    gap = "" if prev_exit is None else f"{(t_wait - prev_exit) / 1e3:8.2f}"
    print(f"{name} {(t_in - t0) / 1e3:10.2f} {(t_wait - t0) / 1e3:12.2f} {(t_out - t0) / 1e3:10.2f} | {(t_out - t_wait) / 1e3:10.2f} | {gap}")
    prev_exit = t_out
