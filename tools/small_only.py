"""Diagnostic (not a test): the small-parameter kernel K2 (k_small) with workload C2's graph (50 frames, 470 edges) on small
images - K2's work does not depend on the image size - as a cheap target for timing and ncu captures.

    python tools/small_only.py [--iters 12]
    ncu --set full --clock-control none --import-source on -k k_small --launch-skip 6 -c 1 -o gpurun_out/prof \
        python tools/small_only.py
"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from align3r_b200.cloud_opt_flow import global_aligner
from align3r_b200.cloud_opt_flow.optimizer import global_alignment_loop
from align3r_b200.synth import CANONICAL_KW, make_scene

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=12)
ap.add_argument("--frames", type=int, default=50)
a = ap.parse_args()
sc = make_scene(a.frames, 32, 64, "swin-5-noncyclic", seed=0, with_flow=True)
net = global_aligner(sc.output, "cuda", flow=(sc.flow_ij, sc.flow_ji), **dict(CANONICAL_KW, num_total_iter=a.iters))
net.verbose = False
net.compute_global_alignment(init="mst", niter=0)
torch.cuda.synchronize()
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    loss = global_alignment_loop(net, lr=0.01, niter=a.iters, schedule="linear")
    e1.record()
    torch.cuda.synchronize()
    print(f"rep {rep}: {a.iters} iterations, {1e3 * e0.elapsed_time(e1) / a.iters:.1f} us / iteration (tiny images: the small-parameter chain dominates)  loss {loss:.6f}")

# phase trace of the last K2 launch (a3r_debug_k2_trace): nanoseconds relative to the earliest CTA's start
import ctypes as C
lib = net._lib
buf = (C.c_ulonglong * 64)()
lib.a3r_debug_k2_trace(2 + a.iters // 2, None)          # a mid-loop launch (it also writes the next iteration's parameter maps)
global_alignment_loop(net, lr=0.01, niter=a.iters, schedule="linear")
lib.a3r_debug_k2_trace(0, buf)
rows = [[buf[c * 8 + k] for k in range(8)] for c in range(8)]
t0 = min(r[0] for r in rows)
print("K2 phase trace (us since the first CTA started): start | gate | phase A1 | barrier 1 | phase A2 | barrier 2 | phase C | end")
for c, r in enumerate(rows):
    print(f"  cta {c}: " + "  ".join(f"{(x - t0) / 1e3:7.2f}" for x in r))
