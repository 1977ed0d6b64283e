"""Control experiment for the 1-GPU vs 8-GPU loss-history comparison (tools/c5_compare.py): the SAME workload on ONE GPU
with two different partitions of the tiles among the scan's warps (the plan's n_ctas: 2368 = 148 SMs x 16 warps, and half of
it).  The per-warp fp32 running sums then cover different tile subsets - exactly what changes between one GPU and a
frame-window shard - and nothing else does.  The free-running loss histories drift apart at the same rate: the drift is
Adam amplifying fp32 summation-order noise (hard part 3), not the sharding.

    python tools/partition_sensitivity.py [--workload c2] [--iters 24]
"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from align3r_b200.cloud_opt_flow import global_aligner
from align3r_b200.synth import CANONICAL_KW, CONFIGS, make_scene

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c2")
ap.add_argument("--iters", type=int, default=24)
a = ap.parse_args()
n, H, W, graph = CONFIGS[a.workload]
sc = make_scene(n, H, W, graph, seed=0, with_flow=True, workers=os.cpu_count() or 1)
hist = {}
for ctas in (None, 1184, 592):
    if ctas is None:
        os.environ.pop("A3R_PLAN_CTAS", None)
    else:
        os.environ["A3R_PLAN_CTAS"] = str(ctas)
    net = global_aligner(sc.output, "cuda", flow=(sc.flow_ij, sc.flow_ji), **dict(CANONICAL_KW, num_total_iter=a.iters))
    net.verbose = False
    net.compute_global_alignment(init="mst", niter=a.iters, schedule="linear", lr=0.01)
    hist[ctas or net._plan.n_ctas] = [float(x) for x in net._bufs.loss_hist[:a.iters].cpu()]
    del net
    torch.cuda.empty_cache()
keys = list(hist)
print(f"{a.workload}: free-running loss history with the tiles cut into {keys} warp ranges (one GPU, identical inputs and initialisation)")
print("  it " + "  ".join(f"{k:>12d}" for k in keys) + "   max rel. difference")
worst = 0.0
for it in range(a.iters):
    v = [hist[k][it] for k in keys]
    d = max(abs(x - v[0]) / abs(v[0]) for x in v)
    worst = max(worst, d)
    print(f"  {it:2d} " + "  ".join(f"{x:12.8f}" for x in v) + f"   {d:.1e}")
print(f"max over {a.iters} iterations: {worst:.1e}")
