"""Diagnostic (not a test): the key-frame alignment of the hierarchical fixture, product vs oracle, iteration by iteration."""
import os, sys, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.helpers import go, make_scene
from tests.test_gpu_parity import set_params
from align3r_b200.cloud_opt import global_aligner
from align3r_b200.cloud_opt_flow import GlobalAlignerMode
N, H, W = 11, 32, 48
sc = make_scene(N, H, W, "complete", seed=4, with_flow=False, edges=[(0, 4), (0, 8), (4, 8)])
out = copy.deepcopy(sc.output)
out["view1"]["idx"] = [0, 0, 1]; out["view2"]["idx"] = [1, 2, 2]
for p in ("pred1", "pred2"):
    c = out[p]["conf"]; c[c > 1] = 10
prob = go.problem_from_output(out, None, None, min_conf_thr=3.0)
params, _ = go.mst_init(prob)
_, losses, _ = go.run_alignment(prob, params, niter=30, schedule="cosine", lr=0.05, lr_min=1e-6)
net = global_aligner(copy.deepcopy(out), False, [], "cuda", mode=GlobalAlignerMode.PointCloudOptimizer, verbose=False, min_conf_thr=3)
net.compute_global_alignment(init="mst", niter=0)
for k in ("pw_poses", "im_poses", "im_focals"):
    print(k, "init diff", float((getattr(net, k).detach().cpu().reshape(params[k].shape) - params[k]).abs().max()))
print("depth init diff", float((net.im_depthmaps.detach().cpu()[:, :prob.P] - params["im_depthmaps"]).abs().max()))
l_own = net.compute_global_alignment(init=None, niter=30, schedule="cosine", lr=0.05)
own = net._bufs.loss_hist[:30].cpu().tolist()
net2 = global_aligner(copy.deepcopy(out), False, [], "cuda", mode=GlobalAlignerMode.PointCloudOptimizer, verbose=False, min_conf_thr=3)
set_params(net2, params)
net2.compute_global_alignment(init=None, niter=30, schedule="cosine", lr=0.05)
same = net2._bufs.loss_hist[:30].cpu().tolist()
for it in range(30):
    print(it, f"oracle {losses[it]:.6f}  product(same init) {same[it]:.6f}  product(own mst) {own[it]:.6f}")
