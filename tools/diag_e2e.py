"""Diagnostic (not a test): where the end-to-end time of a C2 job goes (construct / MST init / loop / read-back)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from align3r_b200.cloud_opt_flow import global_aligner
from align3r_b200.cloud_opt_flow import init_im_poses as init_fun
from align3r_b200.cloud_opt_flow import optimizer as opt
from align3r_b200.synth import make_scene, CANONICAL_KW
import cProfile, pstats

t = time.perf_counter()
sc = make_scene(50, 288, 512, "swin-5-noncyclic", seed=0, with_flow=True, pin_memory=True)
print("scene gen", time.perf_counter() - t)
torch.cuda.synchronize()
for rep in range(3):
    prof = rep == 2
    pr = cProfile.Profile()
    t0 = time.perf_counter()
    if prof: pr.enable()
    net = global_aligner(sc.output, "cuda", flow=(sc.flow_ij, sc.flow_ji), **dict(CANONICAL_KW))
    net.verbose = False
    torch.cuda.synchronize(); t1 = time.perf_counter()
    if prof:
        pr.disable(); print("---- construct"); pstats.Stats(pr).sort_stats("cumulative").print_stats(14); pr = cProfile.Profile(); pr.enable()
    init_fun.init_minimum_spanning_tree(net)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    if prof:
        pr.disable(); print("---- mst"); pstats.Stats(pr).sort_stats("cumulative").print_stats(14); pr = cProfile.Profile()
    loss = opt.global_alignment_loop(net, lr=0.01, niter=300, schedule="linear")
    torch.cuda.synchronize(); t3 = time.perf_counter()
    if prof: pr.enable()
    d = net.get_depthmaps(raw=True).detach().cpu(); torch.cuda.synchronize(); t3a = time.perf_counter()
    p = net.get_im_poses().detach().cpu(); f = net.get_focals().detach().cpu(); h = net._bufs.loss_hist[:300].cpu()
    torch.cuda.synchronize(); t4 = time.perf_counter()
    if prof:
        pr.disable(); print("---- readback"); pstats.Stats(pr).sort_stats("cumulative").print_stats(10)
    print(f"rep {rep}: construct {t1-t0:.3f}  mst-init {t2-t1:.3f}  loop(300) {t3-t2:.3f}  readback depth {t3a-t3:.3f} rest {t4-t3a:.3f}  total {t4-t0:.3f}  loss {loss:.5f}", flush=True)
    del net
