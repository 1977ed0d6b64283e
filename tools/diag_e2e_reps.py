"""Diagnostic: repeat the end-to-end C2 job in one process (A3R_TIMING=1 prints the host-side phases)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from align3r_b200.cloud_opt_flow import global_aligner
from align3r_b200.synth import make_scene, CANONICAL_KW
sc = make_scene(50, 288, 512, "swin-5-noncyclic", seed=0, with_flow=True, pin_memory=True)
res = torch.empty((50, 147456), dtype=torch.float32, pin_memory=True)
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 5):
    import gc; gc.disable(); torch.cuda.synchronize(); t0 = time.perf_counter()
    net = global_aligner(sc.output, "cuda", flow=(sc.flow_ij, sc.flow_ji), **dict(CANONICAL_KW)); net.verbose = False
    t1 = time.perf_counter()
    net.compute_global_alignment(init="mst", niter=0); t2 = time.perf_counter()
    loss = net.compute_global_alignment(init=None, niter=300, schedule="linear", lr=0.01); t3 = time.perf_counter()
    res.copy_(net.get_depthmaps(raw=True).detach(), non_blocking=True); p = net.get_im_poses().detach().cpu(); torch.cuda.synchronize(); t4 = time.perf_counter()
    gc.enable()
    print(f"rep {rep}: construct {t1-t0:.3f}  mst-init {t2-t1:.3f}  loop(300) {t3-t2:.3f}  readback {t4-t3:.3f}  total {t4-t0:.3f} -> {300/(t4-t0):.0f} it/s  loss {loss:.5f}", flush=True)
    del net
    import gc; gc.collect()
