"""Diagnostic (not a test): cost of the self-estimated mask pre-step (use_self_mask=True) on workload C2."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["A3R_TIMING"] = "1"
import torch
from align3r_b200.cloud_opt_flow import global_aligner
from align3r_b200.synth import make_scene, CANONICAL_KW, CONFIGS
n, H, W, graph = CONFIGS["c2"]
sc = make_scene(n, H, W, graph, seed=0, with_flow=True, pin_memory=True)
for rep in range(2):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    net = global_aligner(sc.output, "cuda", flow=(sc.flow_ij, sc.flow_ji), **dict(CANONICAL_KW, use_self_mask=True))
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    dyn = torch.stack(net.dynamic_masks)
    gt = torch.stack([torch.as_tensor(m) for m in sc.dynamic_masks]).to(dyn.device) if hasattr(sc, "dynamic_masks") else None
    msg = f"rep {rep}: construct with self masks {t1 - t0:.3f} s  dynamic fraction {float(dyn.float().mean()):.4f}"
    if gt is not None:
        inter = float((dyn & gt.bool()).sum()); union = float((dyn | gt.bool()).sum())
        msg += f"  IoU with the generator's true dynamic rectangles {inter / max(union, 1):.3f}"
    print(msg, flush=True)
    del net
