"""Diagnostic (not a test): per-phase clock64 timestamps of CTA 0 from a -DA3R_V2_TRACE build."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from align3r_b200 import _lib
from align3r_b200.cloud_opt_flow import global_aligner
from align3r_b200.synth import make_scene, CANONICAL_KW
sc = make_scene(50, 288, 512, "swin-5-noncyclic", seed=0, with_flow=True)
net = global_aligner(sc.output, "cuda", flow=(sc.flow_ij, sc.flow_ji), **dict(CANONICAL_KW))
net.compute_global_alignment(init="mst", niter=0)
for mode in (_lib.M_FLOW_ACTIVE, 0):
    for rep in range(3):
        net._phase_accumulate(mode)
    torch.cuda.synchronize()
    t = net._bufs.edge_scratch[:36 * 12].cpu().view(torch.int64).view(36, 12)
    print("mode", mode, "cols: S.start S.issued S.pixready S.full S.Bdone S.arrived | P.Astart P.Adone P.slotsdone P.stagefree P.Cdone")
    for i in range(4, 24):
        r = t[i]
        print(i, " ".join(f"{int(v):7d}" for v in r[:11]))
