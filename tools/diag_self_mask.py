"""Diagnostic (not a test): self-estimated masks against the fixture of the real reference, term by term."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from tests.helpers import golden_case, go
from tests.test_gpu_parity import build
for tag in ("a", "b"):
    fx, scene, prob = golden_case(f"small_self_mask_{tag}")
    net = build(scene, fx["kw"])
    ours = torch.stack(net.dynamic_masks).cpu()
    ref = fx["dynamic_masks"]
    print(tag, "mismatch per frame", (ours != ref).sum(dim=(1, 2)).tolist())
    pc = net._pair_cameras
    half = len(scene.edges) // 2
    fo = pc["focals"]
    print("  focal diff i", float(np.abs(fo[:half] - fx["cams"]["K_i"][:, 0, 0].numpy()).max()), "j", float(np.abs(fo[half:] - fx["cams"]["K_j"][:, 0, 0].numpy()).max()))
    for k in pc["pairs"]:
        first, pose = pc["rel"][k]
        Rr, Tr = (fx["cams"]["R_j"][k], fx["cams"]["T_j"][k]) if first else (fx["cams"]["R_i"][k], fx["cams"]["T_i"][k])
        print("  pair", k, "first", first, "dR", float((torch.from_numpy(pose[:3, :3]) - Rr).abs().max()), "dT", float((torch.from_numpy(pose[:3, 3:]) - Tr).abs().max()))
    _, mean = go.motion_masks_from_pairs(prob)
    bad = (ours != ref)
    print("  oracle mean at mismatching pixels:", mean[bad][:20].tolist())
