"""TEST INFRASTRUCTURE - generates tests/golden/small_clean_pointcloud.pt by running the REAL reference (CPU).

    python oracle/make_golden_clean.py          (build container only: needs /root/reference)

clean_pointcloud (reference base_opt.py:293-303, 587-623) after a short free run: the inputs it read (poses, intrinsics,
depth maps, world pointmaps, confidences) and the confidences it left.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_harness as rh                                   # noqa: E402
from align3r_b200.synth import make_scene, CANONICAL_KW   # noqa: E402
from make_golden import checksum, params_of               # noqa: E402

RECIPE = dict(n_frames=6, H=48, W=64, scene_graph="swin-2-noncyclic", seed=13, with_flow=True)
NITER = 15


def main():
    rh.import_reference()
    scene = make_scene(**RECIPE)
    kw = dict(CANONICAL_KW, num_total_iter=NITER)
    ref = rh.build_reference_scene(scene, **kw)
    with rh.quiet():
        ref.compute_global_alignment(init="mst", niter=NITER, schedule="linear", lr=0.01)
    out = dict(recipe=RECIPE, kw=kw, niter=NITER, checksum=checksum(scene), edges=scene.edges, params=params_of(ref))
    with torch.no_grad():
        out["before"] = dict(im_conf=torch.stack([c.detach().clone() for c in ref.im_conf]),
                             K=ref.get_intrinsics().detach().clone(), c2w=ref.get_im_poses().detach().clone(),
                             depth=torch.stack([d.detach().clone() for d in ref.get_depthmaps()]),
                             pts3d=torch.stack([p.detach().clone() for p in ref.get_pts3d()]))
        for tol in (0.001, 0.05):
            r2 = rh.build_reference_scene(scene, **kw)
            for k, v in out["params"].items():
                getattr(r2, k).data.copy_(v)
            r2.clean_pointcloud(tol=tol)
            out[f"after_tol{tol}"] = torch.stack([c.detach().clone() for c in r2.im_conf])
    path = os.path.join(ROOT, "tests", "golden", "small_clean_pointcloud.pt")
    torch.save(out, path)
    b = out["before"]["im_conf"]
    print("->", path, f"{os.path.getsize(path) / 1024:.0f} KB", {k: int((out[k] != b).sum()) for k in out if k.startswith("after")}, "of", b.numel())


if __name__ == "__main__":
    main()
