"""TEST INFRASTRUCTURE - generates tests/golden/small_self_mask_{a,b}.pt by running the REAL reference (CPU).

    python oracle/make_golden_self_mask.py          (build container only: needs /root/reference)

use_self_mask=True: PointCloudOptimizer.get_motion_mask_from_pairs (reference optimizer.py:156-235) estimates the dynamic
masks from the symmetric pairs themselves (PairViewer cameras, ego flow vs optical flow).  Stored: the masks, the two-view
cameras / depth maps the reference derived for every pair, and the outcome of a short free-running alignment with those masks.
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

CASES = {"a": (dict(n_frames=6, H=48, W=64, scene_graph="swin-2-noncyclic", seed=0, with_flow=True), 20),
         "b": (dict(n_frames=7, H=40, W=56, scene_graph="swin-3-noncyclic", seed=1, with_flow=True), 20)}


def main():
    rh.import_reference()
    for tag, (recipe, niter) in CASES.items():
        scene = make_scene(**recipe)
        kw = dict(CANONICAL_KW, num_total_iter=niter, use_self_mask=True)
        ref = rh.build_reference_scene(scene, **kw)
        masks = torch.stack(ref.dynamic_masks).clone()
        cams = dict(K_i=ref.intrinsics_i.clone(), K_j=ref.intrinsics_j.clone(), R_i=ref.R_i.clone(), R_j=ref.R_j.clone(),
                    T_i=ref.T_i.clone(), T_j=ref.T_j.clone())
        with rh.quiet():
            final = ref.compute_global_alignment(init="mst", niter=niter, schedule="linear", lr=0.01)
        out = dict(recipe=recipe, kw=kw, niter=niter, checksum=checksum(scene), edges=scene.edges, dynamic_masks=masks, cams=cams,
                   free_run=dict(loss=float(final), flow_dropped=bool(ref.flow_loss_flag), params=params_of(ref),
                                 im_poses_4x4=ref.get_im_poses().detach().clone(), focals=ref.get_focals().detach().clone()))
        path = os.path.join(ROOT, "tests", "golden", f"small_self_mask_{tag}.pt")
        torch.save(out, path)
        print(tag, "->", path, f"{os.path.getsize(path) / 1024:.0f} KB  dynamic fraction {float(masks.float().mean()):.4f}  "
              f"final loss {final}  dropped {out['free_run']['flow_dropped']}")


if __name__ == "__main__":
    main()
