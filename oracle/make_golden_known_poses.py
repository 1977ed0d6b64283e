"""TEST INFRASTRUCTURE - generates tests/golden/small_known_poses.pt by running the REAL reference (CPU).

    python oracle/make_golden_known_poses.py          (build container only: needs /root/reference)

Two initialisations with preset camera poses:
  * ``mst_frozen``: preset_pose(all poses, frozen) then init_minimum_spanning_tree - the align_multiple_poses branch of
    init_from_pts3d (reference init_im_poses.py:109-122);
  * ``known_poses``: preset_pose(all poses, requires_grad=True) then init_from_known_poses (init_im_poses.py:44-85) -
    what compute_global_alignment(init='known_poses') runs (base_opt.py:459-474).
The known poses are the synthetic scene's ground-truth cameras, rescaled and moved by a fixed similarity so that the
global alignment has something to do.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_harness as rh                                   # noqa: E402
from align3r_b200.synth import make_scene, CANONICAL_KW   # noqa: E402
from make_golden import NAMES, checksum, params_of        # noqa: E402

RECIPE = dict(n_frames=6, H=48, W=64, scene_graph="swin-2-noncyclic", seed=7, with_flow=True)


def known_poses_of(scene):
    """Ground-truth cam-to-world poses in another world frame: x -> 1.7 Rz(0.3) x + (0.4, -0.2, 0.1)."""
    c, s = np.cos(0.3), np.sin(0.3)
    S = np.eye(4)
    S[:3, :3] = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])
    out = []
    for P in scene.gt.cam2world:
        Q = S @ np.asarray(P, dtype=np.float64)
        Q[:3, 3] = 1.7 * Q[:3, 3] + np.array([0.4, -0.2, 0.1])
        out.append(Q)
    return torch.from_numpy(np.stack(out)).float()


def main():
    rh.import_reference()
    import dust3r.cloud_opt_flow.init_im_poses as init_fun
    scene = make_scene(**RECIPE)
    kw = dict(CANONICAL_KW, num_total_iter=20)
    known = known_poses_of(scene)
    out = dict(recipe=RECIPE, kw=kw, checksum=checksum(scene), edges=scene.edges, known_c2w=known)

    ref = rh.build_reference_scene(scene, **kw)
    with rh.quiet():
        ref.preset_pose(known)                               # frozen: get_known_poses() sees N known poses
        init_fun.init_minimum_spanning_tree(ref, niter_PnP=10)
    out["mst_frozen"] = dict(params=params_of(ref), im_poses_requires_grad=bool(ref.im_poses.requires_grad),
                             norm_pw_scale=bool(ref.norm_pw_scale))

    ref = rh.build_reference_scene(scene, **kw)
    with rh.quiet():
        ref.preset_pose(known, requires_grad=True)           # what compute_global_alignment(init='known_poses') does
        init_fun.init_from_known_poses(ref, min_conf_thr=ref.min_conf_thr, niter_PnP=10)
    out["known_poses"] = dict(params=params_of(ref), im_poses_requires_grad=bool(ref.im_poses.requires_grad),
                              focals=ref.get_focals().detach().clone())
    path = os.path.join(ROOT, "tests", "golden", "small_known_poses.pt")
    torch.save(out, path)
    print("->", path, f"{os.path.getsize(path) / 1024:.0f} KB")


if __name__ == "__main__":
    main()
