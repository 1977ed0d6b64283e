"""TEST INFRASTRUCTURE - generates tests/golden/small_init_priors_{a,b}.pt by running the REAL reference (CPU).

    python oracle/make_golden_priors.py          (build container only: needs /root/reference)

init_minimum_spanning_tree(init_priors=[key pose, key depth, [key focal]]) - the start of the hierarchical clip driver
from the key frame 0 (reference init_im_poses.py:173-215, tool/demo.py / depth_test.py clip-wise calls).
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
from make_golden import checksum, params_of               # noqa: E402

CASES = {"a": dict(n_frames=6, H=48, W=64, scene_graph="swin-2-noncyclic", seed=17, with_flow=True),
         "b": dict(n_frames=6, H=48, W=64, scene_graph="swin-3-noncyclic", seed=30, with_flow=True)}   # first edge (2*,0*): the j == 0 branch


def key_prior(scene):
    c, s = np.cos(-0.2), np.sin(-0.2)
    P = np.eye(4, dtype=np.float32)
    P[:3, :3] = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], dtype=np.float32)
    P[:3, 3] = (0.3, 0.1, -0.2)
    return [P, None, [0.85 * scene.meta["W"]]]


def main():
    rh.import_reference()
    import dust3r.cloud_opt_flow.init_im_poses as init_fun
    for tag, recipe in CASES.items():
        scene = make_scene(**recipe)
        kw = dict(CANONICAL_KW, num_total_iter=20)
        pri = key_prior(scene)
        ref = rh.build_reference_scene(scene, **kw)
        import io, contextlib
        buf = io.StringIO()
        ref.verbose = True
        with contextlib.redirect_stdout(buf):
            init_fun.init_minimum_spanning_tree(ref, niter_PnP=10, init_priors=pri)
        first = [l for l in buf.getvalue().splitlines() if "init edge" in l][0]
        out = dict(recipe=recipe, kw=kw, checksum=checksum(scene), edges=scene.edges, init_priors=[pri[0], None, pri[2]],
                   params=params_of(ref), first_edge=first.strip())
        path = os.path.join(ROOT, "tests", "golden", f"small_init_priors_{tag}.pt")
        torch.save(out, path)
        print(tag, "->", path, f"{os.path.getsize(path) / 1024:.0f} KB", first.strip())


if __name__ == "__main__":
    main()
