"""TEST INFRASTRUCTURE - generates tests/golden/small_pair_viewer_{a,b}.pt by running the REAL reference (CPU).

    python oracle/make_golden_pair_viewer.py          (build container only: needs /root/reference)

GlobalAlignerMode.PairViewer (reference dust3r/cloud_opt_flow/pair_viewer.py:15-127) on two synthetic two-image pairs: "a"
with edge 0->1 the more confident (camera 0 is the world), "b" with edge 1->0 the more confident (camera 1 is).
"""
import copy
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_harness as rh                                   # noqa: E402
from align3r_b200.synth import make_scene                 # noqa: E402
from make_golden import checksum                           # noqa: E402

RECIPE = dict(n_frames=2, H=48, W=64, scene_graph="complete", seed=9, with_flow=False)


def pair_output(tag):
    scene = make_scene(**RECIPE)
    out = copy.deepcopy(scene.output)
    # lower the first-view confidences of one edge so that the other one is the more confident: "a" -> (0,1), "b" -> (1,0)
    low = (1, 0) if tag == "a" else (0, 1)
    row = [(int(i), int(j)) for i, j in zip(out["view1"]["idx"], out["view2"]["idx"])].index(low)
    out["pred1"]["conf"][row] = 1 + 0.5 * (out["pred1"]["conf"][row] - 1)
    return scene, out


def main():
    global_aligner, Mode, _, _ = rh.import_reference()
    for tag in ("a", "b"):
        scene, out = pair_output(tag)
        with rh.quiet():
            ref = global_aligner(copy.deepcopy(out), "cpu", mode=Mode.PairViewer, verbose=False)
        res = dict(recipe=RECIPE, tag=tag, checksum=checksum(scene), edges=scene.edges,
                   focals=ref.get_focals().detach().clone(), im_poses=ref.get_im_poses().detach().clone(),
                   intrinsics=ref.get_intrinsics().detach().clone(), depth=torch.stack([d.detach() for d in ref.get_depthmaps()]),
                   pts3d=torch.stack([p.detach() for p in ref.get_pts3d()]), masks=torch.stack(ref.get_masks()),
                   conf=torch.stack([c.detach() for c in ref.get_conf()]))
        path = os.path.join(ROOT, "tests", "golden", f"small_pair_viewer_{tag}.pt")
        torch.save(res, path)
        print(tag, "->", path, f"{os.path.getsize(path) / 1024:.0f} KB  focals {res['focals'].tolist()}  world = camera "
              f"{0 if float(res['im_poses'][0].sum()) == 4 else 1}")


if __name__ == "__main__":
    main()
