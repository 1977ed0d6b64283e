"""The result write-out (align3r_b200/cloud_opt_flow/writeout.py, host code) against the REAL reference's own save_* methods
(reference dust3r/cloud_opt_flow/base_opt.py:332-427), live in the build container: both write the read-outs of the SAME
reference scene object, so every file must come out byte-identical.  (save_tum_poses goes through evo in the reference,
which the image lacks: its file is checked against the reference's get_tum_poses array instead.)"""
import filecmp
import os
import sys

import numpy as np
import pytest

from tests.helpers import CANONICAL_KW, make_scene
from align3r_b200.cloud_opt_flow import writeout

ORACLE_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle")
sys.path.insert(0, ORACLE_DIR)
import ref_harness as rh          # noqa: E402

pytestmark = pytest.mark.skipif(not rh.reference_available(), reason="the reference tree only exists in the build container")


def test_written_files_equal_the_reference_methods_output(tmp_path):
    sc = make_scene(4, 24, 32, "swin-2-noncyclic", seed=12, with_flow=True, with_img=True)
    kw = dict(CANONICAL_KW, num_total_iter=4)
    ref = rh.build_reference_scene(sc, **kw)
    with rh.quiet():
        ref.compute_global_alignment(init="mst", niter=4, schedule="linear", lr=0.01)
    theirs, ours = tmp_path / "reference", tmp_path / "ours"
    theirs.mkdir()
    ours.mkdir()
    for d, call in ((theirs, lambda name, *a: getattr(ref, name)(*a)),
                    (ours, lambda name, *a: getattr(writeout, name)(ref, *a))):
        d = str(d)
        with rh.quiet():
            for clip_start in (0, 4):                       # two clips of one video: appended intrinsics and gifs
                call("save_focals", f"{d}/pred_focal.txt")
                call("save_intrinsics", f"{d}/pred_intrinsics.txt")
                call("save_conf_maps", d, clip_start)
                call("save_init_conf_maps", d, clip_start)
                call("save_rgb_imgs", d, clip_start)
                call("save_dynamic_masks", d, clip_start)
                call("save_depth_maps", d, clip_start)
            call("save_depth_maps", d, 0, [x.detach().numpy() * 0.5 for x in ref.get_depthmaps()])      # the refine variant
    names = sorted(os.listdir(str(theirs)))
    assert names == sorted(os.listdir(str(ours))) and len(names) >= 2 + 8 * (2 + 2 + 2) + 2
    match, mismatch, errors = filecmp.cmpfiles(str(theirs), str(ours), names, shallow=False)
    assert not mismatch and not errors, (mismatch, errors)

    # TUM trajectory: "t x y z qw qx qy qz", timestamps 0..N-1 as floats, full repr precision
    with rh.quiet():
        poses = writeout.save_tum_poses(ref, str(ours / "pred_traj.txt"))
    want, tt = ref.get_tum_poses()
    rows = [l.split() for l in open(str(ours / "pred_traj.txt")).read().splitlines()]
    assert [r[0] for r in rows] == [str(t) for t in tt]
    assert np.array_equal(np.array([[float(x) for x in r[1:]] for r in rows]), np.asarray(want, dtype=np.float64))
    assert np.array_equal(poses, want)
