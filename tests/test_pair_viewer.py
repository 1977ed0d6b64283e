"""GlobalAlignerMode.PairViewer (align3r_b200/cloud_opt_flow/pair_viewer.py) against fixtures of the real reference's
PairViewer (oracle/make_golden_pair_viewer.py): both orientations of the "more confident edge" rule."""
import copy
import os
import sys

import pytest
import torch

from tests.helpers import load_golden, make_scene, scene_checksum, rel_err

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))


def fixture_output(tag):
    from make_golden_pair_viewer import pair_output
    fx = load_golden(f"small_pair_viewer_{tag}")
    scene, out = pair_output(tag)
    assert abs(scene_checksum(scene) - fx["checksum"]) <= 1e-9 * abs(fx["checksum"])
    return fx, out


def test_fixtures_cover_both_orientations():
    a, b = load_golden("small_pair_viewer_a"), load_golden("small_pair_viewer_b")
    assert float(a["im_poses"][0].sum()) == 4.0 and float(b["im_poses"][1].sum()) == 4.0        # identity = the world camera
    assert float(a["im_poses"][1].sum()) != 4.0 and float(b["im_poses"][0].sum()) != 4.0


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["a", "b"])
def test_pair_viewer_matches_reference(cuda_device, tag):
    from align3r_b200.cloud_opt_flow import GlobalAlignerMode, global_aligner
    fx, out = fixture_output(tag)
    net = global_aligner(copy.deepcopy(out), "cuda", mode=GlobalAlignerMode.PairViewer, verbose=False)
    assert rel_err(net.get_focals().cpu(), fx["focals"]) < 1e-5 and net.get_focals().shape == (2,)
    assert float((net.get_im_poses().cpu() - fx["im_poses"]).abs().max()) < 1e-4
    assert float((net.get_intrinsics().cpu() - fx["intrinsics"]).abs().max()) < 1e-3
    d = torch.stack(net.get_depthmaps()).cpu()
    assert float((d - fx["depth"]).abs().max()) < 1e-4 * float(fx["depth"].abs().max())
    p = torch.stack(net.get_pts3d()).cpu()
    assert float((p - fx["pts3d"]).abs().max()) < 1e-3
    assert torch.equal(torch.stack(net.get_masks()).cpu(), fx["masks"])
    assert float((torch.stack(net.get_conf()).cpu() - fx["conf"]).abs().max()) < 1e-6
    with pytest.raises(NotImplementedError):
        net.compute_global_alignment(init="mst", niter=10)
