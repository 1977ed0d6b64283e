"""Host code of the product against the REAL reference evaluated live (build container only; skipped on the GPU box).

* The result write-out (align3r_b200/cloud_opt_flow/writeout.py) against the reference's own save_* methods
  (reference dust3r/cloud_opt_flow/base_opt.py:332-427): both write the read-outs of the SAME reference scene object, so
  every file must come out byte-identical.  (save_tum_poses goes through evo in the reference, which the image lacks: its
  file is checked against the reference's get_tum_poses array instead.)
* get_tum_poses / align_poses, the learning-rate schedules, the confidence transforms and the signed log maps against the
  reference's functions of the same names."""
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


def test_tum_read_out_equals_the_reference(tmp_path):
    """get_tum_poses / align_poses (reference base_opt.py:29-42,305-330) of the product class, run on the reference scene
    object, against the reference's own methods - with and without the key pose of the hierarchical clip driver."""
    from align3r_b200.cloud_opt_flow import PointCloudOptimizer as Ours
    sc = make_scene(4, 24, 32, "swin-2-noncyclic", seed=13, with_flow=False)
    ref = rh.build_reference_scene(sc, **dict(CANONICAL_KW, flow_loss_weight=0.0, num_total_iter=3))
    with rh.quiet():
        ref.compute_global_alignment(init="mst", niter=3, schedule="linear", lr=0.01)
    key = np.eye(4)
    key[:3, :3] = np.array([[0.0, -1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]])
    key[:3, 3] = (0.3, -0.2, 1.5)
    poses = ref.get_im_poses().detach().numpy()
    assert np.array_equal(Ours.align_poses(ref, key, poses), ref.align_poses(key, poses))
    for kp in (None, key):
        mine, want = Ours.get_tum_poses(ref, kp), ref.get_tum_poses(kp)
        assert np.array_equal(mine[0], want[0]) and np.array_equal(mine[1], want[1])


def test_schedules_and_small_maps_equal_the_reference():
    """commons.schedule_lr against the learning rate the reference's loop sets (base_opt.py:550-563 -> commons.py:88-103) on
    every iteration of short, odd and canonical runs; confidence transforms (commons.py:53-64); signed maps (:78-85)."""
    import torch
    rh.import_reference()
    import dust3r.cloud_opt_flow.commons as ref_c
    from align3r_b200.cloud_opt_flow import commons as our_c

    def ref_lr(schedule, it, niter, lr, lr_min):
        t = it / niter
        if schedule == 'cosine':
            return ref_c.cosine_schedule(t, lr, lr_min)
        if schedule == 'linear':
            return ref_c.linear_schedule(t, lr, lr_min)
        cycles = int(schedule[5:]) if schedule[5:].isdigit() else 2
        return ref_c.cycled_linear_schedule(t, lr, lr_min, num_cycles=cycles)

    for niter in (1, 7, 300):
        for schedule in ('cosine', 'linear', 'cycle', 'cycle3'):
            for lr, lr_min in ((0.01, 1e-3), (0.05, 1e-6)):
                for it in range(niter):
                    assert our_c.schedule_lr(schedule, it, niter, lr, lr_min) == ref_lr(schedule, it, niter, lr, lr_min)
    x = 1 + 3 * torch.rand(257, generator=torch.Generator().manual_seed(0))
    for mode in ('log', 'sqrt', 'm1', 'id', 'none'):
        w = our_c.get_conf_trf(mode)(x)
        assert torch.equal(w, ref_c.get_conf_trf(mode)(x))
        assert torch.allclose(our_c.get_conf_untrf(mode)(w), x, rtol=1e-6)
    with pytest.raises(ValueError):
        our_c.get_conf_trf('bogus')
    y = torch.tensor([-3.5, -1e-9, -0.0, 0.0, 1e-9, 0.25, 40.0])
    assert torch.equal(our_c.signed_log1p(y), ref_c.signed_log1p(y)) and torch.equal(our_c.signed_expm1(y), ref_c.signed_expm1(y))
    assert our_c.edge_str(3, 14) == ref_c.edge_str(3, 14) and our_c.i_j_ij((3, 14)) == ref_c.i_j_ij((3, 14))


def test_clip_pairs_equal_the_reference_on_a_sweep():
    """hierarchical.my_make_pairs / choose_clip_size against the reference's own my_make_pairs (tool/depth_test.py, taken from its
    source with ast: the tool loads networks at import) and clip-size rule (tool/depth_test.py:626-628) over a sweep of video
    lengths and clip sizes."""
    import ast
    import copy
    from align3r_b200 import hierarchical as hi
    path = os.path.join(rh.REFERENCE_ROOT, "tool", "depth_test.py")
    tree = ast.parse(open(path).read())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "my_make_pairs"][0]
    ns = {}
    exec(compile(ast.Module([fn], []), path, "exec"), ns)
    ref_pairs = ns["my_make_pairs"]

    def ident(pairs):
        return [(a["idx"], b["idx"], a["frame"], b["frame"]) for a, b in pairs]

    checked = 0
    for n in list(range(4, 34)) + [50, 90, 110]:
        for clip in (3, 4, 5, 8, 10, 20):
            imgs = [dict(idx=k, frame=k) for k in range(n)]
            want = clip
            while want > 1 and (n % want == 1 or n % want == 0 or want > n):
                want -= 1
            if want <= 1:
                continue
            assert hi.choose_clip_size(n, clip) == want, (n, clip)
            r_key, r_kid, r_clips, r_cid = ref_pairs(copy.deepcopy(imgs), want)
            o_key, o_kid, o_clips, o_cid = hi.my_make_pairs(imgs, want)
            assert list(o_kid) == list(r_kid) and [list(c) for c in o_cid] == [list(c) for c in r_cid], (n, clip)
            assert ident(o_key) == ident(r_key), (n, clip)
            assert [ident(c) for c in o_clips] == [ident(c) for c in r_clips], (n, clip)
            checked += 1
    assert checked > 100
