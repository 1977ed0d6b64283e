"""Host-side mirror of the reference API (no GPU): names, defaults, error behaviour, schedules."""
import inspect

import numpy as np
import pytest
import torch

from align3r_b200.cloud_opt_flow import GlobalAlignerMode, PointCloudOptimizer, global_aligner, global_alignment_loop
from align3r_b200.cloud_opt_flow import commons
from align3r_b200.synth import make_scene, make_edge_list, CANONICAL_KW


# defaults of reference optimizer.py:36-38 and base_opt.py:61-71
REFERENCE_DEFAULTS = dict(
    optimize_pp=False, focal_break=20, shared_focal=False, flow_loss_fn='smooth_l1', flow_loss_weight=0.0,
    depth_regularize_weight=0.0, num_total_iter=300, temporal_smoothing_weight=0, translation_weight=0.1,
    flow_loss_start_epoch=0.15, flow_loss_thre=50, sintel_ckpt=False, use_self_mask=False, pxl_thre=50,
    sam2_mask_refine=False, motion_mask_thre=0.35, dist='l1', conf='log', min_conf_thr=3, thr_for_init_conf=False,
    base_scale=0.5, allow_pw_adaptors=False, pw_break=20, empty_cache=False, verbose=True)


def test_constructor_keywords_and_defaults_mirror_the_reference():
    sig = inspect.signature(PointCloudOptimizer.__init__)
    for k, v in REFERENCE_DEFAULTS.items():
        assert k in sig.parameters, k
        assert sig.parameters[k].default == v, (k, sig.parameters[k].default, v)
    assert sig.parameters['rand_pose'].default is torch.randn


def test_loop_signature_mirrors_the_reference():
    sig = inspect.signature(global_alignment_loop)
    want = dict(lr=0.01, niter=300, schedule='cosine', lr_min=1e-3, temporal_smoothing_weight=0, depth_map_save_dir=None)
    assert {k: p.default for k, p in list(sig.parameters.items())[1:]} == want
    sig = inspect.signature(PointCloudOptimizer.compute_global_alignment)
    for k, v in dict(init=None, init_priors=None, save_score_path=None, save_score_only=False, niter_PnP=10).items():
        assert sig.parameters[k].default == v


def test_mode_enum_members():
    assert [m.name for m in GlobalAlignerMode] == ["PointCloudOptimizer", "ModularPointCloudOptimizer", "PairViewer"]


def test_cpu_device_is_refused_loudly():
    scene = make_scene(3, 8, 8, "swin-2-noncyclic", with_flow=False)
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        global_aligner(scene.output, "cpu", mode=GlobalAlignerMode.PointCloudOptimizer, verbose=False)
    with pytest.raises(NotImplementedError):
        global_aligner(scene.output, "cuda", mode=GlobalAlignerMode.ModularPointCloudOptimizer)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure path")
def test_without_a_gpu_nothing_falls_back():
    scene = make_scene(3, 8, 8, "swin-2-noncyclic", with_flow=False)
    with pytest.raises(RuntimeError, match="no CUDA device"):
        global_aligner(scene.output, "cuda", mode=GlobalAlignerMode.PointCloudOptimizer, verbose=False)


def test_schedules_match_reference_formulas():
    assert commons.schedule_lr('linear', 0, 300, 0.01, 1e-3) == 0.01
    assert abs(commons.schedule_lr('linear', 150, 300, 0.01, 1e-3) - 0.0055) < 1e-15
    assert abs(commons.schedule_lr('cosine', 300, 300, 0.01, 1e-3) - 1e-3) < 1e-12
    assert abs(commons.schedule_lr('cycle3', 50, 300, 0.01, 1e-3) - commons.linear_schedule(0.5, 0.01, 1e-3)) < 1e-15
    with pytest.raises(ValueError):
        commons.schedule_lr('bogus', 0, 10, 0.01, 1e-3)
    with pytest.raises(AssertionError):
        commons.linear_schedule(1.5, 0.01, 1e-3)


def test_signed_maps_and_quaternions():
    x = torch.tensor([-2.0, -0.0, 0.0, 0.5])
    assert torch.allclose(commons.signed_expm1(commons.signed_log1p(x)), x)
    q = torch.tensor([[0.1, -0.2, 0.3, 0.9]])
    q = q / q.norm()
    R = commons.unitquat_to_rotmat(q)[0]
    assert torch.allclose(R @ R.T, torch.eye(3), atol=1e-6) and abs(float(torch.det(R)) - 1) < 1e-6
    q2 = commons.rotmat_to_unitquat(R)
    assert min(float((q2 - q[0]).norm()), float((q2 + q[0]).norm())) < 1e-6     # xyzw, sign-ambiguous


def test_scene_graph_edge_lists():
    e = make_edge_list(6, "swin-2-noncyclic")
    half = len(e) // 2
    assert e[half:] == [(j, i) for i, j in e[:half]]                   # symmetrised mirror halves
    assert len(make_edge_list(6, "swin-2")) == 2 * 2 * 6               # cyclic: 2*w*N
    assert len(make_edge_list(5, "complete")) == 20
    assert len(make_edge_list(7, "oneref-3")) == 12
    assert all(abs(i - j) in (1, 2, 4) for i, j in make_edge_list(9, "logwin-3-noncyclic"))
    assert all(abs(i - j) in (1, 3, 5) for i, j in make_edge_list(12, "swinstride-3-noncyclic"))


def test_synthetic_scene_is_reproducible_and_sliceable():
    a = make_scene(5, 16, 24, "swin-2-noncyclic", seed=4)
    sub = [a.edges[3], a.edges[0]]
    b = make_scene(5, 16, 24, "swin-2-noncyclic", seed=4, edges=sub)
    assert torch.equal(b.pred1["pts3d"][0], a.pred1["pts3d"][3]) and torch.equal(b.flow_ji[1], a.flow_ji[0])
    assert set(CANONICAL_KW) >= {"shared_focal", "flow_loss_weight", "temporal_smoothing_weight"}


def test_edge_score_map_writer(tmp_path):
    """save_score_path / save_score_only (reference init_im_poses.py:24-42,94-97): the matrix behind the reference's
    heat-map, NaN where the graph has no edge; the picture itself only where matplotlib + seaborn exist."""
    from align3r_b200.cloud_opt_flow.init_im_poses import edge_score_matrix, save_edge_scores_map
    scores = {(0, 1): 2.5, (1, 0): 2.25, (1, 2): 3.0, (3, 1): 1.5}
    m = edge_score_matrix(scores)
    assert m.shape == (4, 4) and np.isnan(m).sum() == 12
    assert m[0, 1] == 2.5 and m[1, 0] == 2.25 and m[1, 2] == 3.0 and m[3, 1] == 1.5
    assert edge_score_matrix(scores, n_imgs=6).shape == (6, 6)
    path = str(tmp_path / "scores.png")
    out = save_edge_scores_map(scores, path, n_imgs=5)
    back = np.load(str(tmp_path / "scores.npy"))
    assert back.shape == (5, 5) and np.array_equal(back, out, equal_nan=True)


def test_tree_file_lists_the_oracle_spanning_tree(tmp_path, monkeypatch):
    """The graph pass of the MST initialisation is host logic: run it on a stand-in object (confidence sums of a fixture
    scene on the CPU, stopped before the first device launch) and compare the tree file it writes
    (reference init_im_poses.py:185-190,229-250) with the oracle's spanning tree."""
    import types
    from tests.helpers import go, golden_case
    from align3r_b200.cloud_opt_flow import init_im_poses as init_fun
    fx, scene, prob = golden_case("small_flow")
    E, N = prob.n_edges, prob.n_imgs
    conf_sum = torch.zeros(2 * E, dtype=torch.float64)
    conf_sum[0::2] = prob.conf_i.reshape(E, -1).double().sum(1)
    conf_sum[1::2] = prob.conf_j.reshape(E, -1).double().sum(1)
    stand_in = types.SimpleNamespace(_shard=None, _conf_sum=conf_sum, _lslot={g: g for g in range(2 * E)},
                                     imshapes=[(prob.H, prob.W)] * N, edges=list(prob.edges), n_imgs=N, n_edges=E,
                                     verbose=False, device='cpu')

    class DeviceWorkStarts(Exception):
        pass

    def stop(*a, **k):
        raise DeviceWorkStarts

    monkeypatch.setattr(init_fun, "_run_procrustes", stop)
    path = str(tmp_path / "scores.png")
    with pytest.raises(DeviceWorkStarts):
        init_fun.init_minimum_spanning_tree(stand_in, save_score_path=path)
    m = np.load(str(tmp_path / "scores.npy"))
    for (i, j), s in go.edge_scores(prob).items():
        assert abs(m[i, j] - s) <= 1e-6 * abs(s)
    lines = open(str(tmp_path / "scores_tree.txt")).read().splitlines()
    tree = go.mst_init(prob)[1]["tree"]
    assert len(lines) == len(tree) == N - 1
    assert lines[0].startswith(f"init edge ({tree[0][0]}*,{tree[0][1]}*) score=")
    for line, (i, j) in zip(lines[1:], tree[1:]):
        assert line.startswith(f"init edge ({i}*,{j}) score=") or line.startswith(f"init edge ({i},{j}*) score="), line
    # scores only: nothing but the matrix is produced
    only = str(tmp_path / "only.png")
    assert init_fun.init_minimum_spanning_tree(stand_in, save_score_path=only, save_score_only=True) is None
    assert np.array_equal(np.load(str(tmp_path / "only.npy"))[~np.isnan(m)], m[~np.isnan(m)])


def test_progress_pictures_of_the_loop(tmp_path):
    """depth_map_save_dir (reference base_opt.py:531-540): one jet-coloured picture per frame, named like the reference's."""
    import cv2
    from align3r_b200.cloud_opt_flow.writeout import save_depthmap_pictures
    maps = [torch.linspace(1, 3, 20 * 30).reshape(20, 30), torch.full((20, 30), 2.0)]
    d = str(tmp_path / "progress")
    files = save_depthmap_pictures(maps, d, 500)
    assert [f.split("/")[-1] for f in files] == ["depthmaps_0_iter_500.png", "depthmaps_1_iter_500.png"]
    a, b = cv2.imread(files[0]), cv2.imread(files[1])
    assert a.shape[:2] == (20, 30) and b.shape[:2] == (20, 30)
    assert int(a[0, 0, 0]) > int(a[0, 0, 2]) and int(a[-1, -1, 2]) > int(a[-1, -1, 0])      # near = blue, far = red (BGR)
    assert len(np.unique(b.reshape(-1, 3), axis=0)) == 1                                      # a constant map is one colour
