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
