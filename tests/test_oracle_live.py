"""The CPU oracle against the REAL reference evaluated live (build container only: /root/reference; skipped on the GPU box,
where the committed fixtures of tests/golden/ stand in).  Covers what the fixtures pin on fixed scenes with fresh seeds:
the flow-free `dust3r.cloud_opt` aligner on a non-symmetrised key-frame graph, the two-image PairViewer, the
self-estimated dynamic masks, and the canonical flow-aware aligner."""
import copy
import os
import sys

import pytest
import torch

from tests.helpers import CANONICAL_KW, go, make_scene

ORACLE_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle")
sys.path.insert(0, ORACLE_DIR)
import ref_harness as rh          # noqa: E402

pytestmark = pytest.mark.skipif(not rh.reference_available(), reason="the reference tree only exists in the build container")


def test_flow_free_aligner_on_a_non_symmetrised_graph():
    """dust3r.cloud_opt (tool/depth_test.py protocol: lr 0.05, cosine, lr_min 1e-6, confidences above 1 set to 10) on three key
    frames with edges (0,1), (0,2), (1,2): frame 2 is never a first image (its focal comes from RANSAC-PnP)."""
    rh.import_reference()
    from dust3r.cloud_opt import global_aligner, GlobalAlignerMode
    import dust3r.cloud_opt.init_im_poses as init_plain
    sc = make_scene(9, 32, 48, "complete", seed=21, with_flow=False, edges=[(0, 4), (0, 8), (4, 8)])
    out = copy.deepcopy(sc.output)
    out["view1"]["idx"], out["view2"]["idx"] = [0, 0, 1], [1, 2, 2]
    for p in ("pred1", "pred2"):
        c = out[p]["conf"]
        c[c > 1] = 10
    with rh.quiet():
        ref = global_aligner(copy.deepcopy(out), False, [], device="cpu", mode=GlobalAlignerMode.PointCloudOptimizer, verbose=False,
                             min_conf_thr=3)
        init_plain.init_minimum_spanning_tree(ref, init_priors=None, niter_PnP=10)
    prob = go.problem_from_output(out, None, None, min_conf_thr=3.0)
    params, _ = go.mst_init(prob)
    for k in ("pw_poses", "im_poses", "im_focals", "im_depthmaps"):
        r = getattr(ref, k).detach()
        assert torch.allclose(params[k].reshape(r.shape), r, rtol=1e-6, atol=1e-6), k
    with rh.quiet():
        final = ref.compute_global_alignment(init=None, niter=20, schedule="cosine", lr=0.05)
    _, losses, _ = go.run_alignment(prob, params, niter=20, schedule="cosine", lr=0.05, lr_min=1e-6)
    assert abs(losses[-1] - float(final)) <= 1e-5 * abs(float(final))


def test_pair_viewer():
    global_aligner, Mode, _, _ = rh.import_reference()
    sc = make_scene(2, 40, 56, "complete", seed=33, with_flow=False)
    out = copy.deepcopy(sc.output)
    with rh.quiet():
        ref = global_aligner(copy.deepcopy(out), "cpu", mode=Mode.PairViewer, verbose=False)
    prob = go.problem_from_output(out, None, None)
    a = prob.edges.index((0, 1))
    b = prob.edges.index((1, 0))
    focals, poses, depth = go.pair_viewer(prob, a, b)
    assert torch.allclose(torch.tensor(focals), ref.get_focals().detach(), rtol=1e-6)
    assert torch.allclose(torch.stack(poses), ref.get_im_poses().detach(), atol=1e-6)
    for d, r in zip(depth, ref.get_depthmaps()):
        assert torch.allclose(d, r.detach(), rtol=1e-5, atol=1e-6)


def test_self_estimated_masks_on_a_fresh_scene():
    rh.import_reference()
    sc = make_scene(5, 40, 48, "swin-2-noncyclic", seed=41, with_flow=True)
    kw = dict(CANONICAL_KW, num_total_iter=10, use_self_mask=True)
    ref = rh.build_reference_scene(sc, **kw)
    prob = go.problem_from_output(sc.output, sc.flow_ij, sc.flow_ji, **kw)
    masks, _ = go.motion_masks_from_pairs(prob)
    assert torch.equal(masks, torch.stack(ref.dynamic_masks))


def test_canonical_aligner_on_a_fresh_scene():
    rh.import_reference()
    import dust3r.cloud_opt_flow.init_im_poses as init_fun
    sc = make_scene(5, 32, 48, "swin-2-noncyclic", seed=57, with_flow=True)
    kw = dict(CANONICAL_KW, num_total_iter=12)
    ref = rh.build_reference_scene(sc, **kw)
    with rh.quiet():
        init_fun.init_minimum_spanning_tree(ref, niter_PnP=10)
    prob = go.problem_from_output(sc.output, sc.flow_ij, sc.flow_ji, **kw)
    params, _ = go.mst_init(prob)
    for k in ("pw_poses", "im_poses", "im_focals", "im_depthmaps"):
        r = getattr(ref, k).detach()
        assert torch.allclose(params[k].reshape(r.shape), r, rtol=1e-5, atol=1e-6), k
    with rh.quiet():
        final = ref.compute_global_alignment(init=None, niter=12, schedule="linear", lr=0.01)
    _, losses, dropped = go.run_alignment(prob, params, niter=12, schedule="linear", lr=0.01)
    assert abs(losses[-1] - float(final)) <= 1e-4 * abs(float(final)) and dropped == bool(ref.flow_loss_flag)


@pytest.mark.parametrize("variant", ["l2", "adaptors_pp", "conf_sqrt", "conf_m1", "conf_id", "perframe_focal_base_scale"])
def test_variant_switches_teacher_forced(variant):
    """The constructor switches no committed fixture turns: dist='l2', trainable adaptors + principal points (with non-zero
    values), the confidence transforms, per-frame focals with another base_scale.  Oracle loss and gradients against the
    real reference's forward + autograd at the same parameters (fp32: loss 2e-6, gradients 2e-4 of the tensor's norm - the
    reference's own fp32 noise level)."""
    rh.import_reference()
    import dust3r.cloud_opt_flow.init_im_poses as init_fun
    sc = make_scene(5, 32, 48, "swin-2-noncyclic", seed=71, with_flow=True)
    kw = dict(CANONICAL_KW, num_total_iter=10, flow_loss_thre=0)
    names = list(go.TRAINABLE)
    if variant == "l2":
        kw["dist"] = "l2"
    elif variant == "adaptors_pp":
        kw.update(allow_pw_adaptors=True, optimize_pp=True)
        names += ["pw_adaptors", "im_pp"]
    elif variant.startswith("conf_"):
        kw["conf"] = variant[5:]
    else:
        kw.update(shared_focal=False, base_scale=0.8)
    ref = rh.build_reference_scene(sc, **kw)
    with rh.quiet():
        init_fun.init_minimum_spanning_tree(ref, niter_PnP=10)
    g = torch.Generator().manual_seed(2)
    with torch.no_grad():
        if variant == "adaptors_pp":
            ref.pw_adaptors.data.copy_(0.3 * torch.randn(ref.pw_adaptors.shape, generator=g))
            ref.im_pp.data.copy_(0.2 * torch.randn(ref.im_pp.shape, generator=g))
    all_names = ("pw_poses", "pw_adaptors", "im_depthmaps", "im_poses", "im_focals", "im_pp")
    params = {k: getattr(ref, k).detach().clone() for k in all_names}
    ref.zero_grad()
    with rh.quiet():
        loss_ref = ref(epoch=9999)
    loss_ref.backward()

    prob = go.problem_from_output(sc.output, sc.flow_ij, sc.flow_ji, **kw)
    mine = {k: v.reshape(go_shape(prob, k, v)) for k, v in params.items()}
    loss, grads, _ = go.loss_and_grads(prob, mine, 9999, trainable=names)
    loss, loss_ref = float(loss.detach()), float(loss_ref.detach())
    assert abs(loss - loss_ref) <= 2e-6 * abs(loss_ref), (loss, loss_ref)
    for k in names:
        r = getattr(ref, k).grad.detach()
        m = grads[k].reshape(r.shape)
        assert float((m - r).norm()) <= 2e-4 * float(r.norm()) + 1e-9, (k, float((m - r).norm()), float(r.norm()))


def go_shape(prob, name, value):
    """Shapes of the oracle's parameter dict (mst_init's): the reference stacks its parameters the same way except the focal."""
    if name == "im_focals":
        return (-1, 1)
    return tuple(value.shape)


def test_preset_frozen_focals_like_use_gt_focal():
    """tool/pose_test.py:181-193 (--use_gt_focal): per-frame focals preset and frozen (preset_focal(..., requires_grad=False),
    optimizer.py:352-363), then the canonical init='mst' run.  The MST initialisation leaves the preset focals alone
    (_set_focal only writes trainable parameters, :392-396) and the loop does not step them."""
    rh.import_reference()
    import dust3r.cloud_opt_flow.init_im_poses as init_fun
    sc = make_scene(5, 32, 48, "swin-2-noncyclic", seed=83, with_flow=True)
    kw = dict(CANONICAL_KW, num_total_iter=10, shared_focal=False)
    gt_focals = [0.9 * 48 * (1 + 0.01 * n) for n in range(5)]
    ref = rh.build_reference_scene(sc, **kw)
    with rh.quiet():
        ref.preset_focal(gt_focals, requires_grad=False)
        init_fun.init_minimum_spanning_tree(ref, niter_PnP=10)
    assert not ref.im_focals.requires_grad
    assert torch.allclose(ref.get_focals().detach().reshape(-1), torch.tensor(gt_focals), rtol=1e-6)

    prob = go.problem_from_output(sc.output, sc.flow_ij, sc.flow_ji, **kw)
    params, _ = go.mst_init(prob)
    params["im_focals"] = (prob.focal_break * torch.log(torch.tensor(gt_focals))).reshape(params["im_focals"].shape).float()
    for k in ("pw_poses", "im_poses", "im_focals", "im_depthmaps"):
        r = getattr(ref, k).detach()
        assert torch.allclose(params[k].reshape(r.shape), r, rtol=1e-5, atol=1e-6), k
    trainable = [k for k in go.TRAINABLE if k != "im_focals"]
    with rh.quiet():
        final = ref.compute_global_alignment(init=None, niter=10, schedule="linear", lr=0.01)
    end, losses, _ = go.run_alignment(prob, params, niter=10, schedule="linear", lr=0.01, trainable=trainable)
    assert abs(losses[-1] - float(final)) <= 1e-4 * abs(float(final))
    assert torch.equal(end["im_focals"], params["im_focals"])
    assert torch.allclose(ref.get_focals().detach().reshape(-1), torch.tensor(gt_focals), rtol=1e-6)


def test_flow_validity_maps():
    """oracle occ_mask against the reference's OccMask (dust3r/utils/goem_opt.py:575-618) on the flows of a synthetic scene and
    on flows that leave the image / disagree between the two directions."""
    rh.import_reference()
    from dust3r.utils.goem_opt import OccMask
    sc = make_scene(5, 32, 48, "swin-2-noncyclic", seed=91, with_flow=True)
    g = torch.Generator().manual_seed(5)
    wild_ij = sc.flow_ij + 6 * torch.randn(sc.flow_ij.shape, generator=g)
    wild_ji = sc.flow_ji + 6 * torch.randn(sc.flow_ji.shape, generator=g)
    for a, b in ((sc.flow_ij, sc.flow_ji), (sc.flow_ji, sc.flow_ij), (wild_ij, wild_ji)):
        want = OccMask(th=3.0)(a.clone(), b.clone())
        mine, _ = go.occ_mask(a, b, th=3.0)
        assert mine.shape == want.shape == (a.shape[0], 1, 32, 48) and torch.equal(mine, want.bool())
    assert 0.02 < float(go.occ_mask(wild_ij, wild_ji)[0].float().mean()) < 0.98          # both outcomes are exercised
