"""Parity of the CUDA path (through the drop-in Python class and the C ABI) against the reference fixtures
and the CPU oracle.  Run on the B200 box:  python -m pytest tests -m gpu"""
import ctypes as C
import math
import os

import numpy as np
import pytest
import torch

from tests.helpers import CANONICAL_KW, go, golden_case, grad_close, load_golden, make_scene, rel_err, scene_checksum
from align3r_b200 import _lib
from align3r_b200.cloud_opt_flow import GlobalAlignerMode, global_aligner
from align3r_b200.cloud_opt_flow import init_im_poses as init_fun
from align3r_b200.problem import adam_schedule

pytestmark = pytest.mark.gpu
CASES = ["small_flow", "small_perframe_focal", "small_gate_trips"]


def build(scene, kw, **extra):
    kw = dict(kw)
    kw.update(extra)
    flow = (scene.flow_ij, scene.flow_ji) if kw.get("flow_loss_weight", 0) > 0 else None
    return global_aligner(scene.output, "cuda", mode=GlobalAlignerMode.PointCloudOptimizer, flow=flow, **kw)


def set_params(net, params):
    P = params["im_depthmaps"].shape[1]
    with torch.no_grad():
        for k, v in params.items():
            if k == "im_depthmaps":
                net.im_depthmaps.data[:, :P] = v[net._plan.lframe_id.tolist()].to(net.device)
            else:
                getattr(net, k).data.copy_(v.to(net.device).view_as(getattr(net, k)))


def grads_of(net, P):
    out = {}
    for k in ("pw_poses", "im_depthmaps", "im_poses", "im_focals"):
        g = getattr(net, k).grad
        out[k] = (g[:, :P] if k == "im_depthmaps" else g).detach().cpu()
    return out


def test_native_library_is_the_one_running(cuda_device):
    lib = _lib.load()
    sm, major, minor = C.c_int(), C.c_int(), C.c_int()
    assert lib.a3r_device_info(0, C.byref(sm), C.byref(major), C.byref(minor)) == 0
    assert major.value >= 10, f"expected a Blackwell GPU, got sm_{major.value}{minor.value}"
    assert "liba3r_ga.so" in open("/proc/self/maps").read()


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("tag", ["pre_flow", "flow"])
def test_teacher_forced_loss_and_gradients(cuda_device, name, tag):
    fx, scene, prob = golden_case(name)
    g32, g64 = fx[f"tf32_{tag}"], fx[f"tf64_{tag}"]
    net = build(scene, fx["kw"])
    set_params(net, fx["init_params"])
    loss = net(epoch=g32["epoch"])
    loss.backward()
    assert abs(float(loss.detach()) - float(g64["loss"])) <= 2e-6 * abs(float(g64["loss"]))
    assert net.flow_loss_flag == g32["flow_dropped"]
    mine = grads_of(net, prob.P)
    for k in mine:
        ok, err, ref_err = grad_close(k, mine[k], g32["grads"][k], g64["grads"][k])
        assert ok, (k, err, ref_err)
    # frozen parameters get no gradient
    assert net.pw_adaptors.grad is None and net.im_pp.grad is None


@pytest.mark.parametrize("name", CASES)
def test_adam_step(cuda_device, name):
    fx, scene, prob = golden_case(name)
    st = fx["adam_step"]
    net = build(scene, fx["kw"])
    set_params(net, fx["init_params"])
    b = net._bufs
    b.reset_adam()
    mode = _lib.M_FLOW_ACTIVE | _lib.M_APPLY_ADAM
    net._phase_accumulate(mode)
    net._phase_update(mode, adam_schedule([st["lr"]], start_step=0, first_iter=0))
    assert abs(float(b.loss_hist[0]) - st["loss"]) <= 2e-6 * abs(st["loss"])
    lr = st["lr"]
    for k in go.TRAINABLE:
        mine = getattr(net, k).detach().cpu()
        mine = mine[:, :prob.P] if k == "im_depthmaps" else mine.reshape(st["params"][k].shape)
        d = (mine - st["params"][k]).abs()
        assert float(d.max()) <= 2.0 * lr + 1e-6, k
        assert float((d > 1e-5).float().mean()) < 0.02, (k, float((d > 1e-5).float().mean()))


@pytest.mark.parametrize("name", CASES)
def test_mst_initialisation(cuda_device, name):
    fx, scene, prob = golden_case(name)
    net = build(scene, fx["kw"])
    init_fun.init_minimum_spanning_tree(net, niter_PnP=10)
    ref = fx["init_params"]
    # pairwise poses: compare the transforms they encode (quaternion sign is free)
    T_ref = go.pw_pose_matrices(prob, ref["pw_poses"])
    T = net.get_pw_poses().detach().cpu()
    assert float((T - T_ref).abs().max()) < 5e-5
    # image poses, focal, depth: frames initialised by RANSAC-PnP inherit its sensitivity to fp32 rounding
    P_ref = go.poses_to_matrices(ref["im_poses"])
    Pm = net.get_im_poses().detach().cpu()
    assert float((Pm - P_ref).abs().max()) < 2e-3
    f_ref = go.focals(prob, ref["im_focals"])
    assert rel_err(net.get_focals().detach().cpu(), f_ref) < 1e-5
    d = (net.im_depthmaps.detach().cpu()[:, :prob.P] - ref["im_depthmaps"]).abs()
    assert float(d.median()) < 1e-5 and float(d.max()) < 2e-3
    # and the loss right after the init is the reference's
    loss0 = float(net(epoch=0))
    assert abs(loss0 - float(fx["tf32_pre_flow"]["loss"])) <= 5e-4 * abs(float(fx["tf32_pre_flow"]["loss"]))


@pytest.mark.parametrize("name", CASES)
def test_free_running_loop_from_reference_init(cuda_device, name):
    """From the reference's own MST initialisation the loop (niter iterations of K0-K3) ends where the reference
    ends: final loss 2e-5 relative, median |log-depth| difference 2e-6, camera matrices 2e-5."""
    fx, scene, prob = golden_case(name)
    net = build(scene, fx["kw"])
    set_params(net, fx["init_params"])
    loss = net.compute_global_alignment(init=None, niter=fx["niter"], schedule='linear', lr=0.01)
    ref = fx["free_run"]
    assert net.flow_loss_flag == ref["flow_dropped"]
    assert abs(loss - ref["loss"]) <= 2e-5 * abs(ref["loss"])
    d = (net.im_depthmaps.detach().cpu()[:, :prob.P] - ref["params"]["im_depthmaps"]).abs()
    assert float(d.median()) < 2e-6
    assert float((net.get_im_poses().detach().cpu() - ref["im_poses_4x4"]).abs().max()) < 2e-5
    assert rel_err(net.get_focals().detach().cpu(), ref["focals"]) < 2e-5


@pytest.mark.parametrize("name", CASES)
def test_free_running_alignment(cuda_device, name):
    """The canonical call (init='mst', linear schedule).  Our MST initialisation differs from the reference's by
    fp32 rounding (~5e-7, different summation order); Adam's sign-like first steps amplify that on parameters
    whose gradient starts near zero (hard part 3), so the end state is compared with the
    stated end-state tolerances: final loss 2e-3 relative, median |log-depth| difference 2e-3, cameras 1e-2."""
    fx, scene, prob = golden_case(name)
    net = build(scene, fx["kw"])
    loss = net.compute_global_alignment(init='mst', niter=fx["niter"], schedule='linear', lr=0.01)
    ref = fx["free_run"]
    assert net.flow_loss_flag == ref["flow_dropped"]
    assert abs(loss - ref["loss"]) <= 2e-3 * abs(ref["loss"])
    d = (net.im_depthmaps.detach().cpu()[:, :prob.P] - ref["params"]["im_depthmaps"]).abs()
    assert float(d.median()) < 2e-3
    assert float((net.get_im_poses().detach().cpu() - ref["im_poses_4x4"]).abs().max()) < 1e-2
    assert rel_err(net.get_focals().detach().cpu(), ref["focals"]) < 5e-3
    # getters keep the reference's shapes
    assert len(net.get_depthmaps()) == prob.n_imgs and net.get_depthmaps()[0].shape == (prob.H, prob.W)
    assert net.get_pts3d()[0].shape == (prob.H, prob.W, 3) and net.get_intrinsics().shape == (prob.n_imgs, 3, 3)
    tum, tt = net.get_tum_poses()
    assert tum.shape == (prob.n_imgs, 7) and len(tt) == prob.n_imgs
    # read-out metrics of the north star: Sintel-protocol depth metrics and ATE / RPE of our end state equal those of
    # the reference's end state (both against the synthetic ground truth).  Stated tolerances: 2e-3 absolute on
    # Abs Rel / delta<1.25 / log RMSE, 5e-3 on ATE (scene units), 0.1 degree on RPE rotation.
    n = prob.n_imgs
    gt_depth = np.stack([scene.gt.depth(k).numpy() for k in range(n)])
    ours_d = torch.stack(net.get_depthmaps()).detach().cpu().numpy()
    ref_d = ref["params"]["im_depthmaps"].exp().reshape(n, prob.H, prob.W).numpy()
    dm, dr = go.depth_metrics(ours_d, gt_depth), go.depth_metrics(ref_d, gt_depth)
    for k in ("abs_rel", "log_rmse", "d1"):
        assert abs(dm[k] - dr[k]) <= 2e-3, (k, dm[k], dr[k])
    pm = go.pose_metrics(net.get_im_poses().detach().cpu().numpy(), scene.gt.cam2world)
    pr = go.pose_metrics(ref["im_poses_4x4"].numpy(), scene.gt.cam2world)
    assert abs(pm["ate"] - pr["ate"]) <= 5e-3 and abs(pm["rpe_rot"] - pr["rpe_rot"]) <= 0.1, (pm, pr)


def test_pointmaps_and_getters_against_oracle(cuda_device):
    fx, scene, prob = golden_case("small_flow")
    net = build(scene, fx["kw"])
    set_params(net, fx["init_params"])
    world = go.unproject(prob, fx["init_params"])
    mine = torch.stack([p.reshape(-1, 3) for p in net.get_pts3d()]).cpu()
    assert float((mine - world).abs().max()) < 2e-5
    K = go.intrinsics(prob, fx["init_params"]["im_focals"], fx["init_params"]["im_pp"])
    assert torch.allclose(net.get_intrinsics().cpu(), K, rtol=1e-6, atol=1e-4)
    assert torch.allclose(net.get_pw_scale().detach().cpu(), go.pw_scale(prob, fx["init_params"]["pw_poses"]), rtol=1e-5)
    assert torch.allclose(torch.stack(net.get_depthmaps()).detach().cpu().reshape(prob.n_imgs, -1),
                          fx["init_params"]["im_depthmaps"].exp(), rtol=1e-6)
    # im_conf = max over incident edges of the raw confidences (reference base_opt.py:187-193)
    conf = torch.zeros(prob.n_imgs, prob.P)
    for e, (i, j) in enumerate(prob.edges):
        conf[i] = torch.maximum(conf[i], prob.conf_i[e])
        conf[j] = torch.maximum(conf[j], prob.conf_j[e])
    assert torch.equal(torch.stack([c.reshape(-1) for c in net.im_conf]).cpu(), conf)
    assert torch.equal(torch.stack(net.get_masks()).cpu().reshape(prob.n_imgs, -1), conf > 3)
    assert torch.allclose(net.pred_i[net.str_edges[2]].cpu(), scene.pred1["pts3d"][2])
    assert torch.allclose(net.conf_j[net.str_edges[1]].cpu(), scene.pred2["conf"][1], rtol=1e-6)


@pytest.mark.parametrize("variant", ["l2", "adaptors_pp", "no_flow_masks_absent"])
def test_medium_scene_against_live_oracle(cuda_device, variant):
    """96x128, 8 frames, swin-3: several tiles per frame, every frame class of the sliding window."""
    scene = make_scene(8, 96, 128, "swin-3-noncyclic", seed=21, with_flow=True)
    kw = dict(CANONICAL_KW, num_total_iter=10, flow_loss_thre=0)
    extra = {}
    if variant == "l2":
        kw["dist"] = "l2"
    if variant == "adaptors_pp":
        extra = dict(allow_pw_adaptors=True, optimize_pp=True)
    if variant == "no_flow_masks_absent":
        kw["flow_loss_weight"] = 0.0
    prob = go.problem_from_output(scene.output, scene.flow_ij, scene.flow_ji, **kw)
    params, _ = go.mst_init(prob)
    g = torch.Generator().manual_seed(1)
    names = list(go.TRAINABLE)
    if variant == "adaptors_pp":
        params["pw_adaptors"] = 0.3 * torch.randn(prob.n_edges, 2, generator=g)
        params["im_pp"] = 0.2 * torch.randn(prob.n_imgs, 2, generator=g)
        names += ["pw_adaptors", "im_pp"]
    p64 = {k: v.double() for k, v in params.items()}
    loss64, g64, _ = go.loss_and_grads(prob.to(torch.float64), p64, 9999, trainable=names)
    _, g32, _ = go.loss_and_grads(prob, params, 9999, trainable=names)
    net = build(scene, kw, **extra)
    set_params(net, params)
    loss = net(epoch=9999)
    loss.backward()
    assert abs(float(loss.detach()) - float(loss64)) <= 2e-6 * abs(float(loss64))
    for k in names:
        m = getattr(net, k).grad.detach().cpu()
        m = m[:, :prob.P] if k == "im_depthmaps" else m.reshape(g64[k].shape)
        ok, err, ref_err = grad_close(k, m, g32[k], g64[k])
        assert ok, (k, err, ref_err)


def test_split_work_frames_complete_graph(cuda_device):
    """A complete graph over 18 frames puts 34 slots on every frame (> a3r_max_slots() = 32): frames are split
    into work frames whose per-pixel gradients are added atomically."""
    scene = make_scene(18, 16, 24, "complete", seed=4, with_flow=False)
    kw = dict(CANONICAL_KW, flow_loss_weight=0.0)
    prob = go.problem_from_output(scene.output, None, None, **kw)
    g = torch.Generator().manual_seed(0)
    E, N = prob.n_edges, prob.n_imgs
    params = dict(pw_poses=0.3 * torch.randn(E, 8, generator=g), pw_adaptors=torch.zeros(E, 2),
                  im_depthmaps=0.7 + 0.1 * torch.randn(N, prob.P, generator=g), im_poses=0.3 * torch.randn(N, 7, generator=g),
                  im_focals=torch.tensor([[20 * math.log(20.0)]]), im_pp=torch.zeros(N, 2))
    params["pw_poses"][:, 3] += 2
    params["im_poses"][:, 3] += 2
    net = build(scene, kw)
    assert net._plan.n_work_frames == 2 * N
    set_params(net, params)
    loss = net(epoch=0)
    loss.backward()
    p64 = {k: v.double() for k, v in params.items()}
    loss64, g64, _ = go.loss_and_grads(prob.to(torch.float64), p64, 0)
    assert abs(float(loss.detach()) - float(loss64)) <= 2e-6 * abs(float(loss64))
    mine = grads_of(net, prob.P)
    for k in mine:
        assert rel_err(mine[k].reshape(g64[k].shape), g64[k]) < 5e-4, k


def test_split_work_frames_with_flow_on_row_tiles(cuda_device):
    """The same with the flow term on and 128-pixel-wide frames: a complete graph over 12 frames puts 22 slots on every frame
    (> a3r_max_slots()), so the row-tile scan runs on split work frames (atomic per-pixel gradient planes, the stand-alone
    depth Adam kernel instead of the folded step).  Loss and gradients against the fp64 oracle, then 8 iterations against the
    oracle's own 8."""
    scene = make_scene(12, 8, 128, "complete", seed=4, with_flow=True)
    kw = dict(CANONICAL_KW, num_total_iter=8, flow_loss_thre=0)
    prob = go.problem_from_output(scene.output, scene.flow_ij, scene.flow_ji, **kw)
    params, _ = go.mst_init(prob)
    net = build(scene, kw)
    net.verbose = False
    assert net._plan.n_work_frames == 2 * prob.n_imgs and net._bufs.flags(net._trainable()) & _lib.F_ROW_TILES
    set_params(net, params)
    loss = net(epoch=9999)
    loss.backward()
    p64 = {k: v.double() for k, v in params.items()}
    loss64, g64, _ = go.loss_and_grads(prob.to(torch.float64), p64, 9999)
    _, g32, _ = go.loss_and_grads(prob, params, 9999)
    assert abs(float(loss.detach()) - float(loss64)) <= 2e-6 * abs(float(loss64))
    mine = grads_of(net, prob.P)
    for k in mine:
        ok, err, ref = grad_close(k, mine[k].reshape(g64[k].shape), g32[k], g64[k])
        assert ok, (k, err, ref)
    run = net.compute_global_alignment(init=None, niter=8, schedule="linear", lr=0.01)
    _, losses, _ = go.run_alignment(prob, params, niter=8, schedule="linear", lr=0.01)
    assert abs(run - losses[-1]) <= 1e-4 * abs(losses[-1]), (run, losses[-1])


@pytest.mark.parametrize("seed", range(6))
def test_device_ransac_pnp_equals_opencv(cuda_device, seed):
    """cloud_opt_flow/pnp_ransac.py: the RANSAC-PnP whose per-hypothesis inlier test runs on the GPU (a3r_pnp_inliers) returns
    what cv2.solvePnPRansac returns on the same points - the same number of inliers, rvec and tvec bit for bit - at the sizes
    and iteration limits of its three call sites (init_im_poses.py:498: 10 iterations; pair_viewer.py:52: 100)."""
    import cv2
    from align3r_b200.cloud_opt_flow import pnp_ransac as pr
    from tests.test_pnp_ransac import scene as pnp_scene
    H, W = ((288, 512), (96, 160), (48, 64))[seed % 3]
    X, msk, K = pnp_scene(H, W, 0.01 * (1 + seed % 3), 0.06 * (seed % 5), 100 + seed)
    iterations = 10 if seed % 2 == 0 else 100
    pix = np.mgrid[:W, :H].T.astype(np.float32)
    ok, rvec, tvec, inl = cv2.solvePnPRansac(X[msk], pix[msk], K, None, iterationsCount=iterations, reprojectionError=5,
                                             flags=cv2.SOLVEPNP_SQPNP)
    ok2, rvec2, tvec2, n2 = pr.solve_pnp_ransac_device(X, msk, K, "cuda", iterations=iterations, reproj=5.0)
    assert bool(ok) == bool(ok2)
    if ok:
        assert n2 == len(inl)
        assert np.array_equal(rvec, rvec2) and np.array_equal(tvec, tvec2)
    # the call sites' entry runs the device path (its once-per-process self-check has passed)
    ok3, rvec3, tvec3, n3 = pr.solve(X, msk, K, "cuda", iterations=iterations, reproj=5)
    assert pr.enabled() and bool(ok3) == bool(ok) and (not ok or (n3 == len(inl) and np.array_equal(rvec, rvec3)))
    # the MST initialisation's form: a raw slot plane on the device + the similarity that makes its points world points
    # (a3r_similarity_points with numpy's rounding, sampled rows through a3r_pnp_gather: no host-to-device copy at all)
    r = np.random.default_rng(seed)
    A = (cv2.Rodrigues(r.normal(0, 0.4, 3))[0] * (0.5 + r.random())).astype(np.float32)
    t32 = r.normal(0, 1, 3).astype(np.float32)
    Xn = X.reshape(-1, 3)
    world = (Xn[:, 0:1] * A[:, 0] + Xn[:, 1:2] * A[:, 1] + Xn[:, 2:3] * A[:, 2] + t32)
    plane = torch.zeros(H * W, 4, device="cuda")
    plane[:, :3] = torch.from_numpy(Xn).cuda()
    frame = pr.PlaneFrame(plane, A, t32, torch.from_numpy(msk).cuda(), H, W)
    assert np.array_equal(frame.d_pts.cpu().numpy(), world)                      # bit for bit numpy's fp32 expression
    ok4, rvec4, tvec4, inl4 = cv2.solvePnPRansac(world[msk.reshape(-1)], pix[msk], K, None, iterationsCount=iterations,
                                                 reprojectionError=5, flags=cv2.SOLVEPNP_SQPNP)
    ok5, rvec5, tvec5, n5 = pr.solve_frame(frame, K, iterations=iterations, reproj=5)
    assert pr.enabled() and bool(ok4) == bool(ok5)
    if ok4:
        assert n5 == len(inl4) and np.array_equal(rvec4, rvec5) and np.array_equal(tvec4, tvec5)


def test_rank_emulated_frame_window_sharding(cuda_device):
    """Two frame-window shards driven from one process on one GPU (their accumulator blocks summed in-process
    instead of by NCCL) take exactly the steps of the unsharded problem."""
    fx, scene, prob = golden_case("small_flow")
    kw = fx["kw"]
    full = build(scene, kw)
    shards = [build(scene, kw, shard=dict(world=2, rank=r, group=None)) for r in range(2)]
    for net in [full] + shards:
        set_params(net, fx["init_params"])
        net._bufs.reset_adam()
        net._reduce_accumulators = lambda: None
    assert shards[0]._plan.n_local_frames + shards[1]._plan.n_local_frames == prob.n_imgs
    mode = _lib.M_FLOW_ACTIVE | _lib.M_APPLY_ADAM
    for it in range(3):
        adam = adam_schedule([0.01], start_step=it, first_iter=it)
        full._phase_accumulate(mode)
        full._phase_update(mode, adam)
        for net in shards:
            net._phase_accumulate(mode)
        total = shards[0]._bufs.acc + shards[1]._bufs.acc
        for net in shards:
            net._bufs.acc.copy_(total)
            net._phase_update(mode, adam)
    for it in range(3):
        a, b0, b1 = (float(n._bufs.loss_hist[it]) for n in [full] + shards)
        assert b0 == b1 and abs(a - b0) <= 1e-6 * abs(a)
    for k in ("pw_poses", "im_poses", "im_focals"):
        assert torch.equal(getattr(shards[0], k), getattr(shards[1], k)), k
        assert torch.allclose(getattr(full, k), getattr(shards[0], k), rtol=0, atol=5e-6), k
    depth = torch.cat([s.im_depthmaps.detach() for s in shards])
    assert torch.allclose(depth, full.im_depthmaps.detach(), rtol=0, atol=5e-6)


def test_full_size_properties_c2(cuda_device):
    """BASELINE config C2 (50 frames, 288x512, swin-5, flow + temporal): size-independent properties."""
    from align3r_b200.synth import CONFIGS
    n, H, W, graph = CONFIGS["c2"]
    scene = make_scene(n, H, W, graph, seed=0, with_flow=True)
    kw = dict(CANONICAL_KW)
    net = build(scene, kw)
    assert net.n_edges == 470 and net._plan.pixel_slots() == 2 * 470 * H * W
    net.compute_global_alignment(init='mst', niter=0)
    l0 = float(net(epoch=0))
    l0b = float(net(epoch=0))
    assert abs(l0 - l0b) <= 1e-6 * abs(l0)                        # only the order of fp64 atomics may differ
    lf = float(net(epoch=300))
    assert math.isfinite(lf) and lf > l0 and not net.flow_loss_flag   # flow term adds a positive amount, gate stays open
    # edge-order invariance: the same scene with its pair list reversed gives the same loss
    perm = list(range(net.n_edges))[::-1]
    out = scene.output
    v1 = {k: ([v[p] for p in perm] if isinstance(v, list) else v[perm]) for k, v in out["view1"].items()}
    v2 = {k: ([v[p] for p in perm] if isinstance(v, list) else v[perm]) for k, v in out["view2"].items()}
    p1 = {k: v[perm] for k, v in out["pred1"].items()}
    p2 = {k: v[perm] for k, v in out["pred2"].items()}
    net2 = global_aligner(dict(view1=v1, view2=v2, pred1=p1, pred2=p2), "cuda", flow=(scene.flow_ij[perm], scene.flow_ji[perm]), **kw)
    with torch.no_grad():
        net2.pw_poses.data.copy_(net.pw_poses.data[perm])
        for k in ("im_depthmaps", "im_poses", "im_focals"):
            getattr(net2, k).data.copy_(getattr(net, k).data)
    assert abs(float(net2(epoch=300)) - lf) <= 1e-6 * abs(lf)
    del net2
    # 60 iterations of the canonical schedule end with finite, sane geometry.  (The loss is NOT monotone: Adam's
    # first lr=0.01 steps kick it up before it settles - the reference does the same on this kind of scene.)
    final = net.compute_global_alignment(init=None, niter=60, schedule='linear', lr=0.01)
    hist = net._bufs.loss_hist[:60].cpu()
    assert math.isfinite(final) and bool(torch.isfinite(hist).all()) and final < 3 * float(hist[0])
    assert float(hist[40:].mean()) < float(hist[1:11].mean())
    depth = torch.stack(net.get_depthmaps()).detach()
    assert bool(torch.isfinite(depth).all()) and float(depth.min()) > 0
    gt_depth = torch.stack([scene.gt.depth(k) for k in range(n)]).float()
    dm = go.depth_metrics(depth.cpu().numpy()[::5], gt_depth.numpy()[::5])
    pm = go.pose_metrics(net.get_im_poses().detach().cpu().numpy(), scene.gt.cam2world)
    print("C2 after 60 iterations:", dm, pm)
    assert dm["abs_rel"] < 0.1 and dm["d1"] > 0.9, dm
    assert pm["ate"] < 0.2 and pm["rpe_rot"] < 2.0, pm


# ---------------------------------------------------------------------------------------------------------------
# preset camera poses against the fixture the real reference produced
# ---------------------------------------------------------------------------------------------------------------
def test_mst_initialisation_with_frozen_known_poses(cuda_device):
    fx, scene, prob = golden_case("small_known_poses")
    ref = fx["mst_frozen"]["params"]
    net = build(scene, fx["kw"])
    net.preset_pose(fx["known_c2w"])
    assert not net.im_poses.requires_grad and net.norm_pw_scale is False
    init_fun.init_minimum_spanning_tree(net, niter_PnP=10)
    assert float((net.get_im_poses().detach().cpu() - fx["known_c2w"]).abs().max()) < 1e-5
    prob.norm_pw_scale = False                                 # preset_pose switches the pair-scale normalisation off
    T_ref = go.pw_pose_matrices(prob, ref["pw_poses"])
    assert float((net.get_pw_poses().detach().cpu() - T_ref).abs().max()) < 2e-3      # PnP-initialised frames: fp32 sensitivity
    d = (net.im_depthmaps.detach().cpu()[:, :prob.P] - ref["im_depthmaps"]).abs()
    assert float(d.median()) < 1e-4 and float(d.max()) < 5e-3
    assert rel_err(net.get_focals().detach().cpu(), go.focals(prob, ref["im_focals"])) < 1e-5
    # the loop then runs with the poses frozen and ends on a finite loss
    loss = net.compute_global_alignment(init=None, niter=10, schedule="linear", lr=0.01)
    assert loss == loss and loss < 1e3
    assert float((net.get_im_poses().detach().cpu() - fx["known_c2w"]).abs().max()) < 1e-5


def test_known_poses_initialisation(cuda_device):
    fx, scene, prob = golden_case("small_known_poses")
    ref = fx["known_poses"]["params"]
    net = build(scene, fx["kw"])
    net.camera_poses = [p for p in fx["known_c2w"]]
    loss = net.compute_global_alignment(init="known_poses", niter=0)
    assert net.im_poses.requires_grad
    assert float((net.get_im_poses().detach().cpu() - fx["known_c2w"]).abs().max()) < 1e-5
    prob.norm_pw_scale = False
    T_ref = go.pw_pose_matrices(prob, ref["pw_poses"])
    assert float((net.get_pw_poses().detach().cpu() - T_ref).abs().max()) < 2e-3
    d = (net.im_depthmaps.detach().cpu()[:, :prob.P] - ref["im_depthmaps"]).abs()
    assert float(d.median()) < 1e-5


# ---------------------------------------------------------------------------------------------------------------
# depth prior against the fixture the real reference produced
# ---------------------------------------------------------------------------------------------------------------
def test_depth_prior_teacher_forced_and_free_run(cuda_device):
    fx, scene, prob = golden_case("small_depth_prior")
    fr, g32, g64 = fx["free_run"], fx["tf32_final"], fx["tf64_final"]
    net = build(scene, fx["kw"])
    with pytest.raises(RuntimeError, match="initial depth"):
        net(epoch=0)                                            # no initialisation yet: nothing to regularise towards
    set_params(net, fr["params"])
    P = prob.P
    with torch.no_grad():                                        # the reference's initial depth maps
        net._bufs.logdepth_init[:, :P] = fr["init_depth"][net._plan.lframe_id.tolist(), :P].clamp(min=1e-6).log().to(net.device)
    net._prior_ready = True
    loss = net(epoch=g32["epoch"])
    loss.backward()
    assert abs(float(loss.detach()) - float(g64["loss"])) <= 2e-6 * abs(float(g64["loss"]))
    assert float(net._bufs.scal[_lib.S_PRIOR]) > 0
    mine = grads_of(net, P)
    for k in mine:
        ok, err, ref_err = grad_close(k, mine[k], g32["grads"][k], g64["grads"][k])
        assert ok, (k, err, ref_err)
    # free run through our own initialisation ends where the reference ends
    net2 = build(scene, fx["kw"])
    final = net2.compute_global_alignment(init="mst", niter=fx["niter"], schedule="linear", lr=0.01)
    assert abs(final - fr["loss"]) <= 2e-3 * abs(fr["loss"])
    d = (net2.im_depthmaps.detach().cpu()[:, :P] - fr["params"]["im_depthmaps"]).abs()
    assert float(d.median()) < 2e-3


# ---------------------------------------------------------------------------------------------------------------
# the two organisations of the scan (A3R_SCAN=v1: thread owns pixels; default v4: warp streams) agree on the device
# ---------------------------------------------------------------------------------------------------------------
_V1_SNIPPET = """
import sys, torch
sys.path.insert(0, {root!r})
from tests.test_gpu_parity import build, set_params, grads_of
from tests.helpers import golden_case
from align3r_b200 import _lib
assert _lib.load().a3r_tile_pixels() == 1024          # the v1 kernel's tile
fx, scene, prob = golden_case("small_flow")
net = build(scene, fx["kw"])
set_params(net, fx["init_params"])
loss = net(epoch=fx["tf32_flow"]["epoch"])
loss.backward()
torch.save(dict(loss=float(loss.detach()), grads=grads_of(net, prob.P)), {out!r})
"""


def test_scan_v1_cross_checks_v4(cuda_device, tmp_path):
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = str(tmp_path / "v1.pt")
    subprocess.run([sys.executable, "-c", _V1_SNIPPET.format(root=root, out=out)], check=True, timeout=600,
                   env=dict(os.environ, A3R_SCAN="v1"), cwd=root)
    v1 = torch.load(out, weights_only=False)
    assert _lib.load().a3r_tile_pixels() == 128
    fx, scene, prob = golden_case("small_flow")
    net = build(scene, fx["kw"])
    set_params(net, fx["init_params"])
    loss = net(epoch=fx["tf32_flow"]["epoch"])
    loss.backward()
    assert abs(float(loss.detach()) - v1["loss"]) <= 1e-6 * abs(v1["loss"])
    mine = grads_of(net, prob.P)
    for k in mine:
        assert rel_err(mine[k], v1["grads"][k]) <= 2e-5, (k, rel_err(mine[k], v1["grads"][k]))


def test_scan_v5_cross_checks_v4(cuda_device, tmp_path):
    """A3R_SCAN=v5 (csrc/ga_scan_v5.cuh: class-specialised warps - align warps and flow warps walk the tiles separately, 8
    pixels per lane, three partitions of the tile list) against the product scan: same loss and gradients, flow term on."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = str(tmp_path / "v5.pt")
    snippet = _V1_SNIPPET.replace("== 1024          # the v1 kernel's tile", "== 256           # the v5 kernel's tile")
    subprocess.run([sys.executable, "-c", snippet.format(root=root, out=out)], check=True, timeout=600,
                   env=dict(os.environ, A3R_SCAN="v5"), cwd=root)
    v5 = torch.load(out, weights_only=False)
    fx, scene, prob = golden_case("small_flow")
    net = build(scene, fx["kw"])
    set_params(net, fx["init_params"])
    loss = net(epoch=fx["tf32_flow"]["epoch"])
    loss.backward()
    assert abs(float(loss.detach()) - v5["loss"]) <= 1e-6 * abs(v5["loss"])
    mine = grads_of(net, prob.P)
    for k in mine:
        assert rel_err(mine[k], v5["grads"][k]) <= 2e-5, (k, rel_err(mine[k], v5["grads"][k]))


# ---------------------------------------------------------------------------------------------------------------
# clean_pointcloud against the fixture the real reference produced
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tol", [0.001, 0.05])
def test_clean_pointcloud(cuda_device, tol):
    fx, scene, prob = golden_case("small_clean_pointcloud")
    net = build(scene, fx["kw"])
    set_params(net, fx["params"])
    before = torch.stack([c.clone() for c in net.im_conf]).cpu()
    assert torch.allclose(before, fx["before"]["im_conf"], rtol=1e-6, atol=1e-6)
    assert net.clean_pointcloud(tol=tol) is net
    after = torch.stack(list(net.im_conf)).cpu()
    ref = fx[f"after_tol{tol}"]
    # a projected coordinate within fp32 rounding of x.5, or a depth within rounding of the (1 - tol) test, may fall on
    # the other side: a handful of pixels out of 18 432 (stated tolerance: 0.2 % of the pixels)
    differ = (after - ref).abs() > 1e-6
    assert float(differ.float().mean()) <= 2e-3, int(differ.sum())
    assert int(((after - before).abs() > 1e-6).sum()) > 1000


# ---------------------------------------------------------------------------------------------------------------
# result write-out: the on-disk formats of reference base_opt.py:332-427 / vo_eval.py:308-316
# ---------------------------------------------------------------------------------------------------------------
def test_result_write_out(cuda_device, tmp_path):
    scene = make_scene(5, 40, 56, "swin-2-noncyclic", seed=2, with_flow=True, with_img=True)
    net = build(scene, dict(CANONICAL_KW, num_total_iter=5))
    net.compute_global_alignment(init="mst", niter=5, schedule="linear", lr=0.01)
    d = str(tmp_path)
    tum = net.save_tum_poses(f"{d}/pred_traj.txt")
    rows = [l.split() for l in open(f"{d}/pred_traj.txt").read().splitlines()]
    assert len(rows) == 5 and all(len(r) == 8 for r in rows) and [float(r[0]) for r in rows] == [0, 1, 2, 3, 4]
    back = np.array([[float(x) for x in r[1:]] for r in rows])
    assert np.allclose(back, tum, rtol=0, atol=1e-12)                      # x y z qw qx qy qz, full precision
    c2w = net.get_im_poses().detach().cpu().numpy()
    assert np.allclose(back[:, :3], c2w[:, :3, 3], atol=1e-6)
    from scipy.spatial.transform import Rotation
    assert np.allclose(Rotation.from_quat(back[:, [4, 5, 6, 3]]).as_matrix(), c2w[:, :3, :3], atol=1e-5)
    net.save_focals(f"{d}/pred_focal.txt")
    assert np.allclose(np.loadtxt(f"{d}/pred_focal.txt"), net.get_focals().detach().cpu().numpy().ravel(), atol=1e-6 + 5e-7)
    net.save_intrinsics(f"{d}/pred_intrinsics.txt")
    net.save_intrinsics(f"{d}/pred_intrinsics.txt")                        # appended: two clips, one file
    K = np.loadtxt(f"{d}/pred_intrinsics.txt")
    assert K.shape == (10, 9) and np.allclose(K[:5], K[5:]) and np.allclose(K[:5].reshape(5, 3, 3), net.get_intrinsics().detach().cpu().numpy(), atol=1e-6)
    net.save_conf_maps(d, 3)
    net.save_init_conf_maps(d, 3)
    assert np.allclose(np.load(f"{d}/conf_3.npy"), net.get_conf()[0].cpu().numpy()) and os.path.exists(f"{d}/init_conf_7.npy")
    net.save_depth_maps(d, 0)
    assert np.allclose(np.load(f"{d}/frame_0004.npy"), net.get_depthmaps()[4].detach().cpu().numpy())
    assert os.path.exists(f"{d}/frame_0000.png") and os.path.exists(f"{d}/_depth_maps.gif")
    net.save_dynamic_masks(d, 0)
    net.save_rgb_imgs(d, 0)
    assert os.path.exists(f"{d}/dynamic_mask_4.png") and os.path.exists(f"{d}/frame_0002_rgb.png") and os.path.exists(f"{d}/_rgb.gif")


# ---------------------------------------------------------------------------------------------------------------
# the plain aligner of tool/depth_test.py (reference dust3r/cloud_opt): same kernels, flow / temporal terms off,
# non-symmetrised complete clip graph, cosine schedule down to 1e-6
# ---------------------------------------------------------------------------------------------------------------
def test_cloud_opt_variant_against_live_oracle(cuda_device):
    from align3r_b200 import cloud_opt
    from align3r_b200.synth import make_edge_list
    edges = make_edge_list(6, "complete", symmetrize=False)
    scene = make_scene(6, 48, 64, "complete", seed=4, with_flow=False, edges=edges)
    assert len(scene.edges) == 15
    net = cloud_opt.global_aligner(scene.output, False, [], "cuda", mode=cloud_opt.GlobalAlignerMode.PointCloudOptimizer,
                                   verbose=False, min_conf_thr=3)
    assert not net.is_symmetrized and net.flow_loss_weight == 0 and net.temporal_smoothing_weight == 0
    prob = go.problem_from_output(scene.output, None, None, flow_loss_weight=0.0, temporal_smoothing_weight=0.0)
    params, _ = go.mst_init(prob)
    # teacher-forced at the oracle's initialisation: loss and gradients on a graph whose frames carry unequal numbers
    # of side-i and side-j slots (frame 0 has no side-i slot at all)
    p64 = {k: v.double() for k, v in params.items()}
    loss64, g64, _ = go.loss_and_grads(prob.to(torch.float64), p64, 9999)
    _, g32, _ = go.loss_and_grads(prob, params, 9999)
    set_params(net, params)
    l0 = net(epoch=9999)
    l0.backward()
    assert abs(float(l0.detach()) - float(loss64)) <= 2e-6 * abs(float(loss64))
    mine = grads_of(net, prob.P)
    for k in mine:
        ok, err, ref_err = grad_close(k, mine[k], g32[k], g64[k])
        assert ok, (k, err, ref_err)
    # the loop from that initialisation (lr 0.05, cosine to 1e-6) ends where the oracle ends
    loss = net.compute_global_alignment(init=None, niter=30, schedule="cosine", lr=0.05)
    final, losses, _ = go.run_alignment(prob, params, niter=30, schedule="cosine", lr=0.05, lr_min=1e-6)
    assert abs(loss - losses[-1]) <= 1e-3 * abs(losses[-1]), (loss, losses[-1])
    d = (net.im_depthmaps.detach().cpu()[:, :prob.P] - final["im_depthmaps"]).abs()
    assert float(d.median()) < 1e-4
    # and the canonical call through our own initialisation ends on a finite, comparable loss
    net2 = cloud_opt.global_aligner(scene.output, False, [], "cuda", verbose=False, min_conf_thr=3)
    loss2 = net2.compute_global_alignment(init="mst", niter=30, schedule="cosine", lr=0.05)
    assert math.isfinite(loss2) and loss2 < 5 * losses[-1], (loss2, losses[-1])
    with pytest.raises(NotImplementedError):
        cloud_opt.global_aligner(scene.output, True, [], "cuda")


# ---------------------------------------------------------------------------------------------------------------
# init_priors (hierarchical clip driver start from the key frame)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["a", "b"])
def test_mst_initialisation_with_priors(cuda_device, tag):
    fx, scene, prob = golden_case(f"small_init_priors_{tag}")
    ref = fx["params"]
    net = build(scene, fx["kw"])
    init_fun.init_minimum_spanning_tree(net, niter_PnP=10, init_priors=fx["init_priors"])
    T_ref = go.pw_pose_matrices(prob, ref["pw_poses"])
    assert float((net.get_pw_poses().detach().cpu() - T_ref).abs().max()) < 5e-5
    P_ref = go.poses_to_matrices(ref["im_poses"])
    Pm = net.get_im_poses().detach().cpu()
    assert float((Pm - P_ref).abs().max()) < 2e-3
    # the key frame keeps the prior pose up to the scale normalisation of its translation
    key = torch.from_numpy(fx["init_priors"][0])
    assert float((Pm[0, :3, :3] - key[:3, :3]).abs().max()) < 1e-5
    assert rel_err(net.get_focals().detach().cpu(), go.focals(prob, ref["im_focals"])) < 1e-5
    d = (net.im_depthmaps.detach().cpu()[:, :prob.P] - ref["im_depthmaps"]).abs()
    assert float(d.median()) < 1e-5 and float(d.max()) < 2e-3


# ---------------------------------------------------------------------------------------------------------------
# BASELINE.json configs[0] at FULL size against the real reference's own 300-iteration run
# (fixture: oracle/make_golden_c1.py, 6.4 minutes of the unmodified reference on 8 host threads)
# ---------------------------------------------------------------------------------------------------------------
def test_c1_full_size_against_reference_run(cuda_device):
    """10-frame 288x512 swin-3 sequence, init='mst', niter=300, canonical hyper-parameters: what a user reads out after
    the run equals what the reference produced.  Stated tolerances (end state after 300 Adam steps from an MST init that
    differs by fp32 rounding): final loss 2e-3 relative, median |log-depth| 2e-3, camera matrices 1e-2, focal 5e-3
    relative, Abs Rel / log RMSE / delta<1.25 2e-3 absolute, ATE 5e-3 scene units, RPE rotation 0.1 degree."""
    fx = load_golden("c1_reference_run")
    scene = make_scene(**fx["recipe"])
    assert scene.edges == [tuple(e) for e in fx["edges"]]
    assert abs(scene_checksum(scene) - fx["checksum"]) <= 1e-9 * abs(fx["checksum"]), "synthetic generator drifted"
    torch.manual_seed(fx["seed"])
    net = build(scene, fx["kw"])
    loss = net.compute_global_alignment(init='mst', niter=fx["niter"], schedule='linear', lr=0.01)
    assert net.flow_loss_flag == fx["flow_dropped"]
    assert abs(loss - fx["loss"]) <= 2e-3 * abs(fx["loss"]), (loss, fx["loss"])
    n, H, W = fx["recipe"]["n_frames"], fx["recipe"]["H"], fx["recipe"]["W"]
    depth = torch.stack(net.get_depthmaps()).detach().cpu()
    assert depth.shape == (n, H, W)
    s = fx["stride"]
    d = (depth[:, ::s, ::s].log() - fx["logdepth_sample"]).abs()
    assert float(d.median()) < 2e-3 and float(d.max()) < 5e-2, (float(d.median()), float(d.max()))
    assert abs(float(depth.log().double().mean()) - fx["logdepth_mean"]) < 2e-3
    poses = net.get_im_poses().detach().cpu()
    assert float((poses - fx["im_poses_4x4"]).abs().max()) < 1e-2
    assert rel_err(net.get_focals().detach().cpu(), fx["focals"]) < 5e-3
    assert float((net.get_pw_poses().detach().cpu() - fx["pw_poses_4x4"]).abs().max()) < 1e-2
    gt_depth = np.stack([scene.gt.depth(k).numpy() for k in range(n)])
    dm = go.depth_metrics(depth.numpy(), gt_depth)
    for k in ("abs_rel", "log_rmse", "d1"):
        assert abs(dm[k] - fx["depth_metrics"][k]) <= 2e-3, (k, dm[k], fx["depth_metrics"][k])
    pm = go.pose_metrics(poses.numpy(), scene.gt.cam2world)
    assert abs(pm["ate"] - fx["pose_metrics"]["ate"]) <= 5e-3 and abs(pm["rpe_rot"] - fx["pose_metrics"]["rpe_rot"]) <= 0.1, (pm, fx["pose_metrics"])


# ---------------------------------------------------------------------------------------------------------------
# use_self_mask=True: the dynamic masks estimated from the pairs themselves (reference optimizer.py:156-235)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["a", "b"])
def test_self_estimated_dynamic_masks(cuda_device, tag):
    """Masks: at most 0.1 % of the pixels may differ from the reference's (a pixel whose mean normalised error sits within
    fp32 rounding of the 0.35 threshold); measured: identical.  With them the loop takes the oracle's steps from the same
    initial parameters (final loss 5e-5 relative), and the canonical call ends near the reference's own result (20
    unconverged iterations from an MST init that differs by fp32 rounding: loss 1e-2 relative, cameras 1e-2)."""
    fx, scene, prob = golden_case(f"small_self_mask_{tag}")
    net = build(scene, fx["kw"])
    ref = fx["dynamic_masks"]
    ours = torch.stack([m for m in net.dynamic_masks]).cpu()
    assert ours.shape == ref.shape and ours.dtype == torch.bool
    assert int((ours != ref).sum()) <= 1e-3 * ref.numel(), int((ours != ref).sum())
    static = net._bufs.static_mask[:, :prob.P].cpu().view(prob.n_imgs, prob.H, prob.W).bool()
    assert torch.equal(static, ~ours)
    # the loop with these masks against the oracle, from the same initial parameters
    prob.dynamic_masks = ref.clone()
    params, _ = go.mst_init(prob)
    _, losses, dropped = go.run_alignment(prob, params, niter=fx["niter"], schedule="linear", lr=0.01)
    set_params(net, params)
    loss = net.compute_global_alignment(init=None, niter=fx["niter"], schedule='linear', lr=0.01)
    assert net.flow_loss_flag == dropped
    assert abs(loss - losses[-1]) <= 5e-5 * abs(losses[-1]), (loss, losses[-1])
    # the canonical call against the reference's own run
    net2 = build(scene, fx["kw"])
    loss2 = net2.compute_global_alignment(init='mst', niter=fx["niter"], schedule='linear', lr=0.01)
    fr = fx["free_run"]
    assert net2.flow_loss_flag == fr["flow_dropped"]
    assert abs(loss2 - fr["loss"]) <= 1e-2 * abs(fr["loss"]), (loss2, fr["loss"])
    assert float((net2.get_im_poses().detach().cpu() - fr["im_poses_4x4"]).abs().max()) < 1e-2
    assert rel_err(net2.get_focals().detach().cpu(), fr["focals"]) < 5e-3


# ---------------------------------------------------------------------------------------------------------------
# long video: many frames and edges (several items per thread in K2, cameras beyond 48 KB of shared memory)
# ---------------------------------------------------------------------------------------------------------------
def test_long_video_small_frames_against_live_oracle(cuda_device):
    """320 frames of 16x32 pixels, swin-5 (3170 directed edges): the loop's small-parameter kernel walks several chain items
    per thread and keeps 320 cameras (51 KB) in shared memory; the MST initialisation, the teacher-forced loss / gradients
    and the per-iteration losses of a ten-iteration free run are compared with the oracle evaluated live."""
    scene = make_scene(320, 16, 32, "swin-5-noncyclic", seed=11, with_flow=True)
    kw = dict(CANONICAL_KW, num_total_iter=10)
    prob = go.problem_from_output(scene.output, scene.flow_ij, scene.flow_ji, **kw)
    assert prob.n_edges == 3170
    params, _ = go.mst_init(prob)
    net = build(scene, kw)
    net.verbose = False
    set_params(net, params)
    p64 = {k: v.double() for k, v in params.items()}
    loss64, g64, _ = go.loss_and_grads(prob.to(torch.float64), p64, 9999)
    _, g32, _ = go.loss_and_grads(prob, params, 9999)
    loss = net(epoch=9999)
    loss.backward()
    assert abs(float(loss.detach()) - float(loss64)) <= 2e-6 * abs(float(loss64))
    for k in go.TRAINABLE:
        m = getattr(net, k).grad.detach().cpu()
        m = m[:, :prob.P] if k == "im_depthmaps" else m.reshape(g64[k].shape)
        ok, err, ref_err = grad_close(k, m, g32[k], g64[k])
        assert ok, (k, err, ref_err)
    net2 = build(scene, kw)
    net2.verbose = False
    set_params(net2, params)
    net2.compute_global_alignment(init=None, niter=10, schedule="linear", lr=0.01)
    ref_full = go.run_alignment(prob, params, niter=10, schedule="linear", lr=0.01)[1]
    ours = net2._bufs.loss_hist[:10].cpu().tolist()
    for it, (a, b) in enumerate(zip(ours, ref_full)):       # 5e-5 while the runs are in lockstep, 2e-3 over all ten
        assert abs(a - b) <= (5e-5 if it < 3 else 2e-3) * abs(b), (it, ours, ref_full)


def test_mst_initialisation_on_a_non_symmetrised_graph(cuda_device):
    """Edges (0,1), (0,2), (1,2): frame 2 is never the first image of a pair, so it has no Weiszfeld focal and takes the
    tentative focal its RANSAC-PnP solve scored best (reference init_im_poses.py:272-279, 482-504); the oracle's
    initialisation equals the real flow-free reference's bit for bit on this case (oracle/make_golden_hierarchical.py)."""
    import copy
    from align3r_b200.cloud_opt import global_aligner as plain_aligner
    sc = make_scene(11, 32, 48, "complete", seed=4, with_flow=False, edges=[(0, 4), (0, 8), (4, 8)])
    out = copy.deepcopy(sc.output)
    out["view1"]["idx"], out["view2"]["idx"] = [0, 0, 1], [1, 2, 2]
    prob = go.problem_from_output(out, None, None, min_conf_thr=3.0)
    params, _ = go.mst_init(prob)
    net = plain_aligner(copy.deepcopy(out), False, [], "cuda", mode=GlobalAlignerMode.PointCloudOptimizer, verbose=False, min_conf_thr=3)
    net.compute_global_alignment(init="mst", niter=0)
    assert float((net.im_focals.detach().cpu().reshape(-1) - params["im_focals"].reshape(-1)).abs().max()) < 1e-4
    assert float((net.im_poses.detach().cpu() - params["im_poses"]).abs().max()) < 1e-5
    assert float((net.pw_poses.detach().cpu() - params["pw_poses"]).abs().max()) < 1e-5
    assert float((net.im_depthmaps.detach().cpu()[:, :prob.P] - params["im_depthmaps"]).abs().max()) < 1e-5


# ---------------------------------------------------------------------------------------------------------------
# variable image shapes (reference optimizer.py:66-92, 601-607): fixture from the real reference, oracle/make_golden_ragged.py
# ---------------------------------------------------------------------------------------------------------------
def test_ragged_image_shapes_against_reference(cuda_device):
    """Frames of 40x56 and 32x56 pixels in one scene, inputs as per-edge lists (flow off: the reference's flow term needs
    uniform shapes).  MST initialisation, teacher-forced loss / gradients (fp64 arbiter criterion) and a 15-iteration free
    run against the reference's own numbers."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    from make_golden_ragged import RECIPE, ragged_output
    fx = load_golden("small_ragged")
    scene = make_scene(**RECIPE)
    assert abs(scene_checksum(scene) - fx["checksum"]) <= 1e-9 * abs(fx["checksum"])
    out = ragged_output(scene)

    def fresh():
        import copy
        return global_aligner(copy.deepcopy(out), "cuda", mode=GlobalAlignerMode.PointCloudOptimizer, **fx["kw"])
    net = fresh()
    assert [tuple(s) for s in net.imshapes] == fx["imshapes"]
    areas = [h * w for h, w in fx["imshapes"]]
    P = max(areas)
    # MST initialisation
    init_fun.init_minimum_spanning_tree(net, niter_PnP=10)
    ref = fx["init_params"]
    assert float((net.pw_poses.detach().cpu() - ref["pw_poses"]).abs().max()) < 5e-5
    assert float((net.im_poses.detach().cpu() - ref["im_poses"]).abs().max()) < 5e-5
    assert float((net.im_focals.detach().cpu().reshape(-1) - ref["im_focals"].reshape(-1)).abs().max()) < 1e-4
    for n, a in enumerate(areas):
        d = (net.im_depthmaps.detach().cpu()[n, :a] - ref["im_depthmaps"][n, :a]).abs()
        assert float(d.max()) < 2e-3 and float(d.median()) < 1e-5, (n, float(d.max()))
    # teacher-forced loss and gradients at the reference's initial parameters
    net = fresh()
    set_params(net, ref)
    loss = net(epoch=9999)
    loss.backward()
    assert abs(float(loss.detach()) - float(fx["tf64"]["loss"])) <= 2e-6 * abs(float(fx["tf64"]["loss"]))
    for k in ("pw_poses", "im_depthmaps", "im_poses", "im_focals"):
        m = getattr(net, k).grad.detach().cpu()
        g32, g64 = fx["tf32"]["grads"][k], fx["tf64"]["grads"][k]
        if k == "im_depthmaps":
            m = m[:, :P].clone()
            for n, a in enumerate(areas):
                assert float(m[n, a:].abs().max()) == 0 if a < P else True      # the padding carries no gradient
        else:
            m = m.reshape(g64.shape)
        ok, err, ref_err = grad_close(k, m, g32, g64)
        assert ok, (k, err, ref_err)
    # free run: from the reference's initial parameters the loop ends on the reference's result (5e-5); through our own MST
    # initialisation (canonical call) within the end-state tolerance of a 15-iteration, still fast-moving run (1e-2)
    fr = fx["free_run"]
    net = fresh()
    set_params(net, ref)
    final = net.compute_global_alignment(init=None, niter=fx["niter"], schedule='linear', lr=0.01)
    assert abs(final - fr["loss"]) <= 5e-5 * abs(fr["loss"]), (final, fr["loss"])
    assert float((net.get_im_poses().detach().cpu() - fr["im_poses_4x4"]).abs().max()) < 1e-4
    net = fresh()
    final = net.compute_global_alignment(init='mst', niter=fx["niter"], schedule='linear', lr=0.01)
    assert abs(final - fr["loss"]) <= 1e-2 * abs(fr["loss"]), (final, fr["loss"])
    assert float((net.get_im_poses().detach().cpu() - fr["im_poses_4x4"]).abs().max()) < 1e-2
    for d, r in zip(net.get_depthmaps(), fr["depths"]):
        assert d.shape == r.shape
        assert float((d.detach().cpu().log() - r.log()).abs().median()) < 2e-3


def test_graft_entry_smoke(cuda_device):
    """The driver's smoke() entry point stays green (it runs before the bench at round end)."""
    import __graft_entry__ as entry
    entry.smoke()


_K2_SCRATCH_SNIPPET = """
import sys, torch
sys.path.insert(0, {root!r})
from tests.test_gpu_parity import loop_case
net, niter = loop_case({case!r})
loss = net.compute_global_alignment(init=None, niter=niter, schedule="linear", lr=0.01)
torch.save(dict(loss=loss, hist=net._bufs.loss_hist[:niter].cpu(), im_poses=net.im_poses.detach().cpu(),
                pw_poses=net.pw_poses.detach().cpu(), depth=net.im_depthmaps.detach().cpu()), {out!r})
"""


def loop_case(case):
    """'small': the 48x64 reference fixture (general form of the scan); 'rowtile': a 128-pixel-wide scene (the row-tile scan)."""
    if case == "small":
        fx, scene, prob = golden_case("small_flow")
        net = build(scene, fx["kw"])
        set_params(net, fx["init_params"])
        return net, fx["niter"]
    scene = make_scene(6, 24, 128, "swin-2-noncyclic", seed=21, with_flow=True)
    kw = dict(CANONICAL_KW, num_total_iter=16)
    prob = go.problem_from_output(scene.output, scene.flow_ij, scene.flow_ji, **kw)
    params, _ = go.mst_init(prob)
    net = build(scene, kw)
    net.verbose = False
    set_params(net, params)
    return net, 16


@pytest.mark.parametrize("case", ["small", "rowtile"])
def test_loop_organisation_changes_no_bit(cuda_device, tmp_path, case):
    """The loop's organisation is invisible in the results: the depth Adam step folded into the next scan's tile prologue
    (vs the stand-alone K3 kernel, A3R_FOLD_DEPTH=0) and programmatic dependent launches (vs plain stream order, A3R_PDL=0)
    - both switched off in a separate process - give the same loss history and the same parameters, bit for bit."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = str(tmp_path / "plain.pt")
    subprocess.run([sys.executable, "-c", _K2_SCRATCH_SNIPPET.format(root=root, out=out, case=case)], check=True, timeout=600,
                   env=dict(os.environ, A3R_FOLD_DEPTH="0", A3R_PDL="0"), cwd=root)
    alt = torch.load(out, weights_only=False)
    net, niter = loop_case(case)
    loss = net.compute_global_alignment(init=None, niter=niter, schedule="linear", lr=0.01)
    assert torch.equal(net._bufs.loss_hist[:niter].cpu(), alt["hist"]) and loss == alt["loss"]
    assert torch.equal(net.im_poses.detach().cpu(), alt["im_poses"]) and torch.equal(net.pw_poses.detach().cpu(), alt["pw_poses"])
    assert torch.equal(net.im_depthmaps.detach().cpu(), alt["depth"])


# ---------------------------------------------------------------------------------------------------------------
# the row-tile specialisation of the scan (widths that are multiples of the 128-pixel tile: the configuration the benchmark
# runs) against the general form of the same kernel and against the oracle
# ---------------------------------------------------------------------------------------------------------------
_ROWT_SNIPPET = """
import sys, torch
sys.path.insert(0, {root!r})
from tests.test_gpu_parity import build, grads_of, set_params
from tests.helpers import make_scene, CANONICAL_KW, go
scene = make_scene(6, 24, 128, "swin-2-noncyclic", seed=21, with_flow=True)
kw = dict(CANONICAL_KW, num_total_iter=12)
prob = go.problem_from_output(scene.output, scene.flow_ij, scene.flow_ji, **kw)
params, _ = go.mst_init(prob)
net = build(scene, kw)
net.verbose = False
set_params(net, params)
loss = net(epoch=9999)
loss.backward()
g = grads_of(net, prob.P)
run = net.compute_global_alignment(init=None, niter=12, schedule="linear", lr=0.01)
torch.save(dict(loss=float(loss.detach()), grads=g, run=run, flags=int(net._bufs.flags(net._trainable()))), {out!r})
"""


def test_row_tile_scan_against_general_form(cuda_device, tmp_path):
    """W = 128: every tile lies in one image row, the product runs the row-tile scan (A3R_F_ROW_TILES; against the oracle:
    test_medium_scene_against_live_oracle and the full-size fixtures).  The same scene through the general form of the kernel
    (A3R_V4_ROWT=0, separate process): same loss (1e-6), same gradients (2e-5: fp32 summation order), and a 12-iteration
    run that ends on the same loss (1e-4)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = {}
    for tag, env in (("general", dict(A3R_V4_ROWT="0")), ("rowtile", {})):
        out = str(tmp_path / f"{tag}.pt")
        subprocess.run([sys.executable, "-c", _ROWT_SNIPPET.format(root=root, out=out)], check=True, timeout=600,
                       env=dict(os.environ, **env), cwd=root)
        res[tag] = torch.load(out, weights_only=False)
    a, b = res["rowtile"], res["general"]
    assert a["flags"] & _lib.F_ROW_TILES
    assert abs(a["loss"] - b["loss"]) <= 1e-6 * abs(b["loss"])
    for k in a["grads"]:
        assert rel_err(a["grads"][k], b["grads"][k]) <= 2e-5, (k, rel_err(a["grads"][k], b["grads"][k]))
    assert abs(a["run"] - b["run"]) <= 1e-4 * abs(b["run"]), (a["run"], b["run"])


# ---------------------------------------------------------------------------------------------------------------
# BASELINE configs[2] and [3] at full size: size-independent properties and the read-out metrics
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["c3", "c4"])
def test_full_size_properties_c3_c4(cuda_device, name):
    """c3 = DAVIS-shaped 90 frames 288x512 with per-frame-coherent ("Depth-Pro-conditioned") pointmap noise; c4 = Bonn/TUM-shaped
    110 frames 384x512 with the pose-metric read-out.  The loss is reproducible (only the order of the fp64 atomics may
    differ), a frame-window shard's accumulators add up to the unsharded ones (what the multi-GPU exchange relies on), and
    40 iterations of the canonical schedule end with sane depth (Sintel protocol) and pose (ATE / RPE) metrics."""
    from align3r_b200.synth import CONFIGS
    n, H, W, graph = CONFIGS[name]
    extra = dict(frame_noise=0.02, pixel_noise=0.003) if name == "c3" else {}
    scene = make_scene(n, H, W, graph, seed=2, with_flow=True, **extra)
    kw = dict(CANONICAL_KW, num_total_iter=40)
    net = build(scene, kw)
    net.verbose = False
    E = {"c3": 870, "c4": 1070}[name]
    assert net.n_edges == E and net._plan.pixel_slots() == 2 * E * H * W
    net.compute_global_alignment(init='mst', niter=0)
    l0, l0b = float(net(epoch=9999)), float(net(epoch=9999))
    assert math.isfinite(l0) and abs(l0 - l0b) <= 1e-6 * abs(l0)
    # additivity over frame windows: two shards, same parameters, accumulator blocks summed = the unsharded block
    full_acc = net._bufs.acc.clone()
    total = torch.zeros_like(full_acc)
    for r in range(2):
        sh = build(scene, kw, shard=dict(world=2, rank=r, group=None))
        sh._reduce_accumulators = lambda: None
        with torch.no_grad():
            for k in ("pw_poses", "im_poses", "im_focals"):
                getattr(sh, k).data.copy_(getattr(net, k).data)
            sh.im_depthmaps.data.copy_(net.im_depthmaps.data[torch.as_tensor(sh._plan.lframe_id.tolist(), device=net.device)])
        sh._phase_accumulate(_lib.M_FLOW_ACTIVE)
        total += sh._bufs.acc
        del sh
    scale = full_acc.abs().max()
    assert float((total - full_acc).abs().max()) <= 1e-6 * float(scale)
    final = net.compute_global_alignment(init=None, niter=40, schedule='linear', lr=0.01)
    hist = net._bufs.loss_hist[:40].cpu()
    assert math.isfinite(final) and bool(torch.isfinite(hist).all()) and not net.flow_loss_flag
    depth = torch.stack(net.get_depthmaps()).detach()
    assert bool(torch.isfinite(depth).all()) and float(depth.min()) > 0
    gt_depth = torch.stack([scene.gt.depth(k) for k in range(n)]).float()
    dm = go.depth_metrics(depth.cpu().numpy()[::10], gt_depth.numpy()[::10])
    pm = go.pose_metrics(net.get_im_poses().detach().cpu().numpy(), scene.gt.cam2world)
    print(name, "after 40 iterations:", dm, pm)
    assert dm["abs_rel"] < 0.1 and dm["d1"] > 0.9, dm
    assert pm["ate"] < 0.3 and pm["rpe_rot"] < 2.0, pm


def test_edge_score_map_and_tree_file(cuda_device, tmp_path):
    """compute_global_alignment(init='mst', save_score_path=..., save_score_only=...) (reference base_opt.py:460-466,
    init_im_poses.py:94-97,185-190,229-250): the score matrix equals the oracle's edge scores, the tree file lists the
    spanning tree in the oracle's order with the reference's star notation, and saving does not change the initialisation."""
    fx, scene, prob = golden_case("small_flow")
    ref = go.edge_scores(prob)
    net = build(scene, fx["kw"])
    before = net.im_poses.detach().clone()
    path = str(tmp_path / "scores.png")
    assert net.compute_global_alignment(init='mst', save_score_path=path, save_score_only=True) is None
    assert torch.equal(net.im_poses.detach(), before)                  # scores only: nothing initialised
    m = np.load(str(tmp_path / "scores.npy"))
    assert int((~np.isnan(m)).sum()) == len(ref)
    for (i, j), s in ref.items():
        assert abs(m[i, j] - s) <= 1e-5 * abs(s), ((i, j), m[i, j], s)

    init_fun.init_minimum_spanning_tree(net, save_score_path=path, niter_PnP=10)
    lines = open(str(tmp_path / "scores_tree.txt")).read().splitlines()
    tree = go.mst_init(prob)[1]["tree"]
    assert len(lines) == len(tree) == prob.n_imgs - 1
    assert lines[0].startswith(f"init edge ({tree[0][0]}*,{tree[0][1]}*) score=")
    for line, (i, j) in zip(lines[1:], tree[1:]):
        assert line.startswith(f"init edge ({i}*,{j}) score=") or line.startswith(f"init edge ({i},{j}*) score="), line
    plain = build(scene, fx["kw"])
    init_fun.init_minimum_spanning_tree(plain, niter_PnP=10)
    assert float((plain.pw_poses.detach() - net.pw_poses.detach()).abs().max()) < 1e-5


@pytest.mark.parametrize("conf", ["sqrt", "m1", "id"])
def test_confidence_transforms_against_live_oracle(cuda_device, conf):
    """conf='sqrt' | 'm1' | 'id' (reference commons.py:53-64): the pack kernel's weight, the raw confidence the Procrustes
    kernel recovers from it (MST initialisation against the oracle's) and the teacher-forced loss / gradients.  The oracle is
    pinned on these switches against the real reference (test_oracle_live.py::test_variant_switches_teacher_forced)."""
    scene = make_scene(6, 48, 64, "swin-2-noncyclic", seed=23, with_flow=True)
    kw = dict(CANONICAL_KW, num_total_iter=10, flow_loss_thre=0, conf=conf)
    prob = go.problem_from_output(scene.output, scene.flow_ij, scene.flow_ji, **kw)
    params, _ = go.mst_init(prob)
    net = build(scene, kw)
    init_fun.init_minimum_spanning_tree(net, niter_PnP=10)
    T_ref = go.pw_pose_matrices(prob, params["pw_poses"])
    assert float((net.get_pw_poses().detach().cpu() - T_ref).abs().max()) < 5e-5
    p64 = {k: v.double() for k, v in params.items()}
    names = list(go.TRAINABLE)
    loss64, g64, _ = go.loss_and_grads(prob.to(torch.float64), p64, 9999, trainable=names)
    _, g32, _ = go.loss_and_grads(prob, params, 9999, trainable=names)
    set_params(net, params)
    loss = net(epoch=9999)
    loss.backward()
    assert abs(float(loss.detach()) - float(loss64)) <= 2e-6 * abs(float(loss64))
    for k in names:
        m = getattr(net, k).grad.detach().cpu()
        m = m[:, :prob.P] if k == "im_depthmaps" else m.reshape(g64[k].shape)
        ok, err, ref_err = grad_close(k, m, g32[k], g64[k])
        assert ok, (k, err, ref_err)


def test_preset_frozen_focals_like_use_gt_focal(cuda_device):
    """tool/pose_test.py:181-193 (--use_gt_focal): per-frame focals preset and frozen, then the canonical init='mst' run.  The
    initialisation leaves the preset focals alone, the loop does not step them, and the run follows the oracle's (which is
    pinned on this flow against the real reference: test_oracle_live.py::test_preset_frozen_focals_like_use_gt_focal)."""
    scene = make_scene(5, 32, 48, "swin-2-noncyclic", seed=83, with_flow=True)
    kw = dict(CANONICAL_KW, num_total_iter=10, shared_focal=False)
    gt_focals = [0.9 * 48 * (1 + 0.01 * n) for n in range(5)]
    net = build(scene, kw)
    net.preset_focal(gt_focals, requires_grad=False)
    assert not net.im_focals.requires_grad
    net.compute_global_alignment(init='mst', niter=0)
    want = torch.tensor(gt_focals)
    assert torch.allclose(net.get_focals().detach().cpu().reshape(-1), want, rtol=1e-6)

    prob = go.problem_from_output(scene.output, scene.flow_ij, scene.flow_ji, **kw)
    params, _ = go.mst_init(prob)
    params["im_focals"] = (prob.focal_break * torch.log(want)).reshape(params["im_focals"].shape).float()
    T_ref = go.pw_pose_matrices(prob, params["pw_poses"])
    assert float((net.get_pw_poses().detach().cpu() - T_ref).abs().max()) < 5e-5
    set_params(net, params)                                   # the same start for the loop
    trainable = [k for k in go.TRAINABLE if k != "im_focals"]
    _, losses, _ = go.run_alignment(prob, params, niter=10, schedule="linear", lr=0.01, trainable=trainable)
    loss = net.compute_global_alignment(init=None, niter=10, schedule="linear", lr=0.01)
    assert abs(loss - losses[-1]) <= 5e-5 * abs(losses[-1]), (loss, losses[-1])
    assert torch.equal(net.im_focals.detach().cpu().reshape(-1), params["im_focals"].reshape(-1))     # never stepped


def test_flow_validity_maps(cuda_device):
    """flow_valid_mask_i / _j (reference get_flow -> OccMask, goem_opt.py:575-618; stored, never read by the loss): evaluated
    by k_occ_mask from the packed flow planes on first access, against the oracle's restatement (pinned live on the reference's
    OccMask).  Noisy flows exercise both outcomes; a pixel may differ only inside a rounding band around the threshold / border."""
    scene = make_scene(5, 32, 48, "swin-2-noncyclic", seed=91, with_flow=True)
    g = torch.Generator().manual_seed(5)
    scene.flow_ij = scene.flow_ij + 6 * torch.randn(scene.flow_ij.shape, generator=g)
    scene.flow_ji = scene.flow_ji + 6 * torch.randn(scene.flow_ji.shape, generator=g)
    net = build(scene, dict(CANONICAL_KW, num_total_iter=10))
    H, W = 32, 48
    vv, uu = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    for mine, a, b in ((net.flow_valid_mask_i, scene.flow_ij, scene.flow_ji), (net.flow_valid_mask_j, scene.flow_ji, scene.flow_ij)):
        want, inc = go.occ_mask(a, b, th=3.0)
        mine = mine.cpu()
        assert mine.dtype == torch.bool and mine.shape == want.shape
        assert 0.02 < float(want.float().mean()) < 0.98
        x, y = uu + a[:, 0], vv + a[:, 1]
        border = torch.minimum(torch.minimum(x.abs(), (x - (W - 1)).abs()), torch.minimum(y.abs(), (y - (H - 1)).abs()))[:, None]
        band = ((inc - 3.0).abs() < 1e-3) | (border < 1e-4)
        assert bool(((mine == want) | band).all()), int(((mine != want) & ~band).sum())
    assert net.flow_valid_mask_i is net.flow_valid_mask_i                      # evaluated once


# ---------------------------------------------------------------------------------------------
# full-size teacher-forced parity (tests/golden/{c2,f200}_teacher_forced.pt, oracle/make_golden_big.py)
# ---------------------------------------------------------------------------------------------
DEPTH_GRAD_TOL = 1e-5          # north star: per-pixel depth gradients within 1e-5 relative


def full_size_case(name, device):
    """Fixture, regenerated scene (inputs by recipe, checksum-guarded), our aligner at the recipe's parameters."""
    import os
    from align3r_b200.synth import load_params, near_solution_params
    from tests.helpers import scene_checksum
    fx = load_golden(f"{name}_teacher_forced")
    scene = make_scene(**fx["recipe"], workers=os.cpu_count() or 1)
    assert scene.edges == [tuple(e) for e in fx["edges"]]
    assert abs(scene_checksum(scene) - fx["checksum"]) <= 1e-9 * abs(fx["checksum"]), "synthetic generator drifted"
    net = global_aligner(scene.output, device, mode=GlobalAlignerMode.PointCloudOptimizer, flow=(scene.flow_ij, scene.flow_ji), **fx["kw"])
    net.verbose = False
    small, logdepth = near_solution_params(scene.gt, scene.edges, seed=fx["param_recipe"]["seed"], shared_focal=fx["kw"]["shared_focal"])
    load_params(net, small, logdepth)
    for k in ("pw_poses", "im_poses", "im_focals"):          # the recipe gives the fixture's parameters on this machine too
        assert abs(float(getattr(net, k).double().sum()) - fx["param_checksum"][k]) <= 1e-6 * max(1.0, abs(fx["param_checksum"][k])), k
    P = scene.gt.H * scene.gt.W
    assert abs(float(net.im_depthmaps[:, :P].double().sum()) - fx["param_checksum"]["im_depthmaps"]) <= 1e-7 * abs(fx["param_checksum"]["im_depthmaps"])
    return fx, scene, net, P


def full_size_rows(fx, net, P, tag):
    """One teacher-forced evaluation: per tensor (error of ours, error of the fp32 reference) against the fp64 arbiter."""
    g64 = fx[f"tf64_{tag}"]
    g32 = fx.get(f"tf32_{tag}") or fx[f"tf32_oracle_{tag}"]       # the real reference's fp32 run where it fits (c2), else the oracle's
    net.flow_loss_flag = False
    for k in ("pw_poses", "im_depthmaps", "im_poses", "im_focals"):          # .backward() accumulates into .grad
        getattr(net, k).grad = None
    loss = net(epoch=g64["epoch"])
    loss.backward()
    rows = {"loss": (abs(float(loss.detach()) - g64["loss"]) / abs(g64["loss"]), abs(g32["loss"] - g64["loss"]) / abs(g64["loss"]))}
    for k in ("pw_poses", "im_poses", "im_focals"):
        mine = getattr(net, k).grad.detach().cpu().reshape(g64["grads"][k].shape)
        ok, err, ref_err = grad_close(k, mine, g32["grads"][k], g64["grads"][k])
        rows[k] = (err, ref_err, ok)
    gd = net.im_depthmaps.grad.detach()[:, :P]
    sample = gd.reshape(-1)[::fx["stride"]].cpu()
    rows["im_depthmaps"] = (rel_err(sample, g64["grads"]["depth_sample"]), rel_err(g32["grads"]["depth_sample"], g64["grads"]["depth_sample"]))
    rows["im_depthmaps (per-frame sums)"] = (rel_err(gd.double().sum(dim=1).cpu(), g64["grads"]["depth_frame_sum"]),
                                             rel_err(g32["grads"]["depth_frame_sum"], g64["grads"]["depth_frame_sum"]))
    rows["im_depthmaps (norm)"] = (abs(float(gd.double().norm()) - g64["grads"]["depth_norm"]) / g64["grads"]["depth_norm"],
                                   abs(g32["grads"]["depth_norm"] - g64["grads"]["depth_norm"]) / g64["grads"]["depth_norm"])
    return rows, bool(net.flow_loss_flag) == bool(g64["flow_dropped"])


@pytest.mark.parametrize("name", ["c2", "f200", "c3", "c4"])
def test_full_size_teacher_forced(cuda_device, name):
    """BASELINE configs[1] at full size (c2: 50 frames 288x512, 470 edges) against the REAL reference's fp32 forward +
    autograd and the fp64 arbiter; a 200-frame 288x512 sub-scene of the long-video workload (1970 edges), configs[2] (c3: 90
    frames, per-frame-coherent pointmap noise) and configs[3] (c4: 110 frames 384x512) against the edge-chunked oracle: loss 2e-6; small gradients by the arbiter criterion (tests/helpers.py::grad_close); per-pixel depth
    gradients within 1e-5 relative of the fp64 gradients (the north star's figure) on a 19 000-pixel strided sample, the
    per-frame sums and the norm of the whole plane."""
    fx, scene, net, P = full_size_case(name, cuda_device)
    for tag in ("pre_flow", "flow"):
        rows, gate_ok = full_size_rows(fx, net, P, tag)
        assert gate_ok
        assert rows["loss"][0] <= 2e-6, (tag, rows["loss"])
        for k in ("pw_poses", "im_poses", "im_focals"):
            assert rows[k][2], (tag, k, rows[k])
        assert rows["im_depthmaps"][0] <= DEPTH_GRAD_TOL, (tag, rows["im_depthmaps"])
        assert rows["im_depthmaps (norm)"][0] <= DEPTH_GRAD_TOL, (tag, rows["im_depthmaps (norm)"])


def test_full_size_mst_initialisation_c2(cuda_device):
    """The MST initialisation at full size (c2) against the real reference's: pair scales, cameras, focal, depth maps - the
    tolerances of test_mst_initialisation (the 12 frames the tree leaves to RANSAC-PnP inherit its sensitivity to fp32
    rounding: 2e-3 on their camera matrices)."""
    fx, scene, net, P = full_size_case("c2", cuda_device)
    ref = fx["mst_init"]
    net.compute_global_alignment(init="mst", niter=0)
    assert float((net.get_im_poses().detach().cpu() - ref["im_poses_4x4"]).abs().max()) < 2e-3
    assert rel_err(net.get_focals().detach().cpu(), ref["focals"]) < 1e-5
    assert rel_err(net.get_pw_scale().detach().cpu(), ref["pw_poses"][:, 7].exp() * (np.log(0.5) - ref["pw_poses"][:, 7].mean()).exp()) < 1e-5
    sample = net.im_depthmaps.detach()[:, :P].reshape(-1)[::fx["stride"]].cpu()
    d = (sample - ref["logdepth_sample"]).abs()
    assert float(d.median()) < 1e-5 and float(d.max()) < 2e-3
    loss0 = float(net(epoch=0).detach())
    assert abs(loss0 - ref["loss_pre_flow"]) <= 5e-4 * ref["loss_pre_flow"]
