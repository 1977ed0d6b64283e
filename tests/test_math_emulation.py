"""The per-element math of the CUDA kernels (align3r_b200/csrc/ga_math.cuh), re-enacted serially on the
CPU by tests/native/ga_emul.cpp, against the reference fixtures and the autograd oracle.  This checks the
hand-derived backward, the parameter chains, the gate and the Adam arithmetic without a GPU; the GPU
tests (test_gpu_parity.py) then check the kernels' parallel orchestration of the same math."""
import pytest
import torch

from tests.helpers import (TRAIN_ALL, cpu_problem, emul_eval, go, golden_case, grad_close, rel_err, make_scene, CANONICAL_KW)

CASES = ["small_flow", "small_perframe_focal", "small_gate_trips"]
# gradient tolerance: tests/helpers.py grad_close (SURVEY.md section 7 hard part 1)


def _grads_ok(mine, g32, g64, P):
    for k in ("pw_poses", "im_depthmaps", "im_poses", "im_focals"):
        m = mine[k][:, :P] if k == "im_depthmaps" else mine[k].reshape(g64[k].shape)
        ok, err, ref = grad_close(k, m, g32[k], g64[k])
        assert ok, (k, err, ref)


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("tag,flow_form", [("pre_flow", 0), ("flow", 0), ("flow", 1), ("flow", 2)])
def test_loss_and_gradients_match_reference_fixtures(name, tag, flow_form):
    """flow_form 1 / 2: the flow pixel-slot as the product scan evaluates it (flow_px_scan, general and row-tile form)."""
    fx, scene, prob = golden_case(name)
    g32, g64 = fx[f"tf32_{tag}"], fx[f"tf64_{tag}"]
    flow_active = prob.flow_loss_weight > 0 and g32["epoch"] >= prob.num_total_iter * prob.flow_loss_start_epoch
    bufs = cpu_problem(prob, fx["init_params"])
    loss, grads, scal = emul_eval(bufs, flow_active, flow_form=flow_form)
    assert abs(loss - float(g64["loss"])) <= 1e-6 * abs(float(g64["loss"]))
    assert bool(scal[4]) == g32["flow_dropped"]
    _grads_ok(grads, g32["grads"], g64["grads"], prob.P)


@pytest.mark.parametrize("name", CASES)
def test_adam_step_matches_reference_fixture(name):
    fx, scene, prob = golden_case(name)
    st = fx["adam_step"]
    bufs = cpu_problem(prob, fx["init_params"])
    loss, _, _ = emul_eval(bufs, True, apply=True, lr=st["lr"], step=0)
    assert abs(loss - st["loss"]) <= 1e-6 * abs(st["loss"])
    lr = st["lr"]
    for k in go.TRAINABLE:
        mine = bufs.params[k]
        mine = mine[:, :prob.P] if k == "im_depthmaps" else mine.reshape(st["params"][k].shape)
        d = (mine - st["params"][k]).abs()
        # first Adam step = lr * sign(g) up to eps: entries whose gradient is fp32 noise may flip sign
        assert float(d.max()) <= 2.0 * lr + 1e-6, k
        assert float((d > 1e-5).float().mean()) < 0.02, (k, float((d > 1e-5).float().mean()))


def test_emulated_free_run_tracks_reference():
    """30 free-running emulated iterations end where the reference ends."""
    fx, scene, prob = golden_case("small_perframe_focal")
    niter = fx["niter"]
    bufs = cpu_problem(prob, fx["init_params"])
    last = None
    for it in range(niter):
        lr = go.lr_schedule("linear", it / niter, 0.01, 1e-3)
        active = it >= prob.num_total_iter * prob.flow_loss_start_epoch
        last, _, _ = emul_eval(bufs, active, apply=True, lr=lr, step=it)
    ref = fx["free_run"]
    assert abs(last - ref["loss"]) <= 2e-4 * abs(ref["loss"])
    d = (bufs.params["im_depthmaps"][:, :prob.P] - ref["params"]["im_depthmaps"]).abs()
    assert float(d.median()) < 2e-4


@pytest.mark.parametrize("variant", ["l2", "adaptors_pp", "no_norm_scale"])
def test_rare_options_against_autograd_oracle(variant):
    """dist='l2', trainable adaptors + principal points, un-normalised pair scales: vs the fp64 oracle."""
    scene = make_scene(5, 32, 40, "swin-2-noncyclic", seed=11, with_flow=True)
    kw = dict(CANONICAL_KW, num_total_iter=10, flow_loss_thre=0)
    if variant == "l2":
        kw["dist"] = "l2"
    prob = go.problem_from_output(scene.output, scene.flow_ij, scene.flow_ji, **kw)
    if variant == "no_norm_scale":
        prob.norm_pw_scale = False
    params, _ = go.mst_init(prob)
    g = torch.Generator().manual_seed(1)
    params["pw_adaptors"] = 0.3 * torch.randn(prob.n_edges, 2, generator=g)
    params["im_pp"] = 0.2 * torch.randn(prob.n_imgs, 2, generator=g)
    names = list(go.TRAINABLE) + (["pw_adaptors", "im_pp"] if variant == "adaptors_pp" else [])
    p64 = {k: v.double() for k, v in params.items()}
    loss64, g64, _ = go.loss_and_grads(prob.to(torch.float64), p64, 9999, trainable=names)
    _, g32, _ = go.loss_and_grads(prob, params, 9999, trainable=names)
    train = dict(TRAIN_ALL, pw_adaptors="pw_adaptors" in names, im_pp="im_pp" in names)
    bufs = cpu_problem(prob, params)
    loss, grads, _ = emul_eval(bufs, True, trainable=train)
    assert abs(loss - float(loss64)) <= 2e-6 * abs(float(loss64))
    for k in names:
        m = grads[k][:, :prob.P] if k == "im_depthmaps" else grads[k].reshape(g64[k].shape)
        ok, err, ref = grad_close(k, m, g32[k], g64[k])
        assert ok, (k, err, ref)


def test_zero_weight_and_masked_pixels_have_no_gradient():
    scene = make_scene(4, 24, 32, "swin-2-noncyclic", seed=2, with_flow=True)
    scene.pred1["conf"][:] = 1.0          # log conf = 0 everywhere
    scene.pred2["conf"][:] = 1.0
    for m in scene.dynamic_masks:
        m[:] = True                       # every pixel dynamic -> flow term has no support
    scene.view1["dynamic_mask"][:] = True
    scene.view2["dynamic_mask"][:] = True
    kw = dict(CANONICAL_KW, num_total_iter=10, temporal_smoothing_weight=0.0)
    prob = go.problem_from_output(scene.output, scene.flow_ij, scene.flow_ji, **kw)
    g = torch.Generator().manual_seed(0)
    E, N = prob.n_edges, prob.n_imgs
    params = dict(pw_poses=torch.randn(E, 8, generator=g), pw_adaptors=torch.zeros(E, 2),
                  im_depthmaps=0.1 * torch.randn(N, prob.P, generator=g), im_poses=torch.randn(N, 7, generator=g),
                  im_focals=torch.tensor([[20 * 3.4]]), im_pp=torch.zeros(N, 2))
    bufs = cpu_problem(prob, params)
    loss, grads, scal = emul_eval(bufs, False)
    assert loss == 0.0
    for k, v in grads.items():
        assert float(v.abs().max()) == 0.0, k


def test_flow_validity_kernel_arithmetic_restated():
    """k_occ_mask (csrc/motion_kernels.cu) restated operation by operation in numpy fp32 - the grid_sample coordinate round
    trip, tap order nw / ne / sw / se with zero padding, |(sx + fx) + (sy + fy)| < th and the bounds test - against the oracle's
    occ_mask (which equals the reference's OccMask, tests/test_oracle_live.py)."""
    import numpy as np
    import torch
    from tests.helpers import go, make_scene
    f32 = np.float32

    def kernel(fw, bw, th=3.0):
        _, H, W = fw.shape
        hx, hy = f32((W - 1) / 2.0), f32((H - 1) / 2.0)
        v, u = np.meshgrid(np.arange(H, dtype=f32), np.arange(W, dtype=f32), indexing="ij")
        x, y = (u + fw[0]).astype(f32), (v + fw[1]).astype(f32)
        inside = ~((x < 0) | (x > f32(W - 1)) | (y < 0) | (y > f32(H - 1)))
        ix = ((((x / hx).astype(f32) - f32(1)) + f32(1)) / f32(2)).astype(f32) * f32(W - 1)
        iy = ((((y / hy).astype(f32) - f32(1)) + f32(1)) / f32(2)).astype(f32) * f32(H - 1)
        fx0, fy0 = np.floor(ix), np.floor(iy)
        wx1, wy1, wx0, wy0 = ix - fx0, iy - fy0, (fx0 + 1) - ix, (fy0 + 1) - iy
        x0, y0 = fx0.astype(np.int64), fy0.astype(np.int64)
        sx, sy = np.zeros_like(x), np.zeros_like(x)
        for t, wg in enumerate((wx0 * wy0, wx1 * wy0, wx0 * wy1, wx1 * wy1)):
            xx, yy = x0 + (t & 1), y0 + (t >> 1)
            ok = (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)
            xs, ys = np.clip(xx, 0, W - 1), np.clip(yy, 0, H - 1)
            sx = sx + np.where(ok, bw[0][ys, xs] * wg, 0).astype(f32)
            sy = sy + np.where(ok, bw[1][ys, xs] * wg, 0).astype(f32)
        inc = np.abs((sx + fw[0]) + (sy + fw[1]))
        return (inc < th) & inside, inc

    sc = make_scene(5, 32, 48, "swin-2-noncyclic", seed=91, with_flow=True)
    g = torch.Generator().manual_seed(5)
    wij = sc.flow_ij + 6 * torch.randn(sc.flow_ij.shape, generator=g)
    wji = sc.flow_ji + 6 * torch.randn(sc.flow_ji.shape, generator=g)
    for a, b in ((sc.flow_ij, sc.flow_ji), (wij, wji)):
        want, inc = go.occ_mask(a, b)
        for e in range(a.shape[0]):
            m, i2 = kernel(a[e].numpy(), b[e].numpy())
            band = np.abs(inc[e, 0].numpy() - 3.0) < 1e-3
            assert ((m == want[e, 0].numpy()) | band).all()
            assert float(np.abs(i2 - inc[e, 0].numpy()).max()) < 1e-4
