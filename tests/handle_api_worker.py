"""Worker of tests/test_handle_api.py (run in its own process on the GPU box): the handle layer of the C ABI
(a3r_ga_create / pack_inputs / forward / grads / reset_adam / step / mst_edge_scores / destroy, include/a3r_ga.h) against the
stateless entry points the Python class drives, on the `small_flow` fixture.  Exits 0 when everything agrees."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch                                                              # noqa: E402

from tests.helpers import golden_case                                     # noqa: E402
from tests.test_gpu_parity import build, set_params                       # noqa: E402
from align3r_b200 import _lib                                             # noqa: E402
from align3r_b200.cloud_opt_flow import init_im_poses as init_fun         # noqa: E402
from align3r_b200.cloud_opt_flow.commons import schedule_lr               # noqa: E402


def main():
    fx, scene, prob = golden_case("small_flow")
    lib = _lib.load()
    E, N, P = prob.n_edges, prob.n_imgs, prob.P
    net = build(scene, fx["kw"])
    set_params(net, fx["init_params"])
    stream = net._stream()
    epoch = fx["tf32_flow"]["epoch"]

    # ---- the class path: loss + gradients of the fixture parameters ----
    loss_cls = float(net._evaluate(epoch).double())
    torch.cuda.synchronize()
    g_cls = {k: v.clone() for k, v in net._last_grads().items()}
    flow_on = net._flow_active(epoch)
    assert flow_on, "the fixture epoch must have the flow term on"

    # ---- handle: create, forward, grads ----
    st = net._bufs.struct(net._trainable())
    h = C.c_void_p()
    _lib.check(lib.a3r_ga_create(C.byref(st), C.byref(h)), "a3r_ga_create")
    assert h.value
    back = lib.a3r_ga_problem(h).contents
    assert (back.n_frames, back.n_edges, back.pix_stride) == (N, E, net._plan.pix_stride)
    loss_buf = torch.zeros(1, dtype=torch.float64, device=net.device)
    _lib.check(lib.a3r_ga_forward(h, _lib.M_FLOW_ACTIVE, _lib.ptr(loss_buf), stream), "a3r_ga_forward")
    torch.cuda.synchronize()
    loss_h = float(loss_buf[0])
    assert abs(loss_h - float(net._bufs.scal[_lib.S_LOSS])) == 0.0
    assert abs(loss_h - float(fx["tf64_flow"]["loss"])) <= 2e-6 * abs(float(fx["tf64_flow"]["loss"])), (loss_h, fx["tf64_flow"]["loss"])
    assert abs(loss_h - loss_cls) <= 1e-6 * abs(loss_cls), (loss_h, loss_cls)       # fp32 read-out on the class path
    gr = _lib.A3rGrads()
    _lib.check(lib.a3r_ga_grads(h, C.byref(gr)), "a3r_ga_grads")
    b = net._bufs
    assert gr.pw_poses == b.grads["pw_poses"].data_ptr() and gr.im_poses == b.grads["im_poses"].data_ptr()
    assert gr.im_focals == b.grads["im_focals"].data_ptr() and gr.logdepth == b.g_align.data_ptr() and gr.scal == b.scal.data_ptr()
    g_h = {k: v.clone() for k, v in net._last_grads().items()}
    for k in g_cls:                                       # same kernels, same inputs: equal up to the fp64 atomics' order
        d = float((g_h[k] - g_cls[k]).abs().max())
        assert d <= 1e-6 * max(1e-12, float(g_cls[k].abs().max())), (k, d)

    # ---- edge scores ----
    idx = torch.tensor([net._lslot[g] for g in range(2 * E)], device=net.device)
    conf_g = net._conf_sum[idx].contiguous()
    scores = torch.zeros(E, dtype=torch.float64, device=net.device)
    _lib.check(lib.a3r_mst_edge_scores(_lib.ptr(conf_g), _lib.ptr(b.edge_ij), _lib.ptr(b.frame_hw), E, _lib.ptr(scores), stream),
               "a3r_mst_edge_scores")
    want = init_fun.compute_edge_scores(net)
    got = scores.cpu().tolist()
    for e, ij in enumerate(net.edges):
        assert got[e] == want[ij], (ij, got[e], want[ij])

    # ---- pack_inputs: packing the same rows again leaves the planes as they are and reproduces the confidence sums ----
    predw_before, flow_before = b.predw.clone(), b.flow.clone()
    sums = torch.zeros_like(net._conf_sum)
    for side, (pts, conf, fl) in enumerate(((scene.pred1["pts3d"], scene.pred1["conf"], scene.flow_ij),
                                             (scene.pred2["pts3d_in_other_view"], scene.pred2["conf"], scene.flow_ji))):
        slots, lfr = net._dest_tables(side)
        dp = pts.reshape(E, P, 3).contiguous().to(net.device)
        dc = conf.reshape(E, P).contiguous().to(net.device)
        df = fl.reshape(E, 2, P).contiguous().to(net.device)
        _lib.check(lib.a3r_ga_pack_inputs(h, _lib.ptr(dp), _lib.ptr(dc), _lib.ptr(df), E, P, _lib.ptr(slots), _lib.ptr(lfr),
                                          _lib.CONF_MODES[net.conf_mode], None, _lib.ptr(sums), stream), "a3r_ga_pack_inputs")
        torch.cuda.synchronize()
    assert torch.equal(b.predw, predw_before) and torch.equal(b.flow, flow_before)
    assert torch.allclose(sums, net._conf_sum, rtol=1e-12, atol=0)

    # ---- three steps through the handle against the class's loop from the same parameters ----
    niter = 3
    ref = build(scene, dict(fx["kw"], num_total_iter=niter))
    set_params(ref, fx["init_params"])
    loss_ref = ref.compute_global_alignment(init=None, niter=niter, schedule="linear", lr=0.01)
    set_params(net, fx["init_params"])
    net.num_total_iter = niter
    net._refresh_hyper()
    st = net._bufs.struct(net._trainable())
    _lib.check(lib.a3r_ga_destroy(h), "a3r_ga_destroy")
    h = C.c_void_p()
    _lib.check(lib.a3r_ga_create(C.byref(st), C.byref(h)), "a3r_ga_create")
    _lib.check(lib.a3r_ga_reset_adam(h, stream), "a3r_ga_reset_adam")
    for it in range(niter):
        lr = float(schedule_lr("linear", it, niter, 0.01, 1e-3))
        flags = _lib.M_FLOW_ACTIVE if ref._flow_active(it) else 0
        _lib.check(lib.a3r_ga_step(h, it, niter, lr, flags, stream), "a3r_ga_step")
    torch.cuda.synchronize()
    hist_h, hist_r = net._bufs.loss_hist[:niter].cpu(), ref._bufs.loss_hist[:niter].cpu()
    assert torch.allclose(hist_h, hist_r, rtol=1e-5, atol=0), (hist_h, hist_r)
    assert abs(float(hist_r[-1]) - loss_ref) <= 1e-5 * abs(loss_ref)
    for k in ("pw_poses", "im_poses", "im_focals", "im_depthmaps"):
        d = (getattr(net, k).detach() - getattr(ref, k).detach()).abs()
        # the criterion of test_adam_step: Adam's first steps are sign-like, an entry whose gradient is rounding noise may step the other way
        assert float(d.max()) <= niter * 2.0 * 0.01 + 1e-6 and float((d > 1e-5).float().mean()) < 0.02, (k, float(d.max()))
    # argument errors are reported, not crashed
    assert lib.a3r_ga_step(h, niter, niter, 0.01, 0, stream) == -1 and b"iteration" in lib.a3r_last_error()
    _lib.check(lib.a3r_ga_destroy(h), "a3r_ga_destroy")
    print("handle layer: ok")


if __name__ == "__main__":
    main()
