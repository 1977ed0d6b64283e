"""Per-tensor parity table (profiles/r02_parity_table.md): for every teacher-forced fixture, the relative L2 error of OUR
gradients against the fp64 arbiter beside the error of the reference's own fp32 gradients against the same arbiter.
Runs on the GPU box:  python tools/parity_table.py > gpurun_out/r02_parity_table.md"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from tests.helpers import golden_case, grad_close, rel_err  # noqa: E402
from tests import test_gpu_parity as T  # noqa: E402

print("# Teacher-forced parity, per tensor\n")
print("Relative L2 error against the fp64 arbiter (the reference run in double precision; for the full-size cases the oracle's")
print("edge-chunked fp64 evaluation, which equals the un-chunked oracle to 1e-16).  `ours` = the CUDA path (fp32 per-pixel math, fp64")
print("accumulators and chains); `reference fp32` = the reference's own fp32 forward + autograd (for f200: the oracle's fp32")
print("evaluation - the reference's graph does not fit in host memory at that size).\n")
print("| fixture | evaluation | tensor | ours vs fp64 | reference fp32 vs fp64 | ours / reference |")
print("|---|---|---|---|---|---|")
for name in T.CASES:
    fx, scene, prob = golden_case(name)
    net = T.build(scene, fx["kw"])
    T.set_params(net, fx["init_params"])
    for tag in ("pre_flow", "flow"):
        g32, g64 = fx[f"tf32_{tag}"], fx[f"tf64_{tag}"]
        net.flow_loss_flag = False
        for k in ("pw_poses", "im_depthmaps", "im_poses", "im_focals"):
            if getattr(net, k).grad is not None:
                getattr(net, k).grad = None
        loss = net(epoch=g32["epoch"])
        loss.backward()
        e_l = abs(float(loss.detach()) - float(g64["loss"])) / abs(float(g64["loss"]))
        r_l = abs(float(g32["loss"]) - float(g64["loss"])) / abs(float(g64["loss"]))
        print(f"| {name} ({prob.n_imgs} frames {prob.H}x{prob.W}, {prob.n_edges} edges) | {tag} | loss | {e_l:.1e} | {r_l:.1e} | {e_l / max(r_l, 1e-30):.2f} |")
        mine = T.grads_of(net, prob.P)
        for k in mine:
            ok, err, ref = grad_close(k, mine[k], g32["grads"][k], g64["grads"][k])
            print(f"| | | {k} | {err:.1e} | {ref:.1e} | {err / max(ref, 1e-30):.2f} |")
    del net
for name in ("c2", "f200", "c3", "c4"):
    fx, scene, net, P = T.full_size_case(name, "cuda")
    for tag in ("pre_flow", "flow"):
        for k in ("pw_poses", "im_depthmaps", "im_poses", "im_focals"):
            getattr(net, k).grad = None
        rows, _ = T.full_size_rows(fx, net, P, tag)
        first = True
        for k, v in rows.items():
            head = f"{name} ({scene.gt.n_frames} frames {scene.gt.H}x{scene.gt.W}, {len(scene.edges)} edges)" if first else ""
            print(f"| {head} | {tag if first else ''} | {k} | {v[0]:.1e} | {v[1]:.1e} | {v[0] / max(v[1], 1e-30):.2f} |")
            first = False
    del net, scene
    torch.cuda.empty_cache()
