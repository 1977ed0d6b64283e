"""TEST INFRASTRUCTURE - generates tests/golden/small_depth_prior.pt by running the REAL reference (CPU).

    python oracle/make_golden_depth_prior.py          (build container only: needs /root/reference)

The depth prior (reference optimizer.py:546-552, goem_opt.py:15-36; depth_regularize_weight > 0 - no caller of the
reference switches it on) is zero with zero gradient right after the initialisation, so the fixture holds a free run
with the prior on and the teacher-forced loss + gradients (fp32 and the fp64 arbiter) at the parameters it ends on.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_harness as rh                                   # noqa: E402
from align3r_b200.synth import make_scene, CANONICAL_KW   # noqa: E402
from make_golden import checksum, load_params, params_of, teacher_forced   # noqa: E402

RECIPE = dict(n_frames=5, H=40, W=56, scene_graph="swin-2-noncyclic", seed=11, with_flow=True)
NITER = 25


def main():
    rh.import_reference()
    scene = make_scene(**RECIPE)
    kw = dict(CANONICAL_KW, num_total_iter=NITER, depth_regularize_weight=0.5)
    out = dict(recipe=RECIPE, kw=kw, niter=NITER, checksum=checksum(scene), edges=scene.edges)

    ref = rh.build_reference_scene(scene, **kw)
    with rh.quiet():
        final = ref.compute_global_alignment(init="mst", niter=NITER, schedule="linear", lr=0.01)
    out["free_run"] = dict(loss=final, params=params_of(ref), flow_dropped=bool(ref.flow_loss_flag),
                           init_depth=torch.stack([d.detach().clone() for d in ref.init_depthmap]))
    ref.flow_loss_flag = False
    loss, grads, flag = teacher_forced(ref, NITER)
    out["tf32_final"] = dict(epoch=NITER, loss=loss, grads=grads, flow_dropped=flag)

    ref64 = rh.build_reference_scene(scene, **kw)
    load_params(ref64, out["free_run"]["params"])
    ref64.init_depthmap = [d.double() for d in out["free_run"]["init_depth"]]
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        rh.to_double(ref64)
        ref64.flow_loss_flag = False
        loss, grads, flag = teacher_forced(ref64, NITER)
        out["tf64_final"] = dict(epoch=NITER, loss=loss, grads=grads, flow_dropped=flag)
    finally:
        torch.set_default_dtype(prev)
    path = os.path.join(ROOT, "tests", "golden", "small_depth_prior.pt")
    torch.save(out, path)
    print("->", path, f"{os.path.getsize(path) / 1024:.0f} KB", "final loss", final, "tf32", float(out["tf32_final"]["loss"]),
          "tf64", float(out["tf64_final"]["loss"]))


if __name__ == "__main__":
    main()
