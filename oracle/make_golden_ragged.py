"""TEST INFRASTRUCTURE - generates tests/golden/small_ragged.pt by running the REAL reference (CPU).

    python oracle/make_golden_ragged.py          (build container only: needs /root/reference)

Variable image shapes (reference optimizer.py:66-92, 601-607: the alignment term pads every per-pixel tensor to the largest
area; the flow term needs uniform shapes, so it is off here).  Frames 1 and 3 of a 5-frame 40x56 scene are cropped to 32x56.
Stored: the MST-initialised parameters, teacher-forced loss + gradients in fp32 and fp64, and a 15-iteration free run.
"""
import copy
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_harness as rh                                   # noqa: E402
from align3r_b200.synth import make_scene, CANONICAL_KW   # noqa: E402
from make_golden import checksum, params_of, load_params, teacher_forced   # noqa: E402

RECIPE = dict(n_frames=5, H=40, W=56, scene_graph="swin-2-noncyclic", seed=13, with_flow=False)
CROPPED = {1: 32, 3: 32}          # frame -> rows kept
NITER = 15


def ragged_output(scene):
    """Per-edge lists instead of stacked tensors; the views of a cropped frame lose their last rows."""
    out = copy.deepcopy(scene.output)
    idx1, idx2 = out["view1"]["idx"], out["view2"]["idx"]

    def crop(t, frames, chan_first=False):
        res = []
        for e, n in enumerate(frames):
            h = CROPPED.get(int(n))
            x = t[e]
            if h is not None:
                x = x[:, :h] if chan_first else x[:h]
            res.append(x.clone())
        return res
    out["pred1"]["pts3d"] = crop(out["pred1"]["pts3d"], idx1)
    out["pred1"]["conf"] = crop(out["pred1"]["conf"], idx1)
    out["pred2"]["pts3d_in_other_view"] = crop(out["pred2"]["pts3d_in_other_view"], idx2)
    out["pred2"]["conf"] = crop(out["pred2"]["conf"], idx2)
    for v, idx in (("view1", idx1), ("view2", idx2)):
        for k in ("img", "dynamic_mask"):
            if k in out[v]:
                out[v][k] = crop(out[v][k], idx, chan_first=(k == "img"))
    return out


def build_ref(out, kw):
    global_aligner, Mode, _, _ = rh.import_reference()
    with rh.quiet():
        return global_aligner(copy.deepcopy(out), "cpu", mode=Mode.PointCloudOptimizer, **kw)


def main():
    rh.import_reference()
    import dust3r.cloud_opt_flow.init_im_poses as init_fun
    scene = make_scene(**RECIPE)
    out = ragged_output(scene)
    kw = dict(CANONICAL_KW, num_total_iter=NITER, flow_loss_weight=0.0)
    ref = build_ref(out, kw)
    with rh.quiet():
        init_fun.init_minimum_spanning_tree(ref, niter_PnP=10)
    res = dict(recipe=RECIPE, cropped=CROPPED, kw=kw, niter=NITER, checksum=checksum(scene), edges=scene.edges,
               imshapes=[tuple(s) for s in ref.imshapes], init_params=params_of(ref))
    loss, grads, _ = teacher_forced(ref, 9999)
    res["tf32"] = dict(loss=loss, grads=grads)
    ref64 = build_ref(out, kw)
    load_params(ref64, res["init_params"])
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        ref64.double()
        loss, grads, _ = teacher_forced(ref64, 9999)
        res["tf64"] = dict(loss=loss, grads=grads)
    finally:
        torch.set_default_dtype(prev)
    ref2 = build_ref(out, kw)
    with rh.quiet():
        final = ref2.compute_global_alignment(init="mst", niter=NITER, schedule="linear", lr=0.01)
    res["free_run"] = dict(loss=float(final), params=params_of(ref2), im_poses_4x4=ref2.get_im_poses().detach().clone(),
                           focals=ref2.get_focals().detach().clone(),
                           depths=[d.detach().clone() for d in ref2.get_depthmaps()])
    path = os.path.join(ROOT, "tests", "golden", "small_ragged.pt")
    torch.save(res, path)
    print("->", path, f"{os.path.getsize(path) / 1024:.0f} KB  shapes {res['imshapes']}  tf loss {float(res['tf32']['loss'])}  final {final}")


if __name__ == "__main__":
    main()
