"""TEST INFRASTRUCTURE - generates tests/golden/*.pt by running the REAL reference (CPU).

Run in the build container only (needs /root/reference):

    python oracle/make_golden.py

Each fixture stores the synthetic-scene recipe (so the inputs are regenerated bit-identically from
``align3r_b200.synth``; an input checksum guards against generator drift) and what the unmodified
reference produced on it: the MST-initialised parameters, teacher-forced loss + gradients (fp32 and
the fp64 arbiter of SURVEY.md 8(c)-4) before and after the flow term switches on, the parameters after
one Adam step, and the outcome of a short free-running ``compute_global_alignment``.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_harness as rh                      # noqa: E402
from align3r_b200.synth import make_scene, CANONICAL_KW   # noqa: E402

NAMES = ["pw_poses", "pw_adaptors", "im_depthmaps", "im_poses", "im_focals", "im_pp"]
TRAIN = ["pw_poses", "im_depthmaps", "im_poses", "im_focals"]

CASES = {
    # name: (scene recipe, optimiser kwargs overrides, niter)
    "small_flow": (dict(n_frames=6, H=48, W=64, scene_graph="swin-2-noncyclic", seed=0, with_flow=True),
                   dict(num_total_iter=40), 40),
    "small_perframe_focal": (dict(n_frames=5, H=40, W=56, scene_graph="swin-3-noncyclic", seed=3, with_flow=True),
                             dict(num_total_iter=30, shared_focal=False, flow_loss_thre=0, translation_weight=0.3), 30),
    "small_gate_trips": (dict(n_frames=5, H=40, W=56, scene_graph="swin-2-noncyclic", seed=5, with_flow=True,
                              flow_noise=30.0),
                         dict(num_total_iter=20, flow_loss_thre=5, pxl_thre=0), 20),
}


def checksum(scene):
    acc = 0.0
    for t in (scene.pred1["pts3d"], scene.pred2["pts3d_in_other_view"], scene.pred1["conf"], scene.pred2["conf"]):
        acc += float(t.double().sum())
    if scene.flow_ij is not None:
        acc += float(scene.flow_ij.double().sum()) + float(scene.flow_ji.double().sum())
    return acc


def params_of(ref):
    return {k: getattr(ref, k).detach().clone() for k in NAMES}


def load_params(ref, params):
    # the reference's own load_state_dict (base_opt.py:179-180) trips over its im_conf ParameterList
    with torch.no_grad():
        for k, v in params.items():
            getattr(ref, k).data.copy_(v)


def teacher_forced(ref, epoch):
    ref.zero_grad()
    with rh.quiet():
        loss = ref(epoch=epoch)
    loss.backward()
    grads = {k: getattr(ref, k).grad.detach().clone() for k in TRAIN}
    return loss.detach().clone(), grads, bool(ref.flow_loss_flag)


def main():
    rh.import_reference()
    import dust3r.cloud_opt_flow.init_im_poses as init_fun
    from dust3r.cloud_opt_flow.base_opt import global_alignment_iter
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    for name, (recipe, over, niter) in CASES.items():
        scene = make_scene(**recipe)
        kw = dict(CANONICAL_KW)
        kw.update(over)
        out = dict(recipe=recipe, kw=kw, niter=niter, checksum=checksum(scene), edges=scene.edges)

        ref = rh.build_reference_scene(scene, **kw)
        with rh.quiet():
            init_fun.init_minimum_spanning_tree(ref, niter_PnP=10)
        out["init_params"] = params_of(ref)

        start = int(kw["num_total_iter"] * kw["flow_loss_start_epoch"])
        for tag, epoch in (("pre_flow", 0), ("flow", start + 1)):
            ref.flow_loss_flag = False
            loss, grads, flag = teacher_forced(ref, epoch)
            out[f"tf32_{tag}"] = dict(epoch=epoch, loss=loss, grads=grads, flow_dropped=flag)

        # fp64 arbiter on a fresh copy
        ref64 = rh.build_reference_scene(scene, **kw)
        load_params(ref64, out["init_params"])
        prev = torch.get_default_dtype()
        torch.set_default_dtype(torch.float64)
        try:
            rh.to_double(ref64)
            for tag, epoch in (("pre_flow", 0), ("flow", start + 1)):
                ref64.flow_loss_flag = False
                loss, grads, flag = teacher_forced(ref64, epoch)
                out[f"tf64_{tag}"] = dict(epoch=epoch, loss=loss, grads=grads, flow_dropped=flag)
        finally:
            torch.set_default_dtype(prev)

        # one Adam step from the init params at epoch start+1 (flow on), lr of the linear schedule
        ref1 = rh.build_reference_scene(scene, **kw)
        load_params(ref1, out["init_params"])
        opt = torch.optim.Adam([p for p in ref1.parameters() if p.requires_grad], lr=0.01, betas=(0.9, 0.9))
        with rh.quiet():
            l1, lr1 = global_alignment_iter(ref1, start + 1, niter, 0.01, 1e-3, opt, "linear")
        out["adam_step"] = dict(epoch=start + 1, loss=l1, lr=lr1, params=params_of(ref1))

        # free-running canonical call
        ref2 = rh.build_reference_scene(scene, **kw)
        with rh.quiet():
            final = ref2.compute_global_alignment(init="mst", niter=niter, schedule="linear", lr=0.01)
        out["free_run"] = dict(loss=final, params=params_of(ref2), flow_dropped=bool(ref2.flow_loss_flag),
                               im_poses_4x4=ref2.get_im_poses().detach().clone(),
                               focals=ref2.get_focals().detach().clone())
        path = os.path.join(ROOT, "tests", "golden", f"{name}.pt")
        torch.save(out, path)
        print(name, "->", path, f"{os.path.getsize(path)/1024:.0f} KB", "final loss", final,
              "dropped", out["free_run"]["flow_dropped"])


if __name__ == "__main__":
    main()
