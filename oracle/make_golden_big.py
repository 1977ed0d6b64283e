"""TEST INFRASTRUCTURE - full-size teacher-forced fixtures (build container only: needs /root/reference).

    python oracle/make_golden_big.py [c2] [f200]

``c2``   BASELINE configs[1] at FULL size (50 frames 288x512, swin-5, 470 directed edges, flow + temporal): the REAL
         reference's forward + autograd backward in fp32 (28 GB of autograd graph on the host), flow term on and off, at
         parameters built by recipe (``align3r_b200.synth.near_solution_params``: a 29 MB depth stack does not fit a
         committed fixture, the recipe does), plus the fp64 arbiter.  The reference's own fp64 run would need ~56 GB, so
         the arbiter is the oracle's edge-chunked evaluation in fp64 (``ga_oracle.chunked_loss_and_grads``: equal to the
         un-chunked oracle to 1e-16 in fp64 / 1e-7 in fp32, tests/test_oracle_golden.py; checked here against the real
         reference's fp32 result at full size, the agreement is stored in the fixture).
``f200`` a 200-frame 288x512 swin-5 sub-scene of the long-video workload (1970 directed edges, 290 M pixel-edges): the
         edge-chunked oracle in fp32 and fp64 (the reference's graph would need > 100 GB).
``c3``, ``c4``  BASELINE configs[2] and [3] at full size (90 frames 288x512 with per-frame-coherent noise; 110 frames 384x512):
         the edge-chunked oracle in fp32 and fp64.

Stored per case: loss, every small gradient in full, and of the depth gradient (N x 147456 floats) a strided sample plus
per-frame sums / sums of squares; the recipe regenerates inputs and parameters bit for bit (input checksum included).
Also the reference's own MST initialisation of c2 (small parameters in full, strided log-depth sample) for the GPU
initialisation test at full size.
"""
import os
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_harness as rh                                   # noqa: E402
import ga_oracle as go                                     # noqa: E402
from align3r_b200.synth import make_scene, near_solution_params, CANONICAL_KW   # noqa: E402
from make_golden import checksum, load_params, teacher_forced                    # noqa: E402

STRIDES = {"c2": 389, "f200": 1553, "c3": 701, "c4": 1151}      # ~19 000 depth-gradient samples per case; co-prime with the image width: the sample
                                         # walks over rows and columns
TRAIN = ("pw_poses", "im_depthmaps", "im_poses", "im_focals")
CASES = {"c2": dict(n_frames=50, H=288, W=512, scene_graph="swin-5-noncyclic", seed=0, with_flow=True),
         "f200": dict(n_frames=200, H=288, W=512, scene_graph="swin-5-noncyclic", seed=0, with_flow=True),
         # BASELINE configs[2] (DAVIS-shaped, per-frame-coherent "Depth-Pro-conditioned" pointmap noise) and configs[3] (Bonn/TUM-shaped)
         "c3": dict(n_frames=90, H=288, W=512, scene_graph="swin-5-noncyclic", seed=2, with_flow=True, frame_noise=0.02, pixel_noise=0.003),
         "c4": dict(n_frames=110, H=384, W=512, scene_graph="swin-5-noncyclic", seed=2, with_flow=True)}


def pack_grads(grads, stride):
    """Small gradients in full; the depth gradient as a strided sample + per-frame moments."""
    out = {k: grads[k].detach().clone() for k in TRAIN if k != "im_depthmaps"}
    gd = grads["im_depthmaps"].detach()
    out["depth_sample"] = gd.reshape(-1)[::stride].clone()
    out["depth_frame_sum"] = gd.double().sum(dim=1)
    out["depth_frame_sq"] = gd.double().square().sum(dim=1)
    out["depth_norm"] = float(gd.double().norm())
    return out


def rel(a, b):
    a, b = a.double().reshape(-1), b.double().reshape(-1)
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


def recipe_params(scene, kw):
    small, logdepth = near_solution_params(scene.gt, scene.edges, seed=0, base_scale=0.5, focal_break=20.0, shared_focal=kw["shared_focal"])
    N, E = scene.gt.n_frames, len(scene.edges)
    return dict(pw_poses=small["pw_poses"], pw_adaptors=torch.zeros(E, 2), im_poses=small["im_poses"], im_focals=small["im_focals"],
                im_pp=torch.zeros(N, 2), im_depthmaps=torch.stack([logdepth(n) for n in range(N)]))


def main():
    which = [a for a in sys.argv[1:] if a in CASES] or list(CASES)
    for name in which:
        recipe = CASES[name]
        STRIDE = STRIDES[name]
        t0 = time.time()
        scene = make_scene(**recipe, workers=os.cpu_count() or 1)
        kw = dict(CANONICAL_KW, num_total_iter=300)
        out = dict(recipe=recipe, kw=kw, checksum=checksum(scene), edges=scene.edges, stride=STRIDE, param_recipe=dict(seed=0))
        prob = go.problem_from_output(scene.output, scene.flow_ij, scene.flow_ji, **kw)
        params = recipe_params(scene, kw)
        out["param_checksum"] = {k: float(v.double().sum()) for k, v in params.items()}
        print(name, "scene + parameters", f"{time.time() - t0:.0f} s", flush=True)

        for tag, epoch in (("flow", 9999), ("pre_flow", 0)):
            for dt, key in ((torch.float64, "tf64"), (torch.float32, "tf32_oracle")):
                t0 = time.time()
                loss, grads, info = go.chunked_loss_and_grads(prob.to(dt), {k: v.to(dt) for k, v in params.items()}, epoch, chunk=40)
                out[f"{key}_{tag}"] = dict(epoch=epoch, loss=float(loss), grads=pack_grads(grads, STRIDE), flow_dropped=info["flow_dropped"],
                                           flow_loss=info["flow_loss"])
                print(name, key, tag, "loss", float(loss), f"{time.time() - t0:.0f} s", flush=True)
                if dt == torch.float64:
                    g64 = grads
                else:
                    out[f"{key}_{tag}"]["err_vs_fp64"] = {k: rel(grads[k], g64[k]) for k in TRAIN}

        if name == "c2":            # the real reference at full size (fp32): teacher-forced, then its own MST initialisation
            rh.import_reference()
            ref = rh.build_reference_scene(scene, **kw)
            load_params(ref, {k: (v.reshape(getattr(ref, k).shape) if k != "im_depthmaps" else v) for k, v in params.items()})
            for tag, epoch in (("flow", 9999), ("pre_flow", 0)):
                t0 = time.time()
                ref.flow_loss_flag = False
                loss, grads, flag = teacher_forced(ref, epoch)
                rec = dict(epoch=epoch, loss=float(loss), grads=pack_grads(grads, STRIDE), flow_dropped=flag)
                o64, o32 = out[f"tf64_{tag}"], out[f"tf32_oracle_{tag}"]
                # the reference's own fp32 error against the arbiter, and the oracle's fp32 evaluation against the reference's
                rec["err_vs_fp64"] = {k: rel(rec["grads"][k], o64["grads"][k]) for k in ("pw_poses", "im_poses", "im_focals")}
                rec["err_vs_fp64"]["im_depthmaps"] = rel(rec["grads"]["depth_sample"], o64["grads"]["depth_sample"])
                rec["oracle32_vs_reference32"] = dict(loss=abs(o32["loss"] - rec["loss"]) / abs(rec["loss"]),
                                                      **{k: rel(o32["grads"][k], rec["grads"][k]) for k in ("pw_poses", "im_poses", "im_focals")},
                                                      im_depthmaps=rel(o32["grads"]["depth_sample"], rec["grads"]["depth_sample"]))
                out[f"tf32_{tag}"] = rec
                print(name, "reference fp32", tag, "loss", float(loss), "vs fp64 oracle", abs(float(loss) - o64["loss"]) / o64["loss"],
                      rec["err_vs_fp64"], rec["oracle32_vs_reference32"], f"{time.time() - t0:.0f} s", flush=True)
            del ref
            ref = rh.build_reference_scene(scene, **kw)
            t0 = time.time()
            with rh.quiet():
                ref.compute_global_alignment(init="mst", niter=0)
            out["mst_init"] = dict(pw_poses=ref.pw_poses.detach().clone(), im_poses=ref.im_poses.detach().clone(),
                                   im_focals=ref.im_focals.detach().clone(),
                                   logdepth_sample=torch.stack([d.detach().reshape(-1) for d in ref.im_depthmaps]).reshape(-1)[::STRIDE].clone(),
                                   im_poses_4x4=ref.get_im_poses().detach().clone(), focals=ref.get_focals().detach().clone())
            with rh.quiet():
                out["mst_init"]["loss_pre_flow"] = float(ref(epoch=0))
            print(name, "reference MST init", f"{time.time() - t0:.0f} s  loss", out["mst_init"]["loss_pre_flow"], flush=True)
            del ref
        path = os.path.join(ROOT, "tests", "golden", f"{name}_teacher_forced.pt")
        torch.save(out, path)
        print("->", path, f"{os.path.getsize(path) / 1024:.0f} KB", flush=True)


if __name__ == "__main__":
    main()
