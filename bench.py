#!/usr/bin/env python
"""Benchmark of the global-alignment hot path (BASELINE.json: GA iterations/s on the 50-frame 288x512
swin-5 graph with flow + temporal terms; % of the HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2]

A "step" is one GA iteration (parameter maps + fused forward/backward scan + parameter chains + Adam).
The K timed steps are the K iterations of ``global_alignment_loop(niter=K, schedule='linear', lr=0.01)``
started from the MST initialisation, so the flow term switches on at iteration 0.1*K exactly as in the
reference's canonical 300-iteration run.  One JSON line is printed by rank 0.

N > 1 (launched by torchrun): the scene is cut into N contiguous frame windows; a rank owns the slots
living on its frames; the fp64 accumulator block of the small parameters is exchanged once per iteration inside our own
kernel over NVLink peer memory (A3R_EXCHANGE=nccl: one NCCL all-reduce instead); total work is fixed ("scaling": "strong").

``--impl reference`` times the UNMODIFIED reference (dust3r.cloud_opt_flow PointCloudOptimizer + global_alignment_iter,
imported from /root/reference or its verbatim git-ignored copy oracle/_ref made by oracle/build_ref.py) on the host cores,
on the same workload: its own MST initialisation, warm-up iterations, then as many of the K iterations as fit the run's
time budget (never fewer than 5; `steps` in its line is what was timed); rank 0 only.  The ``cpu_baseline`` leg of our arm
runs the same function with 1 warm-up + 5 timed iterations.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--long-video", default="auto", choices=["auto", "on", "off"],
                    help="attach a long_video sub-record (the 1000-frame workload c5, device-drawn inputs, 40 iterations) to the line; "
                         "auto: on 8 GPUs with the default workload")
    ap.add_argument("--synth", default="auto", choices=["auto", "host", "device"],
                    help="where the synthetic pair predictions are drawn: host (pinned memory, the default workload) or "
                         "device (rows drawn on the GPU while they are packed: the 1000-frame workload; no e2e leg)")
    ap.add_argument("--cpu-budget", type=float, default=75.0, help="seconds for the cpu_baseline leg (reference construction + MST init + iterations)")
    ap.add_argument("--ref-budget", type=float, default=240.0, help="seconds for the whole --impl reference run")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------------
def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json (of measured)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s of B200_PROFILING.md (of fallback)"


def recorded_traffic():
    """dram bytes per accumulate launch from the committed ncu capture, if any (profiles/traffic.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks + throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def workload_name(name, cfg):
    """One description of the workload for both arms (config.workload)."""
    from align3r_b200.synth import make_edge_list
    n, H, W, graph = cfg
    E = len(make_edge_list(n, graph))
    return (f"{name}: {n}-frame {H}x{W} {graph} PointCloudOptimizer alignment ({E} directed edges, {E * H * W} pixel-edges), "
            "shared focal, flow 0.01 + temporal 0.01, linear lr 0.01->1e-3, init=mst")


def algorithmic_bytes(E, N, P, flow):
    """SURVEY.md section 8(d).  Returns (bytes per whole iteration, bytes of the accumulate kernel alone)."""
    it = E * P * 32 + N * P * 24 + (E * P * 16 + N * P if flow else 0)
    k1 = E * P * 32 + N * P * 4 + (E * P * 16 + N * P if flow else 0)
    return it, k1


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own CPU implementation on the host cores
# ------------------------------------------------------------------------------------------------
def reference_rate(cfg, budget_s, steps, warmup, seed, scene=None):
    """Times the UNMODIFIED reference (dust3r/cloud_opt_flow: PointCloudOptimizer + global_alignment_iter,
    base_opt.py:550-583; imported from /root/reference or its verbatim copy oracle/_ref) on the host cores, on the SAME
    workload as our arm: MST initialisation, then `steps` iterations of the canonical loop (linear schedule, lr 0.01, flow
    term on from iteration 0.1*steps) after `warmup` untimed iterations (flow on; parameters and a fresh Adam restored
    afterwards).  `steps` is cut down to what fits `budget_s` (never below 5; the blend of align-only and flow-active
    iterations stays 10 % / 90 % because the reference is built with num_total_iter = the steps actually run).
    Falls back to the oracle port (kind "port") only when the reference cannot be imported."""
    import torch
    from align3r_b200.synth import make_scene, make_edge_list, CANONICAL_KW
    here = os.path.join(ROOT, "oracle")
    if here not in sys.path:
        sys.path.insert(0, here)
    import ref_harness as rh
    if not rh.reference_available():
        return oracle_rate(cfg, budget_s, min(steps, 3), min(warmup, 1), seed)
    n, H, W, graph = cfg
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    t_start = time.perf_counter()
    _, _, _, base_opt = rh.import_reference()
    if scene is None:
        scene = make_scene(n, H, W, graph, seed=seed, with_flow=True, workers=cores)
    E = len(make_edge_list(n, graph))

    def fresh(k_total):
        torch.manual_seed(seed)
        ref = rh.build_reference_scene(scene, **dict(CANONICAL_KW, num_total_iter=k_total))
        with rh.quiet():
            ref.compute_global_alignment(init="mst", niter=0)            # the reference's own MST initialisation, no iteration
        return ref

    ref = fresh(steps)
    t_init = time.perf_counter() - t_start
    params = [p for p in ref.parameters() if p.requires_grad]
    snap = [p.detach().clone() for p in params]
    opt = torch.optim.Adam(params, lr=0.01, betas=(0.9, 0.9))
    t_w = []
    for w in range(max(1, warmup)):
        t0 = time.perf_counter()
        with rh.quiet():
            base_opt.global_alignment_iter(ref, steps - 1, steps, 0.01, 1e-3, opt, "linear")       # flow on
        t_w.append(time.perf_counter() - t0)
        if w >= 1 and time.perf_counter() - t_start > 0.4 * budget_s:
            break
    with torch.no_grad():
        for p, v in zip(params, snap):
            p.copy_(v)
    left = budget_s - (time.perf_counter() - t_start)
    k_run = min(steps, int(max(5, left / max(min(t_w), 1e-3))))
    ref.num_total_iter = k_run                              # optimizer.py:520: flow on from 0.1 * num_total_iter
    ref.flow_loss_flag = False
    opt = torch.optim.Adam(params, lr=0.01, betas=(0.9, 0.9))          # a fresh Adam, like global_alignment_loop (base_opt.py:524)
    losses = []
    t0 = time.perf_counter()
    for it in range(k_run):
        with rh.quiet():
            loss, _ = base_opt.global_alignment_iter(ref, it, k_run, 0.01, 1e-3, opt, "linear")
        losses.append(loss)
    dt = time.perf_counter() - t0
    n_flow = sum(1 for it in range(k_run) if it >= k_run * CANONICAL_KW["flow_loss_start_epoch"])
    return dict(value=k_run / dt, unit="iterations/s", cores=cores, kind="reference", steps_run=k_run,
                sample=(f"the unmodified reference (dust3r.cloud_opt_flow PointCloudOptimizer, global_alignment_iter) from "
                        f"{os.path.relpath(rh.REFERENCE_ROOT, ROOT) if rh.REFERENCE_ROOT.startswith(ROOT) else rh.REFERENCE_ROOT} "
                        f"on the full workload ({n} frames {H}x{W}, {E} edges): its own MST init ({t_init:.1f} s incl. construction), "
                        f"{len(t_w)} warm-up + {k_run} timed iterations ({n_flow} with the flow term) in {dt:.1f} s on {cores} host threads; "
                        f"roma restated by oracle/roma_shim.py"),
                ms_per_step_sample=1e3 * dt / k_run, torch_threads=torch.get_num_threads(), final_loss=losses[-1],
                same_config=True)


def oracle_rate(cfg, budget_s, steps, warmup, seed):
    """Fallback of reference_rate when the reference tree is neither at /root/reference nor copied to oracle/_ref: times
    `steps` iterations (after `warmup`) of the oracle PORT (oracle/ga_oracle.py) on a sample of workload cfg sized to the
    budget.  Returns dict(value = iterations/s scaled to the full workload, sample description, ...)."""
    import torch
    from align3r_b200.synth import make_scene, make_edge_list, CANONICAL_KW
    from oracle import ga_oracle as go
    n, H, W, graph = cfg
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    full_pe = len(make_edge_list(n, graph)) * H * W
    # ~1.6e-7 s per pixel-edge per iteration on 8 cores with flow active (measured in the build container)
    per_pe = 1.6e-7 * 8 / max(cores, 1) ** 0.7
    ladder = [(10, H, W), (8, H, W), (7, H // 2, W // 2), (6, H // 2, W // 2), (6, H // 4, W // 4), (6, H // 8, W // 8)]
    total_steps = steps + warmup
    pick = ladder[-1]
    for ns, hs, ws in ladder:
        pe = len(make_edge_list(min(ns, n), graph)) * hs * ws
        if pe * per_pe * total_steps <= budget_s:
            pick = (ns, hs, ws)
            break
    ns, hs, ws = min(pick[0], n), pick[1], pick[2]
    scene = make_scene(ns, hs, ws, graph, seed=seed, with_flow=True)
    kw = dict(CANONICAL_KW, num_total_iter=steps)
    prob = go.problem_from_output(scene.output, scene.flow_ij, scene.flow_ji, **kw)
    params, _ = go.mst_init(prob)
    p = {k: v.detach().clone().requires_grad_(k in go.TRAINABLE) for k, v in params.items()}
    opt = torch.optim.Adam([p[k] for k in go.TRAINABLE], lr=0.01, betas=(0.9, 0.9))

    def one(it, niter):
        for g in opt.param_groups:
            g["lr"] = go.lr_schedule("linear", it / niter, 0.01, 1e-3)
        opt.zero_grad()
        loss, _ = go.total_loss(prob, p, it)
        loss.backward()
        opt.step()
        return float(loss)

    for w in range(warmup):
        one(steps - 1, steps)              # warm up with the flow term on
    t0 = time.perf_counter()
    for it in range(steps):
        one(it, steps)
    dt = time.perf_counter() - t0
    sample_pe = prob.n_edges * prob.P
    rate_sample = steps / dt
    return dict(value=rate_sample * sample_pe / full_pe, unit="iterations/s", cores=cores, kind="port", steps_run=steps,
                sample=(f"REFERENCE TREE UNAVAILABLE - {steps} iterations (flow on from 0.1*{steps}) of the oracle port on a {ns}-frame "
                        f"{hs}x{ws} {graph} sample ({prob.n_edges} edges, {sample_pe} pixel-edges, {rate_sample:.3f} it/s), scaled by "
                        f"pixel-edges {sample_pe}/{full_pe} to the full workload"),
                ms_per_step_sample=1e3 * dt / steps, torch_threads=torch.get_num_threads(), same_config=False)


def run_reference(args, cfg, name):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    r = reference_rate(cfg, args.ref_budget, args.steps, args.warmup, args.seed)
    n, H, W, graph = cfg
    line = {
        "impl": "reference", "metric": "ga_iterations_per_sec", "value": r["value"], "unit": "iterations/s",
        "n_gpus": args.gpus, "steps": r["steps_run"], "warmup": args.warmup, "ms_per_step": 1e3 / r["value"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(name, cfg), "same_config": r["same_config"],
                   "steps_requested": args.steps, "steps_run": r["steps_run"],
                   "note": "CPU run of the reference's own code on the host cores; the timed steps are cut to the run's time budget "
                           f"({args.ref_budget:.0f} s), `steps` is what was timed"},
        "cpu_baseline": {"value": r["value"], "unit": "iterations/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "final_loss": r.get("final_loss"), "seconds_total": time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args, cfg, name):
    import torch
    import torch.distributed as dist
    from align3r_b200 import _lib
    from align3r_b200.cloud_opt_flow import GlobalAlignerMode, global_aligner
    from align3r_b200.cloud_opt_flow.optimizer import global_alignment_loop
    from align3r_b200.plan import frame_costs, partition_frames
    from align3r_b200.problem import adam_schedule
    from align3r_b200.synth import CANONICAL_KW, make_edge_list, make_scene, make_scene_device
    import numpy as np
    import ctypes as C

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the alignment kernels have no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world != args.gpus:
        raise SystemExit(f"bench.py --gpus {args.gpus} must be launched with torchrun --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    n, H, W, graph = cfg
    P = H * W
    K, Wm = args.steps, args.warmup
    all_edges = make_edge_list(n, graph)
    E = len(all_edges)
    shard, rows = None, None
    if world > 1:
        wins = partition_frames(frame_costs(np.asarray(all_edges, dtype=np.int32), np.asarray([(H, W)] * n, dtype=np.int32)), world)
        lo, hi = wins[rank]
        rows = [e for e, (i, j) in enumerate(all_edges) if lo <= i < hi or lo <= j < hi]
        shard = dict(world=world, rank=rank, group=None)
    device_synth = args.synth == "device" or (args.synth == "auto" and n >= 400)
    t_synth = time.perf_counter()
    if device_synth:
        args.no_e2e = True                 # no host copy of the inputs exists: nothing to upload
        scene = make_scene_device(n, H, W, graph, seed=args.seed, device=dev,
                                  edges=None if rows is None else [all_edges[e] for e in rows])
    else:
        scene = make_scene(n, H, W, graph, seed=args.seed, with_flow=True, pin_memory=True, workers=max(1, (os.cpu_count() or 1) // world),
                           edges=None if rows is None else [all_edges[e] for e in rows])
    t_synth = time.perf_counter() - t_synth
    kw = dict(CANONICAL_KW, num_total_iter=K)
    if world > 1:                      # host threads of this rank: the box's cores are shared by the ranks
        torch.set_num_threads(max(1, (os.cpu_count() or 1) // world))

    def output_dict():
        out = scene.output
        if rows is not None:          # every rank names the whole graph, rows cover only the edges it touches
            out = dict(out)
            out["view1"] = dict(out["view1"], idx=[i for i, _ in all_edges])
            out["view2"] = dict(out["view2"], idx=[j for _, j in all_edges])
        return out

    def build():
        return global_aligner(output_dict(), dev, mode=GlobalAlignerMode.PointCloudOptimizer,
                              flow=(scene.flow_ij, scene.flow_ji), shard=shard, edge_rows=rows, **kw)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---------------- end-to-end leg: host buffers -> public API -> host results ----------------
    h2d = sum(t.numel() * t.element_size() for t in (scene.pred1["pts3d"], scene.pred2["pts3d_in_other_view"], scene.pred1["conf"],
                                                      scene.pred2["conf"], scene.flow_ij, scene.flow_ji))
    e2e = None
    net = None
    if not args.no_e2e:
        # untimed warm-up job on a tiny scene: loads the CUDA modules, cv2 / scipy code paths and NCCL once, so that the
        # timed job below measures the job itself and not process start-up
        tiny = make_scene(6, 32, 64, "swin-2-noncyclic", seed=1, with_flow=True)
        wnet = global_aligner(tiny.output, dev, mode=GlobalAlignerMode.PointCloudOptimizer, flow=(tiny.flow_ij, tiny.flow_ji),
                              **dict(CANONICAL_KW, num_total_iter=10))
        wnet.compute_global_alignment(init="mst", niter=10, schedule="linear", lr=0.01)
        _ = (wnet.get_depthmaps(raw=True).detach().cpu(), wnet.get_im_poses().detach().cpu(), wnet.get_focals().detach().cpu(),
             wnet._bufs.loss_hist[:10].cpu())
        del wnet, tiny, _
        # ... and one untimed job on the workload itself (3 iterations): first-touch page faults of the host-side temporaries,
        # the PnP thread pool and cv2's buffers at this size, the arena / staging allocations, the peer mappings - the timed job
        # measures a steady-state job, like the W warm-up steps do for the device-timed loop (every world size)
        wnet = build()
        wnet.compute_global_alignment(init="mst", niter=3, schedule="linear", lr=0.01)
        torch.cuda.synchronize(dev)
        wnet.close()
        del wnet
        if world > 1:
            warm = torch.zeros(1024, device=dev, dtype=torch.float64)
            dist.all_reduce(warm)
        # the caller's result buffers (pinned, like the inputs): allocated once, outside the job
        n_own = n if world == 1 else wins[rank][1] - wins[rank][0]
        res_depth = torch.empty((n_own, (P + 127) // 128 * 128), dtype=torch.float32, pin_memory=True)

        def e2e_job(Kj, keep):
            nonlocal res_depth
            kwj = dict(CANONICAL_KW, num_total_iter=Kj)
            # the previous job's optimiser is a reference cycle holding a 3.5 GB arena: collect it HERE, outside the timed region,
            # and keep the cyclic collector off during the job (a collection that frees the arena in the middle of the next
            # constructor costs 0.1-0.3 s at a random place: the run-to-run spread of round 1's e2e figure)
            import gc
            gc.collect()
            torch.cuda.synchronize(dev)
            gc.disable()
            barrier()
            t0 = time.perf_counter()
            netj = global_aligner(output_dict(), dev, mode=GlobalAlignerMode.PointCloudOptimizer,
                                  flow=(scene.flow_ij, scene.flow_ji), shard=shard, edge_rows=rows, **kwj)
            t_c = time.perf_counter()                      # no synchronisation between the phases: the uploads of the
            netj.compute_global_alignment(init="mst", niter=0)      # constructor drain while the initialisation runs
            t_i = time.perf_counter()
            loss_j = netj.compute_global_alignment(init=None, niter=Kj, schedule="linear", lr=0.01)    # returns float(loss): syncs
            t_l = time.perf_counter()
            dd = netj.get_depthmaps(raw=True).detach()                      # (frames, pix_stride) in one copy
            if res_depth.shape != dd.shape:
                res_depth = torch.empty(dd.shape, dtype=torch.float32, pin_memory=True)     # (plan padded differently)
            res_depth.copy_(dd, non_blocking=True)
            poses_h = netj.get_im_poses().detach().cpu()
            foc_h = netj.get_focals().detach().cpu()
            hist_h = netj._bufs.loss_hist[:Kj].cpu()
            torch.cuda.synchronize(dev)
            phases = {"construct_upload_pack_s": t_c - t0, "mst_init_s": t_i - t_c, "loop_s": t_l - t_i, "readback_s": time.perf_counter() - t_l,
                      "note": "host-side wall clock per call, no synchronisation in between (asynchronous uploads overlap the next call)"}
            barrier()
            t_j = time.perf_counter() - t0
            gc.enable()
            if world > 1:
                tt = torch.tensor([t_j], device=dev, dtype=torch.float64)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                t_j = float(tt)
                hh = torch.tensor([float(h2d)], device=dev, dtype=torch.float64)
                dist.all_reduce(hh)
                h2d_total = float(hh)
            else:
                h2d_total = float(h2d)
            d2h = res_depth.numel() * 4 + poses_h.numel() * 4 + foc_h.numel() * 4 + hist_h.numel() * 4
            rec = {"value": Kj / t_j, "unit": "iterations/s", "h2d_bytes_per_step": h2d_total / Kj, "d2h_bytes_per_step": d2h / Kj,
                   "iterations": Kj, "seconds_total": t_j, "phases": phases, "final_loss": loss_j}
            if not keep:
                netj.close()
                netj = None
            return rec, netj

        canon = None
        if K != 300:        # the reference's canonical job length beside the driver's K (first: the kept job below holds its arena)
            canon, _ = e2e_job(300, False)
        e2e, net = e2e_job(K, True)
        e2e["what"] = ("global_aligner(host dict) + compute_global_alignment(init='mst', niter=K) + depth/pose/focal/loss read-back, "
                       "whole job / K; inputs are uploaded once per job (they are constant across iterations); one untimed "
                       "3-iteration job on the same inputs runs first (process warm-up); the cyclic garbage collector runs before "
                       "the job and is off during it")
        if canon is not None:
            e2e["canonical_300_iterations"] = canon
    if net is None:
        net = build()
        net.compute_global_alignment(init="mst", niter=0)
    else:
        # restart from the MST initialisation for the device-timed run
        net.compute_global_alignment(init="mst", niter=0)
    snap = {k: getattr(net, k).detach().clone() for k in ("pw_poses", "im_depthmaps", "im_poses", "im_focals")}

    # ---------------- warm-up: both kernel variants, then restore the initial state ----------------
    verbose_was = net.verbose
    net.verbose = False
    if Wm > 0:
        net.num_total_iter = 2 * Wm          # flow on for the second half
        global_alignment_loop(net, lr=0.01, niter=2 * Wm, schedule="linear")
        net.num_total_iter = K
    with torch.no_grad():
        for k, v in snap.items():
            getattr(net, k).data.copy_(v)
    net.flow_loss_flag = False

    # ---------------- timed region: K iterations, device time, max over ranks ----------------
    # The K-step loop is timed exactly as the contract says (barrier + synchronize on both sides, CUDA events on the launching
    # stream, max over ranks).  When K steps take less than 0.5 s the loop is REPEATED from the restored initial state until
    # at least 1 s of device time has been spent (at most 400 loops), so that the clock sampler sees a loaded GPU for several
    # samples; value = K / median loop time, the spread is reported.
    def restore():
        with torch.no_grad():
            for k, v in snap.items():
                getattr(net, k).data.copy_(v)
        net.flow_loss_flag = False

    sampler = ClockSampler(local).start() if rank == 0 else None
    loops, final_loss, spent = [], None, 0.0
    while True:
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(torch.cuda.current_stream(dev))
        fl = global_alignment_loop(net, lr=0.01, niter=K, schedule="linear")
        ev1.record(torch.cuda.current_stream(dev))
        barrier()
        t_ms = ev0.elapsed_time(ev1)
        if world > 1:
            tt = torch.tensor([t_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t_ms = float(tt)
        if final_loss is None:
            final_loss = fl
            loss_head = [float(x) for x in net._bufs.loss_hist[:min(K, 24)].cpu()]
        loops.append(t_ms)
        spent += t_ms
        if loops[0] >= 500.0 or spent >= 1000.0 or len(loops) >= 400:        # the same decision on every rank (max over ranks)
            break
        restore()
    clocks = sampler.stop() if sampler is not None else None
    ms = sorted(loops)[len(loops) // 2]
    n_flow = sum(1 for it in range(K) if net._flow_active(it))

    # ---------------- roofline of the dominant kernel, timed alone ----------------
    # The loop's dominant kernel is the scan WITH the previous iteration's depth Adam step folded into its tile prologue
    # (every iteration but the first): it moves the whole iteration's algorithmic bytes (48 B per pixel-slot +
    # 25 B per pixel with flow) plus the parked gradient planes, which are overhead, not algorithmic.  The plain scan (the
    # first iteration of a loop, and forward()) is reported beside it.
    b = net._bufs
    st = b.struct(net._trainable())
    stream = net._stream()
    roof = {}
    reps = 20
    prev = adam_schedule([0.0], start_step=10, first_iter=-1)          # lr 0: the repeated depth steps leave the planes alone
    folds = True
    for label, mode, folded in (("flow", _lib.M_FLOW_ACTIVE, True), ("align", 0, True), ("flow_plain", _lib.M_FLOW_ACTIVE, False),
                                ("align_plain", 0, False)):
        def launch():
            if folded and folds:
                return net._lib.a3r_ga_accumulate_folded(C.byref(st), mode, prev, mode, stream)
            return net._lib.a3r_ga_accumulate(C.byref(st), mode, stream)
        _lib.check(net._lib.a3r_ga_prepare(C.byref(st), mode, stream))
        if folded and net._lib.a3r_ga_accumulate_folded(C.byref(st), mode, prev, mode, stream) != 0:
            folds = False                               # (split work frames / frozen depth: the loop runs the plain scan + K3)
        for _ in range(3):
            _lib.check(launch())
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(torch.cuda.current_stream(dev))
        for _ in range(reps):
            _lib.check(launch())
        a1.record(torch.cuda.current_stream(dev))
        torch.cuda.synchronize(dev)
        roof[label] = a0.elapsed_time(a1) / reps
    # per-rank share (pixel-slots are what a launch processes)
    local_ps = net._plan.pixel_slots()
    nl = net._plan.n_local_frames
    adam_b = 24 if folds else 0
    k1_bytes_flow = local_ps * 24 + nl * P * (5 + adam_b)
    k1_bytes_align = local_ps * 16 + nl * P * (4 + adam_b)
    plain_bytes_flow, plain_bytes_align = local_ps * 24 + nl * P * 5, local_ps * 16 + nl * P * 4
    peak, peak_src = measured_peak()
    ach = k1_bytes_flow / (roof["flow"] * 1e-3) / 1e9
    traffic = recorded_traffic()
    it_bytes_flow, _ = algorithmic_bytes(E, n, P, True)
    it_bytes_align, _ = algorithmic_bytes(E, n, P, False)
    blended_bytes = (n_flow * it_bytes_flow + (K - n_flow) * it_bytes_align) / K

    line = None
    if rank == 0:
        line = {
            "metric": "ga_iterations_per_sec", "value": K / (ms * 1e-3), "unit": "iterations/s", "n_gpus": world,
            "steps": K, "warmup": Wm, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(name, cfg),
                       "iterations_with_flow": n_flow, "parallelism": "single GPU" if world == 1 else f"{world} frame windows, "
                       + ("one NCCL all-reduce" if os.environ.get("A3R_EXCHANGE") == "nccl" else "one peer-memory exchange kernel (NVLink loads)")
                       + " of the fp64 accumulator block per iteration",
                       "l2": "inputs exceed L2 (per-iteration scan >> 126 MB)"},
            "pixel_edges_per_sec": E * P * K / (ms * 1e-3),
            "final_loss": final_loss, "loss_hist_head": loss_head,
            "timed_loops": {"count": len(loops), "ms_median": ms, "ms_min": min(loops), "ms_max": max(loops), "ms_first": loops[0],
                            "note": "K-step loops from the restored MST initialisation; value = K / median"},
            "synth": {"where": "device rows" if device_synth else "host, pinned", "seconds": t_synth},
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "traffic": (traffic or {}).get("accumulate_folded_flow_bytes_per_launch" if folds else "accumulate_flow_bytes_per_launch")
                         if world == 1 and name == "c2" else None,
                         "traffic_source": "recorded: profiles/traffic.json (ncu --set full of this kernel on this workload), not measured in this run",
                         "kernel": ("k_accumulate_v4<flow, fold> = the scan with the previous iteration's depth Adam step in its tile prologue"
                                    if folds else "k_accumulate (flow active)") + ", timed alone over %d launches" % reps,
                         "ms_per_launch": roof["flow"], "algorithmic_bytes_per_launch": k1_bytes_flow, "peak_source": peak_src,
                         "align_only": {"ms_per_launch": roof["align"], "achieved": k1_bytes_align / (roof["align"] * 1e-3) / 1e9,
                                        "frac": k1_bytes_align / (roof["align"] * 1e-3) / 1e9 / peak},
                         "plain_scan": {"what": "the scan without the folded depth step (first iteration of a loop, forward())",
                                        "flow": {"ms_per_launch": roof["flow_plain"], "frac": plain_bytes_flow / (roof["flow_plain"] * 1e-3) / 1e9 / peak},
                                        "align_only": {"ms_per_launch": roof["align_plain"], "frac": plain_bytes_align / (roof["align_plain"] * 1e-3) / 1e9 / peak}}},
            "roofline_iteration": {"bytes_per_iteration": blended_bytes, "achieved": blended_bytes / world / (ms / K * 1e-3) / 1e9,
                                   "frac": blended_bytes / (ms / K * 1e-3) / 1e9 / peak / world,
                                   "note": "whole iteration (all kernels), SURVEY 8(d) bytes, per GPU"},
            "gpu_launches": len(loops) * (2 * K + 2),      # per K-step loop: K scans + K update kernels + one k_prepare + the last depth step
            "clocks": clocks,
        }
        if e2e is not None:
            line["e2e"] = e2e
    # ---------------- long video (BASELINE configs[4]) beside it: the north star's scaling target ----------------
    want_long = args.long_video == "on" or (args.long_video == "auto" and world == 8 and name == "c2")
    if want_long:
        del net
        scene = None
        torch.cuda.empty_cache()
        lv = long_video_record(args, world, rank, dev, barrier)
        if rank == 0:
            line["long_video"] = lv
        net = None
    # ---------------- CPU baseline beside it (rank 0, N = 1 only) ----------------
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        net = None
        torch.cuda.empty_cache()
        if scene is None:
            scene = make_scene(n, H, W, graph, seed=args.seed, with_flow=True, workers=os.cpu_count() or 1)
        cb = reference_rate(cfg, args.cpu_budget, 5, 1, args.seed, scene=scene)
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        from align3r_b200.cloud_opt_flow.optimizer import free_peer_regions
        free_peer_regions()
        dist.barrier()
        dist.destroy_process_group()


def long_video_record(args, world, rank, dev, barrier, steps=40, warmup=3):
    """The 1000-frame workload (BASELINE configs[4], c5) on the same ranks: device-drawn inputs, MST initialisation,
    `steps` iterations of the canonical loop timed like the main line (CUDA events, barrier on both sides, max over ranks)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from align3r_b200.cloud_opt_flow import GlobalAlignerMode, global_aligner
    from align3r_b200.cloud_opt_flow.optimizer import global_alignment_loop
    from align3r_b200.plan import frame_costs, partition_frames
    from align3r_b200.synth import CANONICAL_KW, CONFIGS, make_edge_list, make_scene_device
    n, H, W, graph = CONFIGS["c5"]
    all_edges = make_edge_list(n, graph)
    rows, shard = None, None
    if world > 1:
        lo, hi = partition_frames(frame_costs(np.asarray(all_edges, dtype=np.int32), np.asarray([(H, W)] * n, dtype=np.int32)), world)[rank]
        rows = [e for e, (i, j) in enumerate(all_edges) if lo <= i < hi or lo <= j < hi]
        shard = dict(world=world, rank=rank, group=None)
    t0 = time.perf_counter()
    scene = make_scene_device(n, H, W, graph, seed=args.seed, device=dev, edges=None if rows is None else [all_edges[e] for e in rows])
    out = scene.output
    if rows is not None:
        out = dict(out, view1=dict(out["view1"], idx=[i for i, _ in all_edges]), view2=dict(out["view2"], idx=[j for _, j in all_edges]))
    net = global_aligner(out, dev, mode=GlobalAlignerMode.PointCloudOptimizer, flow=(scene.flow_ij, scene.flow_ji), shard=shard,
                         edge_rows=rows, **dict(CANONICAL_KW, num_total_iter=steps))
    net.verbose = False
    net.compute_global_alignment(init="mst", niter=0)
    snap = {k: getattr(net, k).detach().clone() for k in ("pw_poses", "im_depthmaps", "im_poses", "im_focals")}
    net.num_total_iter = 2 * warmup
    global_alignment_loop(net, lr=0.01, niter=2 * warmup, schedule="linear")
    net.num_total_iter = steps
    with torch.no_grad():
        for k, v in snap.items():
            getattr(net, k).data.copy_(v)
    net.flow_loss_flag = False
    torch.cuda.synchronize(dev)
    t_setup = time.perf_counter() - t0
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(torch.cuda.current_stream(dev))
    final = global_alignment_loop(net, lr=0.01, niter=steps, schedule="linear")
    e1.record(torch.cuda.current_stream(dev))
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        tt = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt)
    E, P = len(all_edges), H * W
    n_flow = sum(1 for it in range(steps) if net._flow_active(it))
    it_f, _ = algorithmic_bytes(E, n, P, True)
    it_a, _ = algorithmic_bytes(E, n, P, False)
    blended = (n_flow * it_f + (steps - n_flow) * it_a) / steps
    peak, _ = measured_peak()
    rec = {"workload": workload_name("c5", CONFIGS["c5"]), "value": steps / (ms * 1e-3), "unit": "iterations/s", "n_gpus": world,
           "steps": steps, "warmup": 2 * warmup, "ms_per_step": ms / steps, "pixel_edges_per_sec": E * P * steps / (ms * 1e-3),
           "final_loss": final, "loss_hist_head": [float(x) for x in net._bufs.loss_hist[:min(steps, 12)].cpu()],
           "roofline_iteration_frac_per_gpu": blended / (ms / steps * 1e-3) / 1e9 / peak / world,
           "setup_seconds": t_setup, "synth": "rows drawn on the GPUs while they are packed (no host copy, no e2e leg)",
           "one_gpu_recorded": {"value": 66.2, "source": "profiles/r02_c5_lines.jsonl (same workload and protocol on one B200, round 2, box power-capped at 1635 MHz)"}}
    del net
    torch.cuda.empty_cache()
    return rec


def main():
    args = parse_args()
    from align3r_b200.synth import CONFIGS
    cfg = CONFIGS[args.workload]
    if args.impl == "reference":
        run_reference(args, cfg, args.workload)
    else:
        run_ours(args, cfg, args.workload)


if __name__ == "__main__":
    main()
