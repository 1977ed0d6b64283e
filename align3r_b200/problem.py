"""Device buffers of one alignment problem and the A3rProblem descriptor handed to the C ABI.

PyTorch is used only to own memory (``torch.empty`` on the target device) - every tensor listed here
is filled and consumed by the kernels behind include/a3r_ga.h.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .plan import ShardPlan

PARAM_NAMES = ("pw_poses", "pw_adaptors", "im_depthmaps", "im_poses", "im_focals", "im_pp")


class ProblemBuffers:
    """Allocates every buffer of A3rProblem for a ShardPlan.

    ``params`` maps parameter name -> tensor (already on ``device``) whose storage the kernels update
    in place; the depth parameter holds only the local frames' rows.
    """

    def __init__(self, plan: ShardPlan, device, params: dict, hyper: dict, has_flow: bool, loss_hist_len: int = 4096,
                 has_prior: bool = False):
        self.plan = plan
        self.device = torch.device(device)
        dev = self.device
        N, E = plan.n_frames, plan.n_edges
        nl, nls, stride = plan.n_local_frames, plan.n_local_slots, plan.pix_stride
        self.hyper = dict(hyper)
        self.has_flow = bool(has_flow)
        self.has_prior = bool(has_prior)          # depth prior: masks, initial log-depth and 4 more sums per frame

        def i32(a):
            return torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to(dev)

        def zeros(*shape, dtype=torch.float32):
            return torch.zeros(*shape, dtype=dtype, device=dev)

        self.edge_ij = i32(plan.edges)
        self.frame_hw = i32(plan.frame_hw)
        self.lframe_id = i32(plan.lframe_id)
        self.wf_lframe = i32(plan.wf_lframe)
        self.wf_slot_begin = i32(plan.wf_slot_begin)
        self.wf_tile_begin = i32(plan.wf_tile_begin)
        self.wf_accum = i32(plan.wf_accum)
        self.lslot_gslot = i32(plan.lslot_gslot)
        self.cta_tile_begin = i32(plan.cta_tile_begin)

        # parameters (owned by the caller, e.g. nn.Parameters) + Adam moments + gradients: attach_params()
        self.params, self.moments, self.grads = None, None, None
        if params is not None:
            self.attach_params(params)

        self.frame_par = zeros(N, 16)
        self.edge_par = zeros(E, 16)
        self.slot_par = zeros(2 * E, 16)

        # packed constant inputs, per-pixel gradient planes and the upload staging area: ONE device allocation carved
        # into views (a cudaMalloc costs milliseconds on a B200 box; a job would otherwise pay a dozen of them)
        nls1, nl1 = max(nls, 1), max(nl, 1)
        self.stage_rows = 32                                    # rows (edge-sides) per staged upload chunk
        parts = [("predw", nls1 * stride * 4, torch.float32), ("flow", nls1 * stride * 2 if has_flow else 0, torch.float32),
                 ("g_align", nl1 * stride, torch.float32), ("g_flow_i", nl1 * stride if has_flow else 0, torch.float32),
                 ("g_flow_j", nl1 * stride if has_flow else 0, torch.float32),
                 ("stage", self.stage_rows * stride * 4, torch.float32),
                 ("logdepth_init", nl1 * stride if has_prior else 0, torch.float32),
                 ("static_mask", nl1 * stride if (has_flow or has_prior) else 0, torch.uint8)]
        offs, total = {}, 0
        for name, numel, dt in parts:
            offs[name] = total
            total += (numel * torch.empty((), dtype=dt).element_size() + 255) // 256 * 256
        self.arena = torch.empty(total, dtype=torch.uint8, device=dev)
        zero_to = offs["stage"]                                 # the staging area and what follows need no fill ...
        self.arena[:zero_to].zero_()

        def carve(name, dt, *shape):
            n = int(np.prod(shape)) * torch.empty((), dtype=dt).element_size()
            return self.arena[offs[name]:offs[name] + n].view(dt).view(*shape)

        self.predw = carve("predw", torch.float32, nls1, stride, 4)
        self.flow = carve("flow", torch.float32, nls1, stride, 2) if has_flow else None
        self.static_mask = carve("static_mask", torch.uint8, nl1, stride) if (has_flow or has_prior) else None
        if self.static_mask is not None:
            self.static_mask.zero_()                            # ... except the masks behind it
        self.logdepth_init = carve("logdepth_init", torch.float32, nl1, stride) if has_prior else None
        if has_prior:
            self.logdepth_init.zero_()
        self.g_align = carve("g_align", torch.float32, nl1, stride)
        self.g_flow_i = carve("g_flow_i", torch.float32, nl1, stride) if has_flow else self.g_align
        self.g_flow_j = carve("g_flow_j", torch.float32, nl1, stride) if has_flow else self.g_align
        self.stage = carve("stage", torch.float32, self.stage_rows * stride * 4)

        # fp64 reduction buffers: the accumulator block acc_slot | acc_frame | prior_acc | acc_glob is one contiguous
        # allocation, twice (the loop alternates the two blocks by iteration parity; `acc` is block 0)
        n_sf = 2 * E * _lib.SLOT_ACC + N * _lib.FRAME_ACC
        n_pr = N * _lib.PRIOR_ACC if has_prior else 0
        self.n_acc = n_sf + n_pr + _lib.GLOB_ACC
        both = zeros(2 * self.n_acc, dtype=torch.float64)
        self.acc = both[:self.n_acc]                       # what multi-GPU runs all-reduce
        self.acc_alt = both[self.n_acc:]
        self.acc_slot = self.acc[: 2 * E * _lib.SLOT_ACC].view(2 * E, _lib.SLOT_ACC)
        self.acc_frame = self.acc[2 * E * _lib.SLOT_ACC:n_sf].view(N, _lib.FRAME_ACC)
        self.prior_acc = self.acc[n_sf:n_sf + n_pr].view(N, _lib.PRIOR_ACC) if has_prior else None
        self.acc_glob = self.acc[n_sf + n_pr:]
        self.frame_grad = zeros(N, 16, dtype=torch.float64)
        self.edge_scratch = zeros(E, dtype=torch.float64)
        self.pair_scratch = zeros(2 * E + N, _lib.PAIR_ROW, dtype=torch.float64)
        self.cam_cache = zeros(N, _lib.CAM_ROW, dtype=torch.float64)
        inc_begin, inc_items = incidence_lists(plan.edges, N)
        self.inc_begin, self.inc_items = i32(inc_begin), i32(inc_items)
        self.sync_ctr = torch.zeros(4, dtype=torch.int32, device=dev)
        self.scal = zeros(_lib.SCAL, dtype=torch.float64)
        self.loss_hist = zeros(loss_hist_len)
        self._struct = None

    # -----------------------------------------------------------------------------------
    def attach_params(self, params: dict):
        """Parameter tensors may arrive after the input planes were allocated (the constructor of the optimiser starts the
        big host->device uploads first and draws the random initial parameters on the CPU while they run)."""
        plan, dev = self.plan, self.device
        N, E, nl, stride = plan.n_frames, plan.n_edges, plan.n_local_frames, plan.pix_stride
        self.params = params
        for k in PARAM_NAMES:
            t = params[k]
            assert t.device == dev and t.dtype == torch.float32 and t.is_contiguous(), k
        assert tuple(params["pw_poses"].shape) == (E, 8)
        assert tuple(params["pw_adaptors"].shape) == (E, 2)
        assert tuple(params["im_poses"].shape) == (N, 7)
        assert tuple(params["im_pp"].shape) == (N, 2)
        assert tuple(params["im_depthmaps"].shape) == (nl, stride), (params["im_depthmaps"].shape, nl, stride)
        assert params["im_focals"].numel() == (1 if self.hyper["shared_focal"] else N)
        self.moments = {k: (torch.zeros_like(params[k]), torch.zeros_like(params[k])) for k in PARAM_NAMES}
        self.grads = {k: torch.zeros_like(params[k]) for k in PARAM_NAMES if k != "im_depthmaps"}
        self._struct = None

    # -----------------------------------------------------------------------------------
    def flags(self, trainable: dict) -> int:
        h = self.hyper
        f = 0
        if h["shared_focal"]:
            f |= _lib.F_SHARED_FOCAL
        if h["norm_pw_scale"]:
            f |= _lib.F_NORM_PW_SCALE
        if h["dist"] == "l2":
            f |= _lib.F_DIST_L2
        if self.has_flow:
            f |= _lib.F_HAS_FLOW
        pl = self.plan
        if pl.n_local_frames > 0 and all(int(pl.frame_hw[n, 1]) % pl.tile_pixels == 0 for n in pl.lframe_id.tolist()):
            f |= _lib.F_ROW_TILES      # every tile lies in one image row: the scan specialised for it (csrc/ga_scan_v4.cuh)
        for name, bit in (("pw_poses", _lib.F_TRAIN_PW_POSES), ("im_depthmaps", _lib.F_TRAIN_DEPTH),
                          ("im_poses", _lib.F_TRAIN_IM_POSES), ("im_focals", _lib.F_TRAIN_FOCALS),
                          ("im_pp", _lib.F_TRAIN_PP), ("pw_adaptors", _lib.F_TRAIN_ADAPTORS)):
            if trainable.get(name, False):
                f |= bit
        return f

    def struct(self, trainable: dict, refresh: bool = True) -> _lib.A3rProblem:
        """Fill the descriptor (cheap; call again whenever flags / hyper-parameters change)."""
        pl, h = self.plan, self.hyper
        p = _lib.A3rProblem()
        p.n_frames, p.n_edges = pl.n_frames, pl.n_edges
        p.n_local_frames, p.n_local_slots = pl.n_local_frames, pl.n_local_slots
        p.n_work_frames, p.n_tiles, p.n_ctas = pl.n_work_frames, pl.n_tiles, pl.n_ctas
        p.tile_pixels = pl.tile_pixels
        p.flags = self.flags(trainable)
        p.pix_stride = pl.pix_stride
        p.focal_break, p.pw_break, p.base_scale = h["focal_break"], h["pw_break"], h["base_scale"]
        p.flow_loss_weight, p.flow_loss_thre, p.pxl_thre = h["flow_loss_weight"], h["flow_loss_thre"], h["pxl_thre"]
        p.temporal_smoothing_weight, p.translation_weight = h["temporal_smoothing_weight"], h["translation_weight"]
        p.inv_area_i, p.inv_area_j = 1.0 / pl.total_area_i, 1.0 / pl.total_area_j
        p.adam_beta1, p.adam_beta2, p.adam_eps = h.get("adam_betas", (0.9, 0.9)) + (h.get("adam_eps", 1e-8),)
        p.depth_regularize_weight = float(h.get("depth_regularize_weight", 0.0)) if self.has_prior else 0.0
        P = _lib.ptr
        p.edge_ij, p.frame_hw = P(self.edge_ij), P(self.frame_hw)
        for cname, pname in (("pw_poses", "pw_poses"), ("pw_adaptors", "pw_adaptors"), ("im_poses", "im_poses"),
                             ("im_focals", "im_focals"), ("im_pp", "im_pp")):
            setattr(p, cname, P(self.params[pname]))
            setattr(p, cname + "_m", P(self.moments[pname][0]))
            setattr(p, cname + "_v", P(self.moments[pname][1]))
            setattr(p, "g_" + cname, P(self.grads[pname]))
        p.frame_par, p.edge_par, p.slot_par = P(self.frame_par), P(self.edge_par), P(self.slot_par)
        p.lframe_id, p.wf_lframe = P(self.lframe_id), P(self.wf_lframe)
        p.wf_slot_begin, p.wf_tile_begin, p.wf_accum = P(self.wf_slot_begin), P(self.wf_tile_begin), P(self.wf_accum)
        p.lslot_gslot, p.cta_tile_begin = P(self.lslot_gslot), P(self.cta_tile_begin)
        p.predw, p.flow, p.static_mask = P(self.predw), P(self.flow), P(self.static_mask)
        p.logdepth = P(self.params["im_depthmaps"])
        p.logdepth_m, p.logdepth_v = P(self.moments["im_depthmaps"][0]), P(self.moments["im_depthmaps"][1])
        p.g_align, p.g_flow_i, p.g_flow_j = P(self.g_align), P(self.g_flow_i), P(self.g_flow_j)
        p.acc_slot, p.acc_frame = P(self.acc_slot), P(self.acc_frame)
        p.frame_grad, p.edge_scratch, p.scal = P(self.frame_grad), P(self.edge_scratch), P(self.scal)
        p.loss_hist = P(self.loss_hist)
        p.logdepth_init, p.prior_acc = P(self.logdepth_init), P(self.prior_acc)
        p.acc_glob, p.acc_alt = P(self.acc_glob), P(self.acc_alt)
        p.pair_scratch, p.cam_cache = P(self.pair_scratch), P(self.cam_cache)
        p.inc_begin, p.inc_items, p.sync_ctr = P(self.inc_begin), P(self.inc_items), P(self.sync_ctr)
        self._struct = p
        return p

    def reset_adam(self):
        """A fresh optimiser per compute_global_alignment call (reference base_opt.py:524)."""
        for m, v in self.moments.values():
            m.zero_()
            v.zero_()

    def ensure_loss_hist(self, n: int):
        if self.loss_hist.numel() < n:
            self.loss_hist = torch.zeros(n, dtype=torch.float32, device=self.device)


def incidence_lists(edges, n_frames):
    """Per frame, the rows of A3rProblem.pair_scratch that carry a gradient contribution to its camera, as 2*row + part
    (include/a3r_ga.h): flow slot gs = 2e+side contributes its source half to the frame it lives on and its target half to
    the pair's other frame; temporal pair n (row 2E+n) its source half to frame n+1 and its target half to frame n.
    Ascending row order inside a frame: the gather of the update kernel sums in this fixed order."""
    edges = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
    E = len(edges)
    rows = np.arange(2 * E, dtype=np.int64)
    src = edges[rows >> 1, rows & 1]
    tgt = edges[rows >> 1, 1 - (rows & 1)]
    frame = np.concatenate([src, tgt, np.arange(1, n_frames), np.arange(0, n_frames - 1)])
    code = np.concatenate([2 * rows, 2 * rows + 1, 2 * (2 * E + np.arange(n_frames - 1)), 2 * (2 * E + np.arange(n_frames - 1)) + 1])
    order = np.lexsort((code, frame))
    begin = np.zeros(n_frames + 1, dtype=np.int64)
    np.add.at(begin, frame + 1, 1)
    return np.cumsum(begin).astype(np.int32), code[order].astype(np.int32)


def adam_schedule(lrs, start_step: int = 0, betas=(0.9, 0.9), first_iter: int = 0):
    """Host-side Adam scalars of consecutive steps (torch/optim/adam.py:457-547, python-double math)."""
    steps = (_lib.A3rAdamStep * len(lrs))()
    b1, b2 = betas
    for k, lr in enumerate(lrs):
        t = start_step + k + 1
        steps[k].step_size = lr / (1 - b1 ** t)
        steps[k].bc2_sqrt = (1 - b2 ** t) ** 0.5
        steps[k].iter = first_iter + k
    return steps


def modes_array(modes):
    arr = (C.c_uint32 * len(modes))()
    for k, m in enumerate(modes):
        arr[k] = m
    return arr
