"""PointCloudOptimizer - drop-in for reference dust3r/cloud_opt_flow/optimizer.py:30 (+ the
BasePCOptimizer part, base_opt.py:44) whose hot loop runs in the sm_100a kernels of liba3r_ga.so.

Same constructor keywords, same parameter names (``pw_poses, pw_adaptors, im_depthmaps, im_poses,
im_focals, im_pp``), same getters.  What differs by design:
  * inputs are packed once into frame-major float4 / float2 planes on the GPU (no ParameterDict
    duplicate, reference optimizer.py:80-92 keeps every input twice);
  * ``forward`` / ``compute_global_alignment`` never build an autograd graph - loss, hand-derived
    gradients and the Adam step come from the kernels (``forward`` still returns a tensor whose
    ``.backward()`` fills ``.grad`` of the parameters, so external optimisers keep working);
  * optical flow is not computed here: RAFT is out of scope and
    ``get_flow`` returns the tensors injected with ``flow=(flow_ij, flow_ji)``;
  * a CUDA device is mandatory - there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import numpy as np
import torch
import torch.nn as nn

from .. import _lib
from .._timing import Phases
from ..plan import build_plan, check_edges, frame_costs, partition_frames
from ..problem import ProblemBuffers, adam_schedule, modes_array
from .commons import (edge_str, get_conf_trf, get_conf_untrf, signed_expm1, signed_log1p, schedule_lr,
                      unitquat_to_rotmat, rotmat_to_unitquat)
from . import init_im_poses as init_fun

_TRAIN_ORDER = ("pw_poses", "im_depthmaps", "im_poses", "im_focals")   # reference base_opt.py:514 order

# exchange regions of frame-window sharded runs (a3r_peer_alloc + the peers' CUDA-IPC mappings), cached per
# (device, world, rank, accumulator-block size, group): see PointCloudOptimizer._setup_peers
_PEER_REGIONS = {}


def free_peer_regions():
    """Releases every cached peer-exchange region of this process.  COLLECTIVE on sharded runs (every rank calls it, with no
    alignment in flight): a barrier makes sure no peer still reads a region before it is unmapped and freed."""
    import torch.distributed as dist
    for key in list(_PEER_REGIONS):
        e = _PEER_REGIONS.pop(key)
        torch.cuda.synchronize(e['device'])
        if dist.is_available() and dist.is_initialized():
            dist.barrier(group=e['group'])
        with torch.cuda.device(e['device']):
            for ptr in e['opened']:
                e['lib'].a3r_peer_close(ptr)
            if e['region'].value:
                e['lib'].a3r_peer_free(e['region'])


def _row_hw(rows, r):
    """(H, W) of row r of a stack of pointmaps: one (E,H,W,3) array-like (a tensor, or rows drawn on demand) or a list of
    per-edge (H,W,3) tensors of varying shapes."""
    shp = getattr(rows, 'shape', None)
    if shp is not None and len(shp) == 4:
        return (int(shp[1]), int(shp[2]))
    return tuple(rows[r].shape[0:2])


def _is_stack(rows):
    shp = getattr(rows, 'shape', None)
    return shp is not None and hasattr(rows, '__getitem__') and not isinstance(rows, (list, tuple)) and len(shp) >= 3


class _SlotDict:
    """Read-only dict-like view ``'i_j' -> tensor`` over the packed slot planes (the reference keeps
    ParameterDicts pred_i / pred_j / conf_i / conf_j, base_opt.py:88-100)."""

    def __init__(self, scene, side, what):
        self._scene, self._side, self._what = scene, side, what

    def keys(self):
        return self._scene.str_edges

    def __contains__(self, k):
        return k in self._scene._edge_index

    def __len__(self):
        return self._scene.n_edges

    def __iter__(self):
        return iter(self.keys())

    def items(self):
        return ((k, self[k]) for k in self.keys())

    def __getitem__(self, key):
        sc = self._scene
        e = sc._edge_index[key]
        gs = 2 * e + self._side
        if gs not in sc._lslot:
            raise KeyError(f"slot of edge {key} lives on another GPU")
        n = sc.edges[e][self._side]
        h, w = sc.imshapes[n]
        plane = sc._bufs.predw[sc._lslot[gs], :h * w]
        if self._what == "pts":
            return plane[:, :3].view(h, w, 3)
        return sc._conf_untrf(plane[:, 3]).view(h, w)


class _GALossFn(torch.autograd.Function):
    """Loss tensor whose backward hands out the kernel-computed gradients."""

    @staticmethod
    def forward(ctx, scene, epoch, *params):
        loss = scene._evaluate(epoch)
        ctx.scene = scene
        ctx.version = scene._eval_version
        return loss

    @staticmethod
    def backward(ctx, gout):
        scene = ctx.scene
        if ctx.version != scene._eval_version:
            raise RuntimeError("align3r_b200: backward() of a stale loss - forward() was called again since")
        grads = scene._last_grads()
        return (None, None) + tuple(gout * grads[name] for name in scene._forward_param_names)


class PointCloudOptimizer(nn.Module):
    """ Optimize a global scene, given a list of pairwise observations.
    Graph node: images
    Graph edges: observations = (pred1, pred2)
    """

    def __init__(self, view1, view2, pred1, pred2, *, device=None,
                 # reference optimizer.py:36-38
                 optimize_pp=False, focal_break=20, shared_focal=False, flow_loss_fn='smooth_l1',
                 flow_loss_weight=0.0, depth_regularize_weight=0.0, num_total_iter=300,
                 temporal_smoothing_weight=0, translation_weight=0.1, flow_loss_start_epoch=0.15,
                 flow_loss_thre=50, sintel_ckpt=False, use_self_mask=False, pxl_thre=50,
                 sam2_mask_refine=False, motion_mask_thre=0.35,
                 # reference base_opt.py:61-71
                 dist='l1', conf='log', min_conf_thr=3, thr_for_init_conf=False, base_scale=0.5,
                 allow_pw_adaptors=False, pw_break=20, rand_pose=torch.randn, empty_cache=False, verbose=True,
                 # additions: injected flow, multi-GPU frame-window sharding
                 flow=None, shard=None, edge_rows=None, pack_chunk=32):
        super().__init__()
        if device is None:
            device = 'cuda'
        device = torch.device(device)
        if device.type != 'cuda':
            raise RuntimeError("align3r_b200: PointCloudOptimizer needs a CUDA device (sm_100a); "
                               f"got {device}. There is no CPU fallback.")
        if not torch.cuda.is_available():
            raise RuntimeError("align3r_b200: no CUDA device is available; there is no CPU fallback")
        if device.index is None:
            device = torch.device('cuda', torch.cuda.current_device())
        self._lib = _lib.load()
        self._device = device
        if sam2_mask_refine:
            raise NotImplementedError("sam2_mask_refine is out of scope (SAM-2 pre-step, SURVEY.md section 2 #23)")
        if flow_loss_fn != 'smooth_l1':
            raise NotImplementedError(f"flow_loss_fn={flow_loss_fn!r}: only 'smooth_l1' works in the reference "
                                      "('mse' sets a mistyped attribute, optimizer.py:99-100)")
        if dist not in ('l1', 'l2'):
            raise KeyError(dist)

        # ---------------- graph (reference base_opt.py:73-83) ----------------
        if not isinstance(view1['idx'], list):
            view1['idx'] = view1['idx'].tolist()
        if not isinstance(view2['idx'], list):
            view2['idx'] = view2['idx'].tolist()
        self.edges = [(int(i), int(j)) for i, j in zip(view1['idx'], view2['idx'])]
        self.is_symmetrized = set(self.edges) == {(j, i) for i, j in self.edges}
        self.dist = dist
        self.verbose = verbose
        self.empty_cache = empty_cache
        self.n_imgs = self._check_edges()
        self._edge_index = {edge_str(i, j): e for e, (i, j) in enumerate(self.edges)}

        pred1_pts, pred2_pts = pred1['pts3d'], pred2['pts3d_in_other_view']
        pred1_conf, pred2_conf = pred1['conf'], pred2['conf']
        rows = list(range(len(self.edges))) if edge_rows is None else [int(r) for r in edge_rows]
        self._row_of_edge = {e: r for r, e in enumerate(rows)}
        self.imshapes = self._get_imshapes(pred1_pts, pred2_pts)

        self.min_conf_thr = min_conf_thr
        self.thr_for_init_conf = thr_for_init_conf
        self.conf_mode = conf
        self.conf_trf = get_conf_trf(conf)
        self._conf_untrf = get_conf_untrf(conf)
        self.base_scale = base_scale
        self.norm_pw_scale = True
        self.pw_break = pw_break
        self.POSE_DIM = 7
        self.rand_pose = rand_pose
        self.has_im_poses = True
        self.focal_break = focal_break
        self.num_total_iter = num_total_iter
        self.temporal_smoothing_weight = temporal_smoothing_weight
        self.translation_weight = translation_weight
        self.flow_loss_start_epoch = flow_loss_start_epoch
        self.flow_loss_thre = flow_loss_thre
        self.optimize_pp = optimize_pp
        self.pxl_thre = pxl_thre
        self.motion_mask_thre = motion_mask_thre
        self.shared_focal = shared_focal
        self.flow_loss_weight = flow_loss_weight
        self.depth_regularize_weight = depth_regularize_weight
        self.imshape = self.imshapes[0]
        im_areas = [h * w for h, w in self.imshapes]
        self.max_area = max(im_areas)
        self.total_area_i = sum(im_areas[i] for i, j in self.edges)
        self.total_area_j = sum(im_areas[j] for i, j in self.edges)

        ph = Phases('construct')
        # ---------------- parameters (reference base_opt.py:112-114, optimizer.py:54-75) ----------------
        # drawn on the CPU in the reference's order so that a seeded run starts from the same values - on a worker thread
        # (torch.randn releases the GIL: 7.4 M normal deviates on the 50-frame workload, ~40 ms) while this thread plans,
        # allocates, queues the uploads and sets up the rest; joined below, before the first parameter is written
        nfoc = 1 if shared_focal else self.n_imgs
        drawn = {}

        def _draw():
            try:
                # draws only: an elementwise op on a 147 k-element tensor would open an OpenMP team of its own for this thread
                # (libgomp keeps one pool per master thread), whose spinning threads starve the ranks of a sharded job
                pw = rand_pose((self.n_edges, 1 + self.POSE_DIM))
                depth0 = [torch.randn(H, W) for H, W in self.imshapes]
                im = [rand_pose(self.POSE_DIM) for _ in range(self.n_imgs)]
                drawn.update(pw_poses=pw, depth0=depth0, im_poses=im)
            except BaseException as exc:           # re-raised by the constructor's thread
                drawn['error'] = exc

        import threading
        draw_thread = threading.Thread(target=_draw, name='a3r-rand-params')
        if os.environ.get("A3R_DRAW_THREAD", "1") == "0":      # diagnostic: draw here, on the constructor's thread
            _draw()
            draw_thread = None
        else:
            draw_thread.start()
        # ---------------- plan + shard ----------------
        self._shard = shard
        frame_range = None
        if shard is not None:
            world, rank = int(shard['world']), int(shard['rank'])
            windows = partition_frames(frame_costs(np.asarray(self.edges, dtype=np.int32).reshape(-1, 2),
                                                   np.asarray(self.imshapes, dtype=np.int32)), world)
            frame_range = windows[rank]
            self._windows = windows
        with torch.cuda.device(device):
            n_ctas = self._lib.a3r_max_ctas(device.index)
        if n_ctas <= 0:
            _lib.check(n_ctas, "a3r_max_ctas")
        if os.environ.get("A3R_PLAN_CTAS"):       # diagnostic: another partition of the tiles (tools/partition_sensitivity.py)
            n_ctas = int(os.environ["A3R_PLAN_CTAS"])
        self._plan = build_plan(self.edges, self.imshapes, self._lib.a3r_tile_pixels(), self._lib.a3r_max_slots(),
                                n_ctas, frame_range, align_warps=self._lib.a3r_scan_align_warps())
        plan = self._plan
        self._lslot = plan.local_slot_of()
        self._lframe = plan.local_frame_of()

        ph.mark('plan')
        # ---------------- device buffers; the big host->device uploads start first ----------------
        want_flow = flow_loss_weight > 0
        hyper = self._hyper()
        want_prior = depth_regularize_weight > 0
        self._bufs = ProblemBuffers(plan, device, None, hyper, want_flow, loss_hist_len=max(1024, num_total_iter),
                                    has_prior=want_prior)
        self._conf_sum = torch.zeros(max(plan.n_local_slots, 1), dtype=torch.float64, device=device)
        self._im_conf_plane = torch.zeros(max(plan.n_local_frames, 1), plan.pix_stride, device=device)
        ph.mark('alloc')
        # the parameter tensors exist from here on; the three random ones are filled when the worker is done (below).  Small
        # pageable uploads are synchronous - issued after the big copies they would wait for all of them: they go first
        im_focals = torch.tensor([[focal_break * np.log(max(H, W))] for H, W in self.imshapes[:nfoc]], dtype=torch.float32)
        self.pw_poses = nn.Parameter(torch.empty((self.n_edges, 1 + self.POSE_DIM), device=device))
        self.pw_adaptors = nn.Parameter(torch.zeros((self.n_edges, 2), device=device))
        self.pw_adaptors.requires_grad_(allow_pw_adaptors)
        self.im_depthmaps = nn.Parameter(torch.empty((plan.n_local_frames, plan.pix_stride), device=device))   # log-depth
        self.im_poses = nn.Parameter(torch.empty((self.n_imgs, self.POSE_DIM), device=device))
        self.im_focals = nn.Parameter(im_focals.to(device))
        self.im_pp = nn.Parameter(torch.zeros((self.n_imgs, 2), device=device))
        self.im_pp.requires_grad_(optimize_pp)
        self.register_buffer('_pp', torch.tensor([(w / 2, h / 2) for h, w in self.imshapes], device=device))
        self.register_buffer('_ei', torch.tensor([i for i, j in self.edges], device=device))
        self.register_buffer('_ej', torch.tensor([j for i, j in self.edges], device=device))
        self._bufs.attach_params(dict(pw_poses=self.pw_poses.data, pw_adaptors=self.pw_adaptors.data,
                                      im_depthmaps=self.im_depthmaps.data, im_poses=self.im_poses.data,
                                      im_focals=self.im_focals.data.view(-1), im_pp=self.im_pp.data))
        ph.mark('attach')
        self._dest = None
        self._pack_predictions(pred1_pts, pred2_pts, pred1_conf, pred2_conf, pack_chunk)
        ph.mark('enqueue-pred')
        self._injected_flow = flow
        self._uploads_ready = None
        self._peers, self._peers_tried, self._peer_entry = None, False, None
        self._prior_ready = False
        if want_flow:
            self.flow_ij, self.flow_ji, valid_i, valid_j = self.get_flow(sintel_ckpt)
            self._flow_valid = [valid_i, valid_j]     # None: evaluated from the packed flow planes when somebody reads them
            # on a side stream: the initialisation does not read the flow planes, its kernels need not queue behind
            # these copies; the hot loop joins (self._join_uploads)
            self._upload_stream = torch.cuda.Stream(device)
            self._upload_stream.wait_stream(torch.cuda.current_stream(device))
            self._bufs.arena.record_stream(self._upload_stream)
            with torch.cuda.stream(self._upload_stream):
                self._pack_flow(masks=False)      # asynchronous from pinned host memory; the masks follow below

        ph.mark('enqueue-flow')
        self.im_conf = [self._im_conf_plane[k, :im_areas[n]].view(self.imshapes[n]) for k, n in enumerate(plan.lframe_id.tolist())]
        init_plane = self._im_conf_plane.clone()
        self.init_conf_maps = [init_plane[k, :im_areas[n]].view(self.imshapes[n]) for k, n in enumerate(plan.lframe_id.tolist())]
        self.pred_i, self.pred_j = _SlotDict(self, 0, "pts"), _SlotDict(self, 1, "pts")
        self.conf_i, self.conf_j = _SlotDict(self, 0, "conf"), _SlotDict(self, 1, "conf")

        # ---------------- optional per-frame extras (reference base_opt.py:118-157) ----------------
        self.imgs = None
        if 'img' in view1 and 'img' in view2:
            imgs = [torch.zeros((3,) + hw) for hw in self.imshapes]
            for e in rows:
                r = self._row_of_edge[e]
                imgs[self.edges[e][0]] = view1['img'][r]
                imgs[self.edges[e][1]] = view2['img'][r]
            self.imgs = _rgb(imgs)
        self.dynamic_masks = None
        if 'dynamic_mask' in view1 and 'dynamic_mask' in view2:
            dynamic_masks = [torch.zeros(hw) for hw in self.imshapes]
            for e in rows:
                r = self._row_of_edge[e]
                dynamic_masks[self.edges[e][0]] = view1['dynamic_mask'][r]
                dynamic_masks[self.edges[e][1]] = view2['dynamic_mask'][r]
            self.dynamic_masks = dynamic_masks
        self.camera_poses = None
        if 'camera_pose' in view1 and 'camera_pose' in view2:
            camera_poses = [torch.zeros((4, 4)) for _ in range(self.n_imgs)]
            for e in rows:
                r = self._row_of_edge[e]
                camera_poses[self.edges[e][0]] = view1['camera_pose'][r]
                camera_poses[self.edges[e][1]] = view2['camera_pose'][r]
            if shard is not None and int(shard['world']) > 1:
                # a rank was handed the rows of its own edges only: every frame's pose from the rank(s) that saw it
                acc = np.zeros((self.n_imgs, 17))
                for e in rows:
                    for n in self.edges[e]:
                        acc[n, :16] = torch.as_tensor(camera_poses[n]).double().reshape(-1).numpy()
                        acc[n, 16] = 1.0
                acc = init_fun._allreduce_np(self, acc)
                camera_poses = [torch.from_numpy(acc[n, :16] / max(acc[n, 16], 1.0)).float().reshape(4, 4) for n in range(self.n_imgs)]
            self.camera_poses = camera_poses
        self.img_pathes = None
        if 'instance' in view1 and 'instance' in view2:
            img_pathes = ['' for _ in range(self.n_imgs)]
            for e in rows:
                r = self._row_of_edge[e]
                img_pathes[self.edges[e][0]] = view1['instance'][r]
                img_pathes[self.edges[e][1]] = view2['instance'][r]
            self.img_pathes = img_pathes

        ph.mark('extras')
        # ---------------- flow term inputs (reference optimizer.py:104-116) ----------------
        self.sam2_dynamic_masks = None
        if want_flow:
            assert len(set(self.imshapes)) == 1, "the flow term requires uniform image shapes (reference optimizer.py:528)"
            if use_self_mask:
                # reference optimizer.py:106: the masks are estimated from the pairs themselves (device pre-step)
                self._uploads_ready = self._upload_stream.record_event()
                self.get_motion_mask_from_pairs(view1, view2, pred1, pred2)
            else:
                if self.dynamic_masks is None:
                    raise ValueError("the flow term needs dynamic masks: put 'dynamic_mask' into view1/view2 "
                                     "(the reference fails with torch.stack(None) here, optimizer.py:533)")
                with torch.cuda.stream(self._upload_stream):
                    self._upload_masks()
                    self._uploads_ready = self._upload_stream.record_event()
        if want_prior and not want_flow:
            if self.dynamic_masks is None:
                raise ValueError("depth_regularize_weight > 0 weighs pixels by 1 + dynamic_mask: put 'dynamic_mask' into "
                                 "view1/view2 (the reference fails with torch.stack(None) here, optimizer.py:549)")
            self._upload_masks()
        ph.mark('masks')
        # ---------------- the random initial parameters ----------------
        if draw_thread is not None:
            draw_thread.join()
        if 'error' in drawn:
            raise drawn['error']
        ph.mark('rand-params (join)')
        depth_local = torch.zeros(plan.n_local_frames, plan.pix_stride)
        for k, n in enumerate(plan.lframe_id.tolist()):
            depth_local[k, :im_areas[n]] = drawn['depth0'][n].reshape(-1) / 10 - 3        # reference optimizer.py:54
        with torch.no_grad():
            self.pw_poses.data.copy_(drawn['pw_poses'].float())
            self.im_poses.data.copy_(torch.stack(drawn['im_poses']).float())
            self.im_depthmaps.data.copy_(depth_local)     # (queues behind the prediction uploads of this stream)
        del drawn, depth_local
        ph.mark('param-upload')
        ph.done()
        self._adam_step = 0
        self._eval_version = 0
        self._forward_param_names = ()
        self.init_depthmap = None

    # =====================================================================================
    # construction helpers
    # =====================================================================================
    def _hyper(self):
        return dict(shared_focal=self.shared_focal, norm_pw_scale=self.norm_pw_scale, dist=self.dist,
                    focal_break=float(self.focal_break), pw_break=float(self.pw_break), base_scale=float(self.base_scale),
                    flow_loss_weight=float(self.flow_loss_weight), flow_loss_thre=float(self.flow_loss_thre),
                    pxl_thre=float(self.pxl_thre), temporal_smoothing_weight=float(self.temporal_smoothing_weight),
                    translation_weight=float(self.translation_weight),
                    depth_regularize_weight=float(self.depth_regularize_weight))

    def _check_edges(self):
        return check_edges(self.edges)

    def _get_imshapes(self, pred_i, pred_j):
        """reference commons.py:38-50 (rows may cover only this shard's edges)."""
        imshapes = [None] * self.n_imgs
        for e, r in self._row_of_edge.items():
            i, j = self.edges[e]
            for n, shp in ((i, _row_hw(pred_i, r)), (j, _row_hw(pred_j, r))):
                if imshapes[n]:
                    assert imshapes[n] == shp, f'incorrect shape for image {n}'
                imshapes[n] = shp
        if any(s is None for s in imshapes):
            known = next(s for s in imshapes if s is not None)
            imshapes = [s or known for s in imshapes]     # frames of other shards: uniform-shape assumption
        return imshapes

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self._device).cuda_stream)

    def _check_prior(self):
        if self._bufs.has_prior and not self._prior_ready:
            raise RuntimeError("depth_regularize_weight > 0 needs the initial depth maps: run an initialisation first "
                               "(the reference fails in get_init_depthmaps here, optimizer.py:547)")

    def _join_uploads(self):
        """The flow planes / static masks were uploaded on a side stream: order the current stream after them."""
        if self._uploads_ready is not None:
            torch.cuda.current_stream(self._device).wait_event(self._uploads_ready)
            self._uploads_ready = None

    def _dest_tables(self, side):
        """(local slot, local frame) of every row this device packs, per side - uploaded once, before the first big copy is
        queued (a pageable upload is synchronous: issued later it would wait for all the copies ahead of it)."""
        if self._dest is None:
            rows = [[(e, r) for e, r in sorted(self._row_of_edge.items()) if (2 * e + sd) in self._lslot] for sd in (0, 1)]
            flat = []
            for sd in (0, 1):
                flat += [self._lslot[2 * e + sd] for e, _ in rows[sd]] + [self._lframe[self.edges[e][sd]] for e, _ in rows[sd]]
            t = torch.tensor(flat, dtype=torch.int32).to(self._device)
            n0, n1 = len(rows[0]), len(rows[1])
            self._dest = ((t[:n0], t[n0:2 * n0]), (t[2 * n0:2 * n0 + n1], t[2 * n0 + n1:]))
        return self._dest[side]

    def _pack_predictions(self, pts1, pts2, conf1, conf2, chunk):
        chunk = max(1, min(int(chunk), self._bufs.stage_rows))
        """Upload the raw pair predictions chunk by chunk and pack them into the slot planes."""
        plan, dev, b = self._plan, self._device, self._bufs
        mode = _lib.CONF_MODES[self.conf_mode]
        with torch.cuda.device(dev):
            for side, (pts, conf) in enumerate(((pts1, conf1), (pts2, conf2))):
                todo = [(e, r) for e, r in sorted(self._row_of_edge.items()) if (2 * e + side) in self._lslot]
                slots_all, lfr_all = self._dest_tables(side)
                # group consecutive rows of equal shape into one upload
                k = 0
                while k < len(todo):
                    e0, r0 = todo[k]
                    shp = _row_hw(pts, r0)
                    m = 1
                    while (k + m < len(todo) and m < chunk and todo[k + m][1] == r0 + m
                           and _row_hw(pts, r0 + m) == shp):
                        m += 1
                    P = shp[0] * shp[1]
                    # staged through the problem's one staging area: copy(k+1) queues behind pack(k) on this stream
                    dp = b.stage[:m * P * 3].view(m, P, 3)
                    dc = b.stage[m * P * 3:m * P * 4].view(m, P)
                    if _is_stack(pts) and _is_stack(conf):      # one (E,H,W,3) tensor, or rows drawn on the device on demand
                        dp.copy_(pts[r0:r0 + m].reshape(m, P, 3), non_blocking=True)
                        dc.copy_(conf[r0:r0 + m].reshape(m, P), non_blocking=True)
                    else:
                        dp.copy_(torch.stack([torch.as_tensor(pts[r0 + q]).reshape(P, 3) for q in range(m)]), non_blocking=True)
                        dc.copy_(torch.stack([torch.as_tensor(conf[r0 + q]).reshape(P) for q in range(m)]), non_blocking=True)
                    slots, lfr = slots_all[k:k + m], lfr_all[k:k + m]
                    _lib.check(self._lib.a3r_pack_pred(_lib.ptr(dp), _lib.ptr(dc), m, P, _lib.ptr(slots), _lib.ptr(lfr),
                                                       plan.pix_stride, mode, _lib.ptr(b.predw), _lib.ptr(self._im_conf_plane),
                                                       _lib.ptr(self._conf_sum), self._stream()), "a3r_pack_pred")
                    k += m

    def _pack_flow(self, chunk=32, masks=True):
        chunk = max(1, min(int(chunk), self._bufs.stage_rows))
        plan, dev, b = self._plan, self._device, self._bufs
        H, W = self.imshape
        P = H * W
        with torch.cuda.device(dev):
            for side, fl in enumerate((self.flow_ij, self.flow_ji)):
                todo = [(e, r) for e, r in sorted(self._row_of_edge.items()) if (2 * e + side) in self._lslot]
                slots_all, _ = self._dest_tables(side)
                k = 0
                while k < len(todo):
                    # one upload per run of consecutive rows (a frame-window shard packs a subset of its rows on side j):
                    # slices of the caller's pinned tensor copy asynchronously, a gathered copy would be pageable
                    r0 = todo[k][1]
                    m = 1
                    while k + m < len(todo) and m < chunk and todo[k + m][1] == r0 + m:
                        m += 1
                    d = b.stage[:m * 2 * P].view(m, 2, P)
                    d.copy_(fl[r0:r0 + m].reshape(m, 2, P), non_blocking=True)
                    slots = slots_all[k:k + m]
                    _lib.check(self._lib.a3r_pack_flow(_lib.ptr(d), m, P, _lib.ptr(slots), plan.pix_stride,
                                                       _lib.ptr(b.flow), self._stream()), "a3r_pack_flow")
                    k += m
            if masks:
                self._upload_masks()

    def _upload_masks(self):
        """static_mask = ~dynamic_mask per local frame (reference optimizer.py:533-537)."""
        b, plan = self._bufs, self._plan
        masks = self.sam2_dynamic_masks if self.sam2_dynamic_masks is not None else self.dynamic_masks
        lfr = plan.lframe_id.tolist()
        # through pinned memory, asynchronously: a pageable upload is synchronous - on the side stream it would hold the
        # constructor until the flow planes queued ahead of it have landed (the caching host allocator keeps the pinned
        # block until the copy has run)
        def up(t):
            t = t.contiguous()
            return (t if t.is_cuda else t.pin_memory()).to(self._device, non_blocking=True)

        if len(lfr) and len({tuple(torch.as_tensor(masks[n]).shape) for n in lfr}) == 1:      # one upload for all frames
            m = torch.stack([torch.as_tensor(masks[n]).bool().reshape(-1) for n in lfr])
            b.static_mask[:, :m.shape[1]] = up((~m).to(torch.uint8))
            return
        for k, n in enumerate(lfr):
            m = torch.as_tensor(masks[n]).bool().reshape(-1)
            b.static_mask[k, :m.numel()] = up((~m).to(torch.uint8))

    def get_motion_mask_from_pairs(self, view1, view2, pred1, pred2):
        """reference optimizer.py:156-235 (``use_self_mask=True``): see cloud_opt_flow/motion_mask.py."""
        from .motion_mask import get_motion_mask_from_pairs
        return get_motion_mask_from_pairs(self, view1, view2, pred1, pred2)

    def get_flow(self, sintel_ckpt=False):
        """reference optimizer.py:118-154 runs RAFT here.  RAFT (and its weights) are outside this
        package: hand the precomputed flows in with ``flow=(flow_ij, flow_ji)`` - shape (E,2,H,W) each,
        the layout RAFT returns - or override this method exactly as one would on the reference."""
        if self._injected_flow is None:
            raise NotImplementedError("flow_loss_weight > 0 needs optical flow: pass flow=(flow_ij, flow_ji) to "
                                      "global_aligner / PointCloudOptimizer (RAFT inference is out of scope)")
        flow_ij, flow_ji = self._injected_flow[:2]
        # the reference also returns OccMask validity maps, which it stores and never reads (optimizer.py:105-111,146-147):
        # None here - flow_valid_mask_i / _j evaluate them on the device from the packed flow planes on first access
        return flow_ij, flow_ji, None, None

    def _owner_of_frame(self, n):
        for r, (a, b) in enumerate(self._windows):
            if a <= n < b:
                return r
        raise ValueError(n)

    def _fetch_remote_planes(self, planes, want):
        """Frame-window shards: the rows of ``planes`` ([local slots][pix_stride](,C)) of the global slots in ``want`` that
        live on other GPUs, fetched from their owners (point-to-point, one batch).  Collective: every rank calls it with
        its own wish list.  Returns {global slot: tensor}."""
        import torch.distributed as dist
        world, rank, group = int(self._shard['world']), int(self._shard['rank']), self._shard.get('group')
        want = sorted(int(g) for g in want)
        wishes = [None] * world
        dist.all_gather_object(wishes, want, group=group)
        ops, got = [], {}
        for r, wl in enumerate(wishes):              # the same order on sender and receiver: (receiver rank, slot)
            for gs in wl:
                owner = self._owner_of_frame(self.edges[gs >> 1][gs & 1])
                peer_g = (lambda q: dist.get_global_rank(group, q) if group is not None else q)
                if owner == rank and r != rank:
                    ops.append(dist.P2POp(dist.isend, planes[self._lslot[gs]].contiguous(), peer_g(r), group=group))
                elif r == rank and owner != rank:
                    got[gs] = torch.empty_like(planes[0])
                    ops.append(dist.P2POp(dist.irecv, got[gs], peer_g(owner), group=group))
        if ops:
            for q in dist.batch_isend_irecv(ops):
                q.wait()
        return got

    def _flow_valid_mask(self, side):
        """reference goem_opt.py:575-618 (OccMask, th = 3): (E,1,H,W) bool, forward flow = this side's, backward = the other's.
        Frame-window shards evaluate the rows whose forward flow lives here (``flow_valid_rows(side)``; the other rows stay
        False); the backward planes of the pairs that straddle a window boundary are fetched from the neighbour."""
        if getattr(self, '_flow_valid', None) is None:
            return None
        if self._flow_valid[side] is None:
            E, (H, W) = self.n_edges, self.imshape
            dev = self._device
            rows = self.flow_valid_rows(side)
            valid = torch.zeros((E, H * W), dtype=torch.uint8, device=dev)

            def run(table, fwd_idx, bwd_idx, dst_rows):
                if not dst_rows:
                    return
                fwd = torch.tensor(fwd_idx, dtype=torch.int32).to(dev)
                bwd = torch.tensor(bwd_idx, dtype=torch.int32).to(dev)
                part = valid if len(dst_rows) == E else torch.empty((len(dst_rows), H * W), dtype=torch.uint8, device=dev)
                with torch.cuda.device(dev):
                    _lib.check(self._lib.a3r_occ_masks(_lib.ptr(table), self._plan.pix_stride, _lib.ptr(fwd), _lib.ptr(bwd), len(dst_rows),
                                                       H, W, 3.0, _lib.ptr(part), self._stream()), "a3r_occ_masks")
                if part is not valid:
                    valid[torch.tensor(dst_rows, device=dev)] = part

            self._join_uploads()
            here = [e for e in rows if (2 * e + 1 - side) in self._lslot]
            run(self._bufs.flow, [self._lslot[2 * e + side] for e in here], [self._lslot[2 * e + 1 - side] for e in here], here)
            if self._shard is not None and int(self._shard['world']) > 1:
                away = [e for e in rows if (2 * e + 1 - side) not in self._lslot]         # pairs that straddle a window boundary
                remote = self._fetch_remote_planes(self._bufs.flow, [2 * e + 1 - side for e in away])
                if away:                     # a small table: [forward planes of those rows | the fetched backward planes]
                    table = torch.stack([self._bufs.flow[self._lslot[2 * e + side]] for e in away] + [remote[2 * e + 1 - side] for e in away])
                    run(table, list(range(len(away))), list(range(len(away), 2 * len(away))), away)
            self._flow_valid[side] = valid.view(E, 1, H, W).bool()
        return self._flow_valid[side]

    def flow_valid_rows(self, side):
        """Edges whose row of flow_valid_mask_i (side 0) / _j (side 1) this device evaluates: all of them on one GPU, the
        edges whose side-`side` frame lives here on a frame-window shard."""
        return [e for e in range(self.n_edges) if (2 * e + side) in self._lslot]

    @property
    def flow_valid_mask_i(self):
        return self._flow_valid_mask(0)

    @flow_valid_mask_i.setter
    def flow_valid_mask_i(self, v):
        self._flow_valid = [v, (getattr(self, '_flow_valid', None) or [None, None])[1]]

    @property
    def flow_valid_mask_j(self):
        return self._flow_valid_mask(1)

    @flow_valid_mask_j.setter
    def flow_valid_mask_j(self, v):
        self._flow_valid = [(getattr(self, '_flow_valid', None) or [None, None])[0], v]

    # =====================================================================================
    # properties (reference base_opt.py:159-173)
    # =====================================================================================
    @property
    def n_edges(self):
        return len(self.edges)

    @property
    def str_edges(self):
        return [edge_str(i, j) for i, j in self.edges]

    @property
    def imsizes(self):
        return [(w, h) for h, w in self.imshapes]

    @property
    def device(self):
        return self._device

    @property
    def flow_loss_flag(self):
        """True once the flow gate has tripped (reference optimizer.py:46,540-542); reads the device."""
        return bool(self._bufs.scal[_lib.S_FLAG_STICKY].item() != 0)

    @flow_loss_flag.setter
    def flow_loss_flag(self, v):
        self._bufs.scal[_lib.S_FLAG_STICKY] = 1.0 if v else 0.0

    def to(self, *args, **kwargs):
        """global_aligner(...).to(device) compatibility: buffers already live on the CUDA device."""
        dev = kwargs.get('device', args[0] if args else None)
        if dev is not None and isinstance(dev, (str, torch.device, int)):
            d = torch.device(dev) if not isinstance(dev, int) else torch.device('cuda', dev)
            if d.type != 'cuda' or (d.index is not None and d.index != self._device.index):
                raise RuntimeError(f"align3r_b200: the optimiser lives on {self._device} and cannot be moved to {d}")
        return self

    def state_dict(self, trainable=True):
        """reference base_opt.py:175-177."""
        all_params = super().state_dict()
        return {k: v for k, v in all_params.items() if k.startswith(('_', 'pred_i.', 'pred_j.', 'conf_i.', 'conf_j.')) != trainable}

    def load_state_dict(self, data):
        """reference base_opt.py:179-180."""
        return super().load_state_dict(self.state_dict(trainable=False) | data)

    # =====================================================================================
    # parameter maps / getters (reference base_opt.py:195-253, optimizer.py:392-498)
    # =====================================================================================
    def get_adaptors(self):
        adapt = self.pw_adaptors
        adapt = torch.cat((adapt[:, 0:1], adapt), dim=-1)  # (scale_xy, scale_xy, scale_z)
        if self.norm_pw_scale:  # normalize so that the product == 1
            adapt = adapt - adapt.mean(dim=1, keepdim=True)
        return (adapt / self.pw_break).exp()

    def _get_poses(self, poses):
        Q = poses[:, :4]
        T = signed_expm1(poses[:, 4:7])
        Q = Q / torch.norm(Q, dim=-1, keepdim=True)
        R = unitquat_to_rotmat(Q)
        top = torch.cat((R, T.unsqueeze(-1)), dim=-1)
        bottom = torch.zeros_like(top[:, :1, :])
        bottom[:, 0, 3] = 1
        return torch.cat((top, bottom), dim=1)

    def _set_pose(self, poses, idx, R, T=None, scale=None, force=False):
        # all poses == cam-to-world  (reference base_opt.py:209-228)
        pose = poses[idx]
        if not (pose.requires_grad or force):
            return pose
        if R.shape == (4, 4):
            assert T is None
            T = R[:3, 3]
            R = R[:3, :3]
        if R is not None:
            pose.data[0:4] = rotmat_to_unitquat(torch.as_tensor(R)).to(pose.device, pose.dtype)
        if T is not None:
            pose.data[4:7] = signed_log1p(torch.as_tensor(T, dtype=torch.float32).to(pose.device) / (scale or 1))
        if scale is not None:
            assert poses.shape[-1] in (8, 13)
            pose.data[-1] = np.log(float(scale))
        return pose

    def get_pw_norm_scale_factor(self):
        if self.norm_pw_scale:
            # normalize scales so that things cannot go south: exp(scale) ~= self.base_scale
            return (np.log(self.base_scale) - self.pw_poses[:, -1].mean()).exp()
        return 1  # don't norm scale for known poses

    def get_pw_scale(self):
        scale = self.pw_poses[:, -1].exp()  # (n_edges,)
        return scale * self.get_pw_norm_scale_factor()

    def get_pw_poses(self):  # cam to world
        RT = self._get_poses(self.pw_poses)
        scaled_RT = RT.clone()
        scaled_RT[:, :3] *= self.get_pw_scale().view(-1, 1, 1)  # scale the rotation AND translation
        return scaled_RT

    def get_masks(self):
        maps = self.init_conf_maps if self.thr_for_init_conf else self.im_conf
        return [(conf > self.min_conf_thr) for conf in maps]

    def get_conf(self, mode=None):
        trf = self.conf_trf if mode is None else get_conf_trf(mode)
        return [trf(c) for c in self.im_conf]

    def get_init_conf(self, mode=None):
        trf = self.conf_trf if mode is None else get_conf_trf(mode)
        return [trf(c) for c in self.init_conf_maps]

    def _set_focal(self, idx, focal, force=False):
        param = self.im_focals[idx]
        if param.requires_grad or force:  # can only init a parameter not already initialized
            param.data[:] = self.focal_break * np.log(float(focal))
        return param

    def get_focals(self):
        if self.shared_focal:
            log_focals = torch.stack([self.im_focals[0]] * self.n_imgs, dim=0)
        else:
            log_focals = self.im_focals
        return (log_focals / self.focal_break).exp()

    def get_known_focal_mask(self):
        return torch.tensor([not self.im_focals.requires_grad] * len(self.im_focals))

    def _set_principal_point(self, idx, pp, force=False):
        param = self.im_pp[idx]
        H, W = self.imshapes[idx]
        if param.requires_grad or force:
            param.data[:] = (torch.as_tensor(np.asarray(pp, dtype=np.float32)) - torch.tensor((W / 2, H / 2))).to(param.device) / 10
        return param

    def get_principal_points(self):
        return self._pp + 10 * self.im_pp

    def get_intrinsics(self):
        K = torch.zeros((self.n_imgs, 3, 3), device=self.device)
        focals = self.get_focals().flatten()
        K[:, 0, 0] = K[:, 1, 1] = focals
        K[:, :2, 2] = self.get_principal_points()
        K[:, 2, 2] = 1
        return K

    def get_im_poses(self):  # cam to world
        return self._get_poses(self.im_poses)

    def _ravel_depth(self, depth):
        depth = torch.as_tensor(depth, dtype=torch.float32).reshape(-1)
        if depth.numel() < self._plan.pix_stride:
            depth = torch.cat((depth, depth.new_zeros(self._plan.pix_stride - depth.numel())))
        return depth

    def _set_depthmap(self, idx, depth, force=False):
        """reference optimizer.py:430-436; idx is a GLOBAL frame index owned by this device."""
        k = self._lframe[idx]
        param = self.im_depthmaps[k]
        if param.requires_grad or force:
            d = self._ravel_depth(depth).to(self.device)
            param.data[:] = d.log().nan_to_num(neginf=0)
        return param

    def _set_init_depthmap(self):
        depth_maps = self.get_depthmaps(raw=True)
        self.init_depthmap = [dm.detach().clone() for dm in depth_maps]
        if self._bufs.has_prior:        # what the depth prior compares with: log(clamp(init depth, 1e-6)), goem_opt.py:17-20
            self._bufs.logdepth_init.copy_(torch.stack(self.init_depthmap).clamp(min=1e-6).log())
            self._prior_ready = True

    def get_init_depthmaps(self, raw=False):
        res = self.init_depthmap
        if not raw:
            res = [dm[:h * w].view(h, w) for dm, (h, w) in zip(res, self._local_shapes())]
        return res

    def _local_shapes(self):
        return [self.imshapes[n] for n in self._plan.lframe_id.tolist()]

    def get_depthmaps(self, raw=False):
        """Depth maps of the frames this device owns (all frames on a single GPU)."""
        res = self.im_depthmaps.exp()
        if not raw:
            res = [dm[:h * w].view(h, w) for dm, (h, w) in zip(res, self._local_shapes())]
        return res

    def depth_to_pts3d(self):
        """World pointmaps (local frames, pix_stride, 3) - reference optimizer.py:468-478, by kernel."""
        b = self._bufs
        out = torch.empty(max(self._plan.n_local_frames, 1), self._plan.pix_stride, 3, device=self.device)
        with torch.cuda.device(self._device):
            st = b.struct(self._trainable())
            _lib.check(self._lib.a3r_ga_prepare(C.byref(st), 0, self._stream()), "a3r_ga_prepare")
            _lib.check(self._lib.a3r_ga_unproject(C.byref(st), _lib.ptr(out), self._stream()), "a3r_ga_unproject")
        return out

    @torch.no_grad()
    def clean_pointcloud(self, tol=0.001, bad_conf=0, dbg=()):
        """reference base_opt.py:293-303,587-623: lower ``im_conf`` where a point of one frame lands in front of another
        frame's depth map while being less confident.  One kernel launch per frame (a3r_clean_pointcloud).
        Frame-window shards: every frame meets every other frame's depth map and confidence, and frame i reads the FINAL
        confidences of the frames before it - the depth and confidence planes of all frames are gathered on every rank
        (N x P floats each), then the windows are cleaned in rank order, each rank broadcasting its frames' confidences
        when it is done (the reference's sequential semantics, exactly)."""
        assert 0 <= tol < 1
        dev, plan = self._device, self._plan
        cams = torch.linalg.inv(self.get_im_poses().detach()).float().contiguous()            # world to camera
        K = self.get_intrinsics().detach()
        intr = torch.stack((K[:, 0, 0], K[:, 1, 1], K[:, 0, 2], K[:, 1, 2]), dim=1).float().contiguous()
        pts = self.depth_to_pts3d().contiguous()
        depth = self.im_depthmaps.detach().exp().contiguous()
        hw = torch.tensor(self.imshapes, dtype=torch.int32).to(dev)
        if self._shard is not None and int(self._shard['world']) > 1:
            import torch.distributed as dist
            group, rank = self._shard.get('group'), int(self._shard['rank'])
            depth_all = torch.zeros(self.n_imgs, plan.pix_stride, device=dev)
            conf_all = torch.zeros(self.n_imgs, plan.pix_stride, device=dev)
            lo, hi = self._windows[rank]
            depth_all[lo:hi] = depth
            conf_all[lo:hi] = self._im_conf_plane[:hi - lo]
            for r, (a, b) in enumerate(self._windows):                  # everybody's depth maps and initial confidences
                if b > a:
                    dist.broadcast(depth_all[a:b], src=dist.get_global_rank(group, r) if group is not None else r, group=group)
                    dist.broadcast(conf_all[a:b], src=dist.get_global_rank(group, r) if group is not None else r, group=group)
            for r, (a, b) in enumerate(self._windows):                  # windows in frame order; the results travel on
                if b <= a:
                    continue
                if r == rank:
                    with torch.cuda.device(dev):
                        _lib.check(self._lib.a3r_clean_pointcloud_range(_lib.ptr(pts), _lib.ptr(depth_all), _lib.ptr(cams), _lib.ptr(intr),
                                                                        _lib.ptr(hw), self.n_imgs, a, b, plan.pix_stride, float(tol),
                                                                        float(bad_conf), _lib.ptr(conf_all), self._stream()),
                                   "a3r_clean_pointcloud_range")
                dist.broadcast(conf_all[a:b], src=dist.get_global_rank(group, r) if group is not None else r, group=group)
            self._im_conf_plane[:hi - lo] = conf_all[lo:hi]
            return self
        with torch.cuda.device(dev):
            _lib.check(self._lib.a3r_clean_pointcloud(_lib.ptr(pts), _lib.ptr(depth), _lib.ptr(cams), _lib.ptr(intr), _lib.ptr(hw),
                                                      self.n_imgs, plan.pix_stride, float(tol), float(bad_conf),
                                                      _lib.ptr(self._im_conf_plane), self._stream()), "a3r_clean_pointcloud")
        return self

    def get_pts3d(self, raw=False, **kwargs):
        res = self.depth_to_pts3d()
        if not raw:
            res = [dm[:h * w].view(h, w, 3) for dm, (h, w) in zip(res, self._local_shapes())]
        return res

    # =====================================================================================
    # presets (reference optimizer.py:315-450)
    # =====================================================================================
    def _check_all_imgs_are_selected(self, msk):
        assert np.all(np.asarray(self._get_msk_indices(msk)) == np.arange(self.n_imgs)), 'incomplete mask!'

    def _get_msk_indices(self, msk):
        if msk is None:
            return range(self.n_imgs)
        if isinstance(msk, int):
            return [msk]
        if isinstance(msk, (tuple, list)):
            return self._get_msk_indices(np.array(msk))
        if msk.dtype in (bool, torch.bool, np.bool_):
            assert len(msk) == self.n_imgs
            return np.where(msk)[0]
        if np.issubdtype(msk.dtype, np.integer):
            return msk
        raise ValueError(f'bad {msk=}')

    def _no_grad(self, tensor):
        assert tensor.requires_grad, 'it must be True at this point, otherwise no modification occurs'

    def preset_pose(self, known_poses, pose_msk=None, requires_grad=False):  # cam-to-world
        self._check_all_imgs_are_selected(pose_msk)
        if isinstance(known_poses, torch.Tensor) and known_poses.ndim == 2:
            known_poses = [known_poses]
        for idx, pose in zip(self._get_msk_indices(pose_msk), known_poses):
            pose = torch.as_tensor(np.asarray(pose), dtype=torch.float32)
            if pose.shape == (7,):   # tum pose x y z qw qx qy qz (the reference's helper for this is broken, goem_opt.py:8-13)
                from scipy.spatial.transform import Rotation
                M = torch.eye(4)
                M[:3, :3] = torch.from_numpy(Rotation.from_quat(pose[[4, 5, 6, 3]].numpy()).as_matrix()).float()
                M[:3, 3] = pose[:3]
                pose = M
            if self.verbose:
                print(f' (setting pose #{idx} = {pose[:3, 3]})')
            self._no_grad(self._set_pose(self.im_poses, idx, pose))
        if len(known_poses) == self.n_imgs:
            self.im_poses.requires_grad_(bool(requires_grad))
        self.norm_pw_scale = False        # reference optimizer.py:340

    def preset_intrinsics(self, known_intrinsics, msk=None):
        if isinstance(known_intrinsics, torch.Tensor) and known_intrinsics.ndim == 2:
            known_intrinsics = [known_intrinsics]
        for K in known_intrinsics:
            assert K.shape == (3, 3)
        self.preset_focal([K.diagonal()[:2].mean() for K in known_intrinsics], msk)
        if self.optimize_pp:
            self.preset_principal_point([K[:2, 2] for K in known_intrinsics], msk)

    def preset_focal(self, known_focals, msk=None, requires_grad=False):
        self._check_all_imgs_are_selected(msk)
        for idx, focal in zip(self._get_msk_indices(msk), known_focals):
            if self.verbose:
                print(f' (setting focal #{idx} = {focal})')
            if self.shared_focal and idx > 0:
                continue
            self._no_grad(self._set_focal(idx, focal))
        if len(known_focals) == self.n_imgs:
            self.im_focals.requires_grad_(bool(requires_grad))

    def preset_principal_point(self, known_pp, msk=None):
        self._check_all_imgs_are_selected(msk)
        for idx, pp in zip(self._get_msk_indices(msk), known_pp):
            if self.verbose:
                print(f' (setting principal point #{idx} = {pp})')
            self._no_grad(self._set_principal_point(idx, pp))
        self.im_pp.requires_grad_(False)

    def preset_depthmap(self, known_depthmaps, msk=None, requires_grad=False):
        self._check_all_imgs_are_selected(msk)
        for idx, depth in zip(self._get_msk_indices(msk), known_depthmaps):
            if self.verbose:
                print(f' (setting depthmap #{idx})')
            if idx in self._lframe:
                self._no_grad(self._set_depthmap(idx, depth))
        if len(known_depthmaps) == self.n_imgs:
            self.im_depthmaps.requires_grad_(bool(requires_grad))

    # =====================================================================================
    # the hot loop
    # =====================================================================================
    def _trainable(self):
        return dict(pw_poses=self.pw_poses.requires_grad, im_depthmaps=self.im_depthmaps.requires_grad,
                    im_poses=self.im_poses.requires_grad, im_focals=self.im_focals.requires_grad,
                    im_pp=self.im_pp.requires_grad, pw_adaptors=self.pw_adaptors.requires_grad)

    def _flow_active(self, epoch):
        """reference optimizer.py:520."""
        return self.flow_loss_weight > 0 and epoch >= self.num_total_iter * self.flow_loss_start_epoch

    def _refresh_hyper(self):
        self._bufs.hyper.update(self._hyper())

    def _reduce_accumulators(self):
        """Multi-GPU exchange step: sum the small-parameter gradient accumulators of all frame windows
        (SURVEY.md section 8(e)); nothing per-pixel crosses GPUs."""
        if self._shard is not None and int(self._shard['world']) > 1:
            import torch.distributed as dist
            dist.all_reduce(self._bufs.acc, op=dist.ReduceOp.SUM, group=self._shard.get('group'))

    def _setup_peers(self):
        """Frame-window sharded runs on one node: map every rank's accumulator region (CUDA IPC over NVLink) so that the
        per-iteration exchange happens inside our own update kernel (a3r_ga_run_peers: every value is read from the one rank
        that owns it) instead of a library all-reduce between launches.  Returns the A3rPeers descriptor, or None when the
        NCCL path is to be used (A3R_EXCHANGE=nccl, more ranks than A3R_MAX_PEERS, or peer mapping unavailable on some rank -
        all ranks take the same decision).  Regions and mappings are cached per (device, world, rank, block size) and
        outlive the optimiser: the clips of a hierarchical run, or consecutive jobs of one process, pay the cudaMalloc + IPC
        handshake (milliseconds) once - ``free_peer_regions()`` releases them."""
        if self._peers is not None or self._peers_tried:
            return self._peers
        self._peers_tried = True
        import torch.distributed as dist
        world, rank, group = int(self._shard['world']), int(self._shard['rank']), self._shard.get('group')
        if not (dist.is_available() and dist.is_initialized()):
            return None
        want = os.environ.get("A3R_EXCHANGE", "peer") != "nccl" and world <= _lib.MAX_PEERS
        n_acc = int(self._bufs.n_acc)
        key = (self._device.index, world, rank, n_acc, id(group))
        entry = _PEER_REGIONS.get(key)
        if entry is None:
            region, handle, ok = C.c_void_p(), C.create_string_buffer(64), want
            with torch.cuda.device(self._device):
                if want and self._lib.a3r_peer_alloc(n_acc, C.byref(region), handle) != 0:
                    ok = False
                handles = [None] * world
                dist.all_gather_object(handles, (bool(ok), bytes(handle.raw)), group=group)
                ok = all(h[0] for h in handles)
                peers = _lib.A3rPeers()
                peers.world, peers.rank, peers.n_acc = world, rank, n_acc
                opened = []
                if ok:
                    for r in range(world):
                        if r == rank:
                            peers.base[r] = region.value
                            continue
                        ptr = C.c_void_p()
                        if self._lib.a3r_peer_open(handles[r][1], C.byref(ptr)) != 0:
                            ok = False
                            break
                        opened.append(ptr)
                        peers.base[r] = ptr.value
                flag = torch.tensor([1 if ok else 0], device=self._device)
                dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
                if int(flag.item()) == 0:                       # some rank cannot map its peers: everybody uses NCCL
                    for ptr in opened:
                        self._lib.a3r_peer_close(ptr)
                    if region.value:
                        self._lib.a3r_peer_free(region)
                    return None
            entry = dict(peers=peers, region=region, opened=opened, tag=0, group=group, device=self._device, lib=self._lib)
            _PEER_REGIONS[key] = entry
        for r, (lo, hi) in enumerate(self._windows):
            entry['peers'].win[r], entry['peers'].win[r + 1] = lo, hi
        self._peers, self._peer_entry = entry['peers'], entry
        return self._peers

    @property
    def _peer_tag(self):
        return self._peer_entry['tag']

    @_peer_tag.setter
    def _peer_tag(self, v):
        self._peer_entry['tag'] = int(v)

    def close(self):
        """Detaches the optimiser from the cached peer-exchange region of a frame-window sharded run (the region itself
        stays mapped for the next optimiser of this process - ``free_peer_regions()`` releases it).  Not collective."""
        self._peers, self._peers_tried, self._peer_entry = None, False, None

    def _phase_accumulate(self, mode):
        """parameter maps + fused forward/backward scan over this device's pixel-slots."""
        with torch.cuda.device(self._device):
            self._join_uploads()
            self._check_prior()
            st = self._bufs.struct(self._trainable())
            s = self._stream()
            _lib.check(self._lib.a3r_ga_prepare(C.byref(st), mode, s), "a3r_ga_prepare")
            _lib.check(self._lib.a3r_ga_accumulate(C.byref(st), mode, s), "a3r_ga_accumulate")

    def _phase_update(self, mode, adam):
        """loss/gate + parameter chains (+ Adam when mode has M_APPLY_ADAM)."""
        with torch.cuda.device(self._device):
            st = self._bufs.struct(self._trainable())
            _lib.check(self._lib.a3r_ga_update(C.byref(st), mode, adam, self._stream()), "a3r_ga_update")

    def _evaluate(self, epoch):
        """loss + gradients of the current parameters, no update.  Returns a 0-dim fp32 tensor."""
        self._refresh_hyper()
        mode = _lib.M_FLOW_ACTIVE if self._flow_active(epoch) else 0
        self._phase_accumulate(mode)
        self._reduce_accumulators()
        self._phase_update(mode, adam_schedule([0.0], first_iter=-1))
        self._eval_version += 1
        return self._bufs.scal[_lib.S_LOSS].to(torch.float32)

    def _last_grads(self):
        b = self._bufs
        g = {k: v.view_as(getattr(self, k)).clone() for k, v in b.grads.items()}
        g["im_depthmaps"] = b.g_align[:self._plan.n_local_frames].clone()
        return g

    def forward(self, epoch=9999):
        """Loss of the current parameters (reference optimizer.py:500-557).  The returned tensor supports
        ``.backward()``: it deposits the kernel-computed gradients into the parameters' ``.grad``."""
        names = [n for n in ("pw_poses", "pw_adaptors", "im_depthmaps", "im_poses", "im_focals", "im_pp")
                 if getattr(self, n).requires_grad]
        if torch.is_grad_enabled() and names:
            self._forward_param_names = tuple(names)
            return _GALossFn.apply(self, epoch, *[getattr(self, n) for n in names])
        return self._evaluate(epoch)

    @torch.no_grad()
    def compute_global_alignment(self, init=None, init_priors=None, save_score_path=None, save_score_only=False,
                                 niter_PnP=10, **kw):
        """reference base_opt.py:459-474."""
        if init is None:
            pass
        elif init == 'msp' or init == 'mst':
            init_fun.init_minimum_spanning_tree(self, save_score_path=save_score_path, save_score_only=save_score_only,
                                                niter_PnP=niter_PnP, init_priors=init_priors)
            if save_score_only:
                return None
        elif init == 'known_poses':
            self.preset_pose(known_poses=self.camera_poses, requires_grad=True)
            init_fun.init_from_known_poses(self, min_conf_thr=self.min_conf_thr, niter_PnP=niter_PnP)
        else:
            raise ValueError(f'bad value for {init=}')
        return global_alignment_loop(self, **kw)

    # ---- read-out helpers (reference base_opt.py:29-42, 305-330) ----
    def align_poses(self, init_keypose, poses):
        T_relative = init_keypose @ np.linalg.inv(poses[0])
        aligned = np.zeros_like(poses)
        aligned[0] = init_keypose
        for i in range(1, poses.shape[0]):
            aligned[i] = T_relative @ poses[i]
        return aligned

    def get_tum_poses(self, init_keypose=None):
        from scipy.spatial.transform import Rotation
        poses = self.get_im_poses().detach().cpu().numpy()
        if init_keypose is not None:
            poses = self.align_poses(np.array(init_keypose), poses)
        tt = np.arange(len(poses)).astype(float)
        tum = []
        for c2w in poses:
            qx, qy, qz, qw = Rotation.from_matrix(c2w[:3, :3]).as_quat()
            tum.append(np.concatenate([c2w[:3, -1], [qw, qx, qy, qz]]))
        return [np.stack(tum, 0), tt]


from . import writeout as _writeout      # noqa: E402   save_tum_poses, save_focals, save_intrinsics, save_conf_maps, ...
for _m in _writeout.METHODS:
    setattr(PointCloudOptimizer, _m.__name__, _m)


def _rgb(ftensor, true_shape=None):
    """reference dust3r/utils/image.py:44-61: network-normalised images -> [0,1] HWC numpy."""
    if isinstance(ftensor, list):
        return [_rgb(x) for x in ftensor]
    if isinstance(ftensor, torch.Tensor):
        ftensor = ftensor.detach().cpu().numpy()
    if ftensor.ndim == 3 and ftensor.shape[0] == 3:
        ftensor = ftensor.transpose(1, 2, 0)
    elif ftensor.ndim == 4 and ftensor.shape[1] == 3:
        ftensor = ftensor.transpose(0, 2, 3, 1)
    if ftensor.dtype == np.uint8:
        img = np.float32(ftensor) / 255
    else:
        img = (ftensor * 0.5) + 0.5
    return img.clip(min=0, max=1)


# =========================================================================================
# optimisation loop (reference base_opt.py:513-583)
# =========================================================================================
def global_alignment_loop(net, lr=0.01, niter=300, schedule='cosine', lr_min=1e-3, temporal_smoothing_weight=0,
                          depth_map_save_dir=None):
    """Adam loop of the global alignment.  ``temporal_smoothing_weight`` is accepted and ignored exactly as in
    the reference (base_opt.py:513,550): the constructor attribute is what counts.  ``depth_map_save_dir`` (verbose runs
    only, as in the reference, base_opt.py:531-540): jet-coloured pictures of the depth maps before iterations 0, 500, ..."""
    trainable = net._trainable()
    if not any(trainable.values()):
        return net
    if net.verbose:
        print('Global alignement - optimizing for:')
        print([name for name, on in trainable.items() if on])
    niter = int(niter)
    if niter <= 0:
        return float('inf')
    net._refresh_hyper()
    b = net._bufs
    b.reset_adam()                      # a fresh torch.optim.Adam per call (reference base_opt.py:524)
    b.ensure_loss_hist(niter)
    lrs = [float(schedule_lr(schedule, it, niter, lr, lr_min)) for it in range(niter)]
    modes = [(_lib.M_FLOW_ACTIVE if net._flow_active(it) else 0) | _lib.M_APPLY_ADAM for it in range(niter)]
    steps = adam_schedule(lrs, start_step=0, betas=(0.9, 0.9), first_iter=0)
    marr = modes_array(modes)
    sharded = net._shard is not None and int(net._shard['world']) > 1
    peers = net._setup_peers() if sharded else None
    chunk = niter if not (net.verbose or (sharded and peers is None)) else (1 if sharded and peers is None else max(1, min(50, niter // 10 or 1)))

    bar = None
    if net.verbose:
        import tqdm
        bar = tqdm.tqdm(total=niter)
    with torch.cuda.device(net._device):
        net._join_uploads()
        net._check_prior()
        st = b.struct(trainable)
        s = net._stream()
        it = 0
        dump = net.verbose and depth_map_save_dir is not None
        while it < niter:
            n = min(chunk, niter - it)
            if dump:
                if it % 500 == 0:
                    _writeout.save_depthmap_pictures(net.get_depthmaps(), depth_map_save_dir, it)
                    print(f"Saved depthmaps at iteration {it} to {depth_map_save_dir}")
                n = min(n, 500 - it % 500)                      # the next picture is due on a chunk boundary
            if sharded and peers is not None:
                sub_m = (C.c_uint32 * n).from_buffer(marr, 4 * it)
                sub_s = (_lib.A3rAdamStep * n).from_buffer(steps, C.sizeof(_lib.A3rAdamStep) * it)
                _lib.check(net._lib.a3r_ga_run_peers(C.byref(st), C.byref(peers), net._peer_tag, n, sub_m, sub_s, s), "a3r_ga_run_peers")
                net._peer_tag += n
            elif sharded:
                for k in range(it, it + n):
                    _lib.check(net._lib.a3r_ga_prepare(C.byref(st), marr[k], s), "a3r_ga_prepare")
                    _lib.check(net._lib.a3r_ga_accumulate(C.byref(st), marr[k], s), "a3r_ga_accumulate")
                    net._reduce_accumulators()
                    _lib.check(net._lib.a3r_ga_update(C.byref(st), marr[k], C.byref(steps[k]), s), "a3r_ga_update")
            else:
                sub_m = (C.c_uint32 * n).from_buffer(marr, 4 * it)
                sub_s = (_lib.A3rAdamStep * n).from_buffer(steps, C.sizeof(_lib.A3rAdamStep) * it)
                _lib.check(net._lib.a3r_ga_run(C.byref(st), n, sub_m, sub_s, s), "a3r_ga_run")
            it += n
            if bar is not None:
                loss = float(b.loss_hist[it - 1].item())
                bar.set_postfix_str(f'lr={lrs[it - 1]:g} loss={loss:g}')
                bar.update(n)
    if bar is not None:
        bar.close()
    net._adam_step = niter
    net._eval_version += 1
    return float(b.loss_hist[niter - 1].item())
