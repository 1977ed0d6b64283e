"""Self-estimated dynamic masks of the flow term (``use_self_mask=True``).

Reference: ``PointCloudOptimizer.get_motion_mask_from_pairs`` (dust3r/cloud_opt_flow/optimizer.py:156-235) builds, for every
symmetric pair of edges, a two-image ``PairViewer`` (pair_viewer.py:21-78: Weiszfeld focals, one RANSAC-PnP pose, the depth of
both images), warps both images with those cameras (``DepthBasedWarping``, utils/goem_opt.py:195-236), compares the ego flow
with the optical flow, min-max normalises every error map and thresholds the per-frame mean at ``motion_mask_thre``.

Here: the focals come from the batched Weiszfeld kernel and the pair confidences from the pack kernel's sums (both already
used by the spanning-tree initialisation), the E/2 PnP solves run on a host thread pool (cv2, like the reference), and the
per-pixel work - E error maps, their extrema, the per-frame means and the threshold - is two launches over the slot planes
that are already resident for the hot loop (csrc/motion_kernels.cu); the masks are produced on the device, where the hot loop
reads them.  Every error map lives on one frame's pixel grid, so a frame-window shard computes the masks of its own frames.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib
from .._timing import Phases, host_workers

PAIR_MIN_CONF_THR = 3.0      # the PairViewer is built with the base class default (reference base_opt.py:66, optimizer.py:188)


def _pixel_grid(H, W):
    return np.mgrid[:W, :H].T.astype(np.float32)


def _pair_pnp(pts3d, msk, focal, H, W, device=None):
    """pair_viewer.py:44-60: cam-to-world of the image whose points ``pts3d`` are expressed in the other image's frame
    (cv2.solvePnPRansac, or its device-assisted equal: pnp_ransac.py)."""
    import cv2
    from . import pnp_ransac
    K = np.float32([(focal, 0, W / 2), (0, focal, H / 2), (0, 0, 1)])
    try:
        success, R, T, _ = pnp_ransac.solve(pts3d, msk, K, device, iterations=100, reproj=5, flags=cv2.SOLVEPNP_SQPNP)
        assert success
        R = cv2.Rodrigues(R)[0]                                        # world to cam
        pose = np.linalg.inv(np.r_[np.c_[R, T], [(0, 0, 0, 1)]])       # cam to world
    except Exception:
        pose = np.eye(4)
    return pose.astype(np.float32)


def _row(t, r):
    v = t[r]
    return v.numpy() if isinstance(v, torch.Tensor) else np.asarray(v)


def get_motion_mask_from_pairs(self, view1, view2, pred1, pred2):
    """Sets ``self.dynamic_masks`` (list of N bool (H,W) device tensors; None for frames of other shards) and fills the
    static-mask planes of the hot loop.  Same signature as the reference method (optimizer.py:156)."""
    from . import init_im_poses as init_fun
    assert self.is_symmetrized, 'only support symmetric case'
    E = len(self.edges)
    half = E // 2
    # the reference pairs edge k with edge k + E/2 (optimizer.py:158): the symmetrised list layout of make_pairs
    assert all(self.edges[k + half] == (self.edges[k][1], self.edges[k][0]) for k in range(half)), \
        'only support symmetric case: edge k + E/2 must be the mirror of edge k'
    ph = Phases('self-mask', self.device)
    H, W = self.imshape
    dev = self.device
    pts2, conf1, conf2 = pred2['pts3d_in_other_view'], pred1['conf'], pred2['conf']

    focals = init_fun._run_focals(self, [2 * e for e in range(E)])     # Weiszfeld focal of every pred_i pointmap
    scores = init_fun.compute_edge_scores(self)                        # mean(conf_i) * mean(conf_j) per edge
    ph.mark('focals+scores')

    # ---- which PnP each pair needs: the more confident edge's frame is the pair's world (pair_viewer.py:63-71) ----
    mine = [k for k in range(half) if self.edges[k][0] in self._lframe or self.edges[k][1] in self._lframe]
    jobs = []
    for k in mine:
        a, b = k, k + half
        first = scores[self.edges[a]] > scores[self.edges[b]]
        if first:        # image 1 (= j) is posed: its points in image 0's frame are pred_j['0_1'], its focal from pred_i['1_0']
            ra, rb = self._row_of_edge[a], self._row_of_edge[b]
            jobs.append((k, first, ra, np.maximum(_row(conf2, ra), _row(conf1, rb)), focals[b]))
        else:            # image 0 (= i) is posed: pred_j['1_0'], focal from pred_i['0_1']
            ra, rb = self._row_of_edge[a], self._row_of_edge[b]
            jobs.append((k, first, rb, np.maximum(_row(conf1, ra), _row(conf2, rb)), focals[a]))
    rel = {}
    if jobs:
        from concurrent.futures import ThreadPoolExecutor
        import os as _os

        def _solve(job):
            k, first, r, conf, focal = job
            return k, first, _pair_pnp(_row(pts2, r).reshape(H, W, 3), conf.reshape(H, W) > PAIR_MIN_CONF_THR, float(focal), H, W,
                                       device=self.device)

        with ThreadPoolExecutor(max_workers=host_workers(len(jobs), self._shard)) as pool:
            for k, first, pose in pool.map(_solve, jobs):
                rel[k] = (first, pose)
    ph.mark(f'pnp({len(jobs)})')

    # ---- items: two per pair (side i, side j) on the frames this device owns ----
    eye = torch.eye(4)
    item_slot, item_dslot, item_frame, K_s, K_t, P_s, P_t, drow = [], [], [], [], [], [], [], []
    for k in mine:
        a, b = k, k + half
        i, j = self.edges[a]
        first, pose = rel[k]
        pose = torch.from_numpy(pose)
        inv = torch.linalg.inv(pose)
        Ki = torch.tensor([[focals[a], 0, W / 2], [0, focals[a], H / 2], [0, 0, 1]], dtype=torch.float32)
        Kj = torch.tensor([[focals[b], 0, W / 2], [0, focals[b], H / 2], [0, 0, 1]], dtype=torch.float32)
        Pi, Pj = (eye, pose) if first else (pose, eye)
        z = torch.tensor([0.0, 0.0, 1.0, 0.0])
        # depth of image i / j and the slot whose points it is made of (pair_viewer.py:63-71)
        di = (2 * a, z) if first else (2 * b + 1, inv[2])
        dj = (2 * a + 1, inv[2]) if first else (2 * b, z)
        if i in self._lframe:        # side i: flow_ij[a] against the ego flow image 0 -> image 1
            item_slot.append(self._lslot[2 * a]); item_dslot.append(self._lslot[di[0]]); item_frame.append(self._lframe[i])
            K_s.append(Ki); K_t.append(Kj); P_s.append(Pi); P_t.append(Pj); drow.append(di[1])
        if j in self._lframe:        # side j: flow_ji[a] against the ego flow image 1 -> image 0
            item_slot.append(self._lslot[2 * a + 1]); item_dslot.append(self._lslot[dj[0]]); item_frame.append(self._lframe[j])
            K_s.append(Kj); K_t.append(Ki); P_s.append(Pj); P_t.append(Pi); drow.append(dj[1])
    n_items = len(item_slot)
    plan, bufs = self._plan, self._bufs
    nl = plan.n_local_frames
    dyn = torch.zeros((max(nl, 1), plan.pix_stride), dtype=torch.uint8, device=dev)
    if n_items:
        K_s, K_t, P_s, P_t = torch.stack(K_s), torch.stack(K_t), torch.stack(P_s), torch.stack(P_t)
        R_s, T_s, R_t, T_t = P_s[:, :3, :3], P_s[:, :3, 3:], P_t[:, :3, :3], P_t[:, :3, 3:]
        relR = R_t.transpose(1, 2) @ R_s                                # goem_opt.py:149-154
        relt = R_t.transpose(1, 2) @ (T_s - T_t)
        Hm = K_t @ (relR @ torch.linalg.inv(K_s))                       # goem_opt.py:225
        bt = K_t @ relt
        cam = torch.cat((Hm.reshape(n_items, 9), bt.reshape(n_items, 3), torch.stack(drow)), dim=1).float().contiguous()
        # per-frame item lists in the reference's stacking order (pairs in edge order, side i before side j)
        order = sorted(range(n_items), key=lambda q: item_frame[q])     # stable: keeps the pair order inside a frame
        begin = np.zeros(nl + 1, dtype=np.int32)
        for q in item_frame:
            begin[q + 1] += 1
        begin = np.cumsum(begin).astype(np.int32)
        ints = torch.tensor(item_slot + item_dslot + [H, W] * n_items + order + [H, W] * nl, dtype=torch.int32)
        ints = torch.cat((ints, torch.from_numpy(begin))).to(dev)
        o = 0
        t_slot = ints[o:o + n_items]; o += n_items
        t_dslot = ints[o:o + n_items]; o += n_items
        t_hw = ints[o:o + 2 * n_items]; o += 2 * n_items
        t_items = ints[o:o + n_items]; o += n_items
        t_lhw = ints[o:o + 2 * nl]; o += 2 * nl
        t_begin = ints[o:]
        cam_d = cam.to(dev)
        minmax = torch.empty((n_items, 2), dtype=torch.int32, device=dev)
        # the flow planes are packed on the side stream: this stream reads them now
        if getattr(self, '_upload_stream', None) is not None:
            torch.cuda.current_stream(dev).wait_stream(self._upload_stream)
        with torch.cuda.device(dev):
            st = self._stream()
            _lib.check(self._lib.a3r_motion_minmax(_lib.ptr(bufs.predw), _lib.ptr(bufs.flow), plan.pix_stride, _lib.ptr(t_slot),
                                                   _lib.ptr(t_dslot), _lib.ptr(t_hw), _lib.ptr(cam_d), n_items,
                                                   _lib.ptr(minmax), st), "a3r_motion_minmax")
            _lib.check(self._lib.a3r_motion_masks(_lib.ptr(bufs.predw), _lib.ptr(bufs.flow), plan.pix_stride, _lib.ptr(t_slot),
                                                  _lib.ptr(t_dslot), _lib.ptr(cam_d), _lib.ptr(minmax), _lib.ptr(t_begin),
                                                  _lib.ptr(t_items), _lib.ptr(t_lhw), nl, float(self.motion_mask_thre),
                                                  _lib.ptr(bufs.static_mask), _lib.ptr(dyn), None, st), "a3r_motion_masks")
        self._motion_keep = (ints, cam_d, minmax)     # alive until the launches have run
    ph.mark(f'error-maps({n_items})')
    self._pair_cameras = dict(pairs=mine, focals=focals, rel={k: rel[k] for k in mine})      # read-out for tests / diagnostics
    masks = [None] * self.n_imgs
    for k, n in enumerate(plan.lframe_id.tolist()):
        masks[n] = dyn[k, :H * W].view(H, W).bool()
    self.dynamic_masks = masks
    ph.done()
