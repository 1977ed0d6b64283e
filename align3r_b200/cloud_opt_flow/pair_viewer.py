"""Two-image "dummy optimiser" - reference dust3r/cloud_opt_flow/pair_viewer.py:15-127 (GlobalAlignerMode.PairViewer, what the
tools select for two images: tool/demo.py:151-153).  Nothing is optimised: the focals are the Weiszfeld estimates of the two
first-view pointmaps, the more confident edge's first image is the world frame, the other image is posed by one RANSAC-PnP
solve and its depth is the z of its points moved into its own camera.

Built on the PointCloudOptimizer's device planes: focals from the batched Weiszfeld kernel, confidences from the pack kernel's
sums, PnP on the host (cv2, as in the reference), depth through the depth-initialisation kernel; the read-out getters
(`get_im_poses`, `get_focals`, `get_intrinsics`, `get_depthmaps`, `get_pts3d`, `get_masks`, `get_conf`) are the class's own.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib
from .commons import edge_str
from .motion_mask import _pair_pnp, _row
from .optimizer import PointCloudOptimizer


class PairViewer(PointCloudOptimizer):
    def __init__(self, view1, view2, pred1, pred2, **kw):
        from . import init_im_poses as init_fun
        for k in ("flow_loss_weight", "temporal_smoothing_weight", "depth_regularize_weight"):
            kw.pop(k, None)
        super().__init__(view1, view2, pred1, pred2, **kw)
        assert self.is_symmetrized and self.n_edges == 2          # pair_viewer.py:23
        if self._shard is not None and int(self._shard['world']) > 1:
            raise NotImplementedError("PairViewer on a frame-window sharded problem")
        self.has_im_poses = True
        H0W0, H1W1 = self.imshapes[0], self.imshapes[1]
        e01, e10 = self._edge_index[edge_str(0, 1)], self._edge_index[edge_str(1, 0)]
        # pair_viewer.py:31-42: confidence of edge n -> 1-n and focal of image n from its own first-view pointmap
        scores = init_fun.compute_edge_scores(self)
        confs = [scores[(0, 1)], scores[(1, 0)]]
        focals = init_fun._run_focals(self, [2 * e01, 2 * e10])
        self.focals_px = [float(focals[0]), float(focals[1])]
        conf1, conf2, pts2 = pred1['conf'], pred2['conf'], pred2['pts3d_in_other_view']
        r01, r10 = self._row_of_edge[e01], self._row_of_edge[e10]
        eye = np.eye(4, dtype=np.float32)
        if confs[0] > confs[1]:
            # the point cloud is expressed in camera 0; image 1 is posed from its points in that frame, pred_j['0_1']
            H, W = H1W1
            msk = np.maximum(_row(conf2, r01), _row(conf1, r10)).reshape(H, W) > self.min_conf_thr
            rel = _pair_pnp(_row(pts2, r01).reshape(H, W, 3), msk, self.focals_px[1], H, W, device=self.device)
            poses = [eye, rel]
            depth_src = {0: (2 * e01, eye), 1: (2 * e01 + 1, np.linalg.inv(rel))}
        else:
            H, W = H0W0
            msk = np.maximum(_row(conf1, r01), _row(conf2, r10)).reshape(H, W) > self.min_conf_thr
            rel = _pair_pnp(_row(pts2, r10).reshape(H, W, 3), msk, self.focals_px[0], H, W, device=self.device)
            poses = [rel, eye]
            depth_src = {0: (2 * e10 + 1, np.linalg.inv(rel)), 1: (2 * e10, eye)}
        for p in (self.pw_poses, self.pw_adaptors, self.im_depthmaps, self.im_poses, self.im_focals, self.im_pp):
            p.requires_grad_(False)
        for n in (0, 1):
            self._set_pose(self.im_poses, n, torch.from_numpy(np.asarray(poses[n], dtype=np.float32)), force=True)
            if not self.shared_focal:
                self._set_focal(n, self.focals_px[n], force=True)
        if self.shared_focal:
            self._set_focal(0, sum(self.focals_px) / 2, force=True)
        # depth maps = z of the source points in the image's own camera (pair_viewer.py:63-71), one launch
        rows = np.zeros((2, 4))
        srcs = np.zeros(2, dtype=np.int32)
        npix = np.zeros(2, dtype=np.int32)
        for k, n in enumerate(self._plan.lframe_id.tolist()):
            gslot, w2c = depth_src[n]
            rows[k] = np.asarray(w2c, dtype=np.float64)[2]
            srcs[k] = self._lslot[gslot]
            npix[k] = self.imshapes[n][0] * self.imshapes[n][1]
        dev = self.device
        d_rows, d_src, d_np = torch.from_numpy(rows).to(dev), torch.from_numpy(srcs).to(dev), torch.from_numpy(npix).to(dev)
        with torch.cuda.device(dev):
            _lib.check(self._lib.a3r_init_depth(_lib.ptr(self._bufs.predw), self._plan.pix_stride, _lib.ptr(d_src), _lib.ptr(d_np),
                                                _lib.ptr(d_rows), 2, _lib.ptr(self.im_depthmaps.data), self._stream()),
                       "a3r_init_depth")
        torch.cuda.current_stream(dev).synchronize()

    def get_focals(self):
        return super().get_focals().flatten()                 # pair_viewer.py:89-90: shape (2,)

    def get_known_focal_mask(self):
        return torch.tensor([True, True])

    def _set_depthmap(self, idx, depth, force=False):
        if self.verbose:
            print('_set_depthmap is ignored in PairViewer')
        return

    def compute_global_alignment(self, *a, **kw):
        raise NotImplementedError("PairViewer does not optimise anything (reference pair_viewer.py:15-19); the tools never "
                                  "call compute_global_alignment on it (tool/demo.py:159-160)")

    def forward(self, *a, **kw):
        return float('nan')                                    # pair_viewer.py:126-127
