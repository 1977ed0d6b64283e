"""Spanning-tree initialisation of the global alignment (reference dust3r/cloud_opt_flow/init_im_poses.py).

The reference grows the tree with N-2 *sequential* weighted Procrustes solves on (H,W,3) CPU tensors and
then runs E more, each followed by host syncs (init_im_poses.py:107-154,219-262).  Here the graph logic
(scores -> scipy MST -> traversal order, including the reference's stale-variable focal quirk at
init_im_poses.py:223-224) stays on the host, but every pointmap reduction is one batched kernel launch:

  * world pointmaps are never materialised: pts3d[n] = T_n o X[src_n] is kept symbolically as a
    similarity T_n (fp64 on the host) and the packed slot src_n it applies to;
  * similarity-equivariance of the registration (register(x -> T o y) = T o register(x -> y)) turns all
    E + N-2 solves into slot-to-slot registrations that do not depend on each other
    -> a3r_procrustes_batched, composed afterwards in tree order;
  * focals -> a3r_weiszfeld_focal_batched, initial log-depth planes -> a3r_init_depth;
  * RANSAC-PnP for frames the tree leaves without a pose stays cv2 on the host, as in the reference.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import scipy.sparse as sp
import torch

from .. import _lib
from .._timing import Phases, host_workers
from .commons import edge_str, signed_log1p


def _allreduce_np(self, arr: np.ndarray) -> np.ndarray:
    """Sum a small host array over the frame-window shards (identity on one GPU)."""
    if self._shard is None or int(self._shard['world']) <= 1:
        return arr
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(arr)).to(self.device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self._shard.get('group'))
    return t.cpu().numpy()


def compute_edge_scores(self):
    """score(e) = mean(conf_i[e]) * mean(conf_j[e]) (reference commons.py:21-31) from the per-slot
    confidence sums the pack kernel accumulated."""
    E = self.n_edges
    sums = np.zeros(2 * E)
    local = self._conf_sum.cpu().numpy()
    for gs, s in self._lslot.items():
        sums[gs] = local[s]
    sums = _allreduce_np(self, sums)
    areas = np.array([h * w for h, w in self.imshapes], dtype=np.float64)
    scores = {}
    for e, (i, j) in enumerate(self.edges):
        mi = np.float32(sums[2 * e] / areas[i])
        mj = np.float32(sums[2 * e + 1] / areas[j])
        scores[(i, j)] = float(mi * mj)
    return scores


def _is_rank0(self):
    return self._shard is None or int(self._shard.get('rank', 0)) == 0


def edge_score_matrix(edge_scores, n_imgs=None):
    """(n_imgs, n_imgs) matrix of the directed edge scores, NaN where the graph has no edge - what the reference's
    draw_edge_scores_map (init_im_poses.py:24-34) renders."""
    if n_imgs is None:
        n_imgs = max(max(edge) for edge in edge_scores) + 1
    m = np.full((n_imgs, n_imgs), np.nan)
    for (i, j), score in edge_scores.items():
        m[i, j] = score
    return m


def save_edge_scores_map(edge_scores, save_path, n_imgs=None):
    """reference draw_edge_scores_map (init_im_poses.py:24-42): the annotated seaborn heat-map of the score matrix at
    ``save_path``.  The data behind the picture is always written next to it (``<save_path minus .png>.npy``); the picture
    itself needs matplotlib + seaborn, which this image does not have - without them only the matrix is written."""
    import os
    m = edge_score_matrix(edge_scores, n_imgs)
    np.save(os.path.splitext(save_path)[0] + ".npy", m)
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
        import seaborn as sns
    except ImportError:
        return m
    n = m.shape[0]
    fig = plt.figure(figsize=(int(5.5 * np.log(n) - 2), int((5.5 * np.log(n) - 2) * 3 / 4)))
    sns.heatmap(m, annot=True, fmt=".1f", cmap="viridis", cbar=True, annot_kws={"fontsize": int(-4.2 * np.log(n) + 22.4)})
    plt.title("Heatmap of Edge Scores")
    plt.xlabel("Node")
    plt.ylabel("Node")
    plt.savefig(save_path)
    plt.close(fig)
    return m


def dict_to_sparse_graph(dic):
    """reference init_im_poses.py:438-443."""
    n_imgs = max(max(e) for e in dic) + 1
    res = sp.dok_array((n_imgs, n_imgs))
    for edge, value in dic.items():
        res[edge] = value
    return res


def _compose(a, b):
    """Similarity composition a o b for (s, R, t) triples: x -> s_a R_a (s_b R_b x + t_b) + t_a."""
    sa, Ra, ta = a
    sb, Rb, tb = b
    return sa * sb, Ra @ Rb, sa * (Ra @ tb) + ta


def _umeyama(x, y):
    """Unweighted similarity registration y ~ s R x + t (roma.rigid_points_registration(compute_scaling=True)), fp64."""
    xm, ym = x.mean(0), y.mean(0)
    xc, yc = x - xm, y - ym
    U, D, Vt = np.linalg.svd(yc.T @ xc)
    d = np.sign(np.linalg.det(U @ Vt))
    R = U @ np.diag([1.0, 1.0, d]) @ Vt
    s = (D[0] + D[1] + d * D[2]) / (xc ** 2).sum()
    return float(s), R, ym - s * (R @ xm)


def align_multiple_poses(src_poses, target_poses):
    """reference init_im_poses.py:534-542 on (N,4,4) numpy cam-to-world poses -> (s, R, t)."""
    from scipy.spatial.distance import pdist

    def center_and_z(P):
        eps = float(np.median(pdist(P[:, :3, 3]))) / 100          # utils/geometry.py:370-372
        return np.concatenate((P[:, :3, 3], P[:, :3, 3] + eps * P[:, :3, 2]))

    return _umeyama(center_and_z(np.asarray(src_poses, dtype=np.float64)), center_and_z(np.asarray(target_poses, dtype=np.float64)))


def _rigid4x4(R, t):
    M = np.eye(4)
    M[:3, :3] = R
    M[:3, 3] = t
    return M


def _run_procrustes(self, pairs):
    """pairs: list of (x_gslot, y_gslot) living on one frame.  Returns (len, 13) fp64 [s, R(9), t(3)]
    for every pair (pairs of other shards are filled in by the all-reduce)."""
    out = np.zeros((len(pairs), 13))
    mine = [k for k, (x, y) in enumerate(pairs) if x in self._lslot]
    if mine:
        dev = self.device
        xs = torch.tensor([self._lslot[pairs[k][0]] for k in mine], dtype=torch.int32, device=dev)
        ys = torch.tensor([self._lslot[pairs[k][1]] for k in mine], dtype=torch.int32, device=dev)
        npx = torch.tensor([self.imshapes[self.edges[pairs[k][0] >> 1][pairs[k][0] & 1]][0]
                            * self.imshapes[self.edges[pairs[k][0] >> 1][pairs[k][0] & 1]][1] for k in mine],
                           dtype=torch.int32, device=dev)
        sums = torch.zeros(len(mine), 20, dtype=torch.float64, device=dev)
        res = torch.zeros(len(mine), 16, dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            _lib.check(self._lib.a3r_procrustes_batched(_lib.ptr(self._bufs.predw), self._plan.pix_stride, _lib.ptr(xs),
                                                        _lib.ptr(ys), _lib.ptr(npx), len(mine),
                                                        _lib.CONF_MODES[self.conf_mode], _lib.ptr(sums), _lib.ptr(res),
                                                        self._stream()), "a3r_procrustes_batched")
        out[mine] = res.cpu().numpy()[:, :13]
    return _allreduce_np(self, out)


def _run_focals(self, gslots):
    """Weiszfeld focal of the pointmaps in the given global slots (reference init_im_poses.py:459-465)."""
    out = np.zeros(len(gslots))
    mine = [k for k, g in enumerate(gslots) if g in self._lslot]
    if mine:
        dev = self.device
        sl = torch.tensor([self._lslot[gslots[k]] for k in mine], dtype=torch.int32, device=dev)
        hw = torch.tensor([self.imshapes[self.edges[gslots[k] >> 1][gslots[k] & 1]] for k in mine], dtype=torch.int32, device=dev)
        sums = torch.zeros(len(mine), 2, dtype=torch.float64, device=dev)
        foc = torch.zeros(len(mine), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            _lib.check(self._lib.a3r_weiszfeld_focal_batched(_lib.ptr(self._bufs.predw), self._plan.pix_stride, _lib.ptr(sl),
                                                             _lib.ptr(hw), len(mine), _lib.ptr(sums), _lib.ptr(foc),
                                                             self._stream()), "a3r_weiszfeld_focal_batched")
        out[mine] = foc.cpu().numpy()
    return _allreduce_np(self, out)


def _pixel_grid(H, W):
    return np.mgrid[:W, :H].T.astype(np.float32)


def fast_pnp(pts3d, focal, msk, pp=None, niter_PnP=10, device=None, frame=None):
    """RANSAC-PnP, reference init_im_poses.py:473-513.  Returns (focal, cam2world 4x4) or None.  With a CUDA ``device`` the
    re-projection test of every hypothesis runs there (pnp_ransac.py: the same inlier sets and pose as cv2.solvePnPRansac,
    which is what runs without a device); ``frame``: a pnp_ransac.PlaneFrame instead of the host arrays (points that already
    live on the device)."""
    import cv2
    from . import pnp_ransac
    if frame is None:
        if msk.sum() < 4:
            return None
        H, W, THREE = pts3d.shape
        assert THREE == 3
        if device is not None and pnp_ransac.enabled():
            frame = pnp_ransac.HostFrame(pts3d, msk, device)          # uploaded once for all tentative focals
    else:
        if frame.count < 4:
            return None
        H, W = frame.H, frame.W
    if focal is None:
        S = max(W, H)
        tentative_focals = np.geomspace(S / 2, S * 3, 21)
    else:
        tentative_focals = [focal]
    if pp is None:
        pp = (W / 2, H / 2)
    best = (0,)
    for f in tentative_focals:
        K = np.float32([(f, 0, pp[0]), (0, f, pp[1]), (0, 0, 1)])
        if frame is not None:
            success, R, T, score = pnp_ransac.solve_frame(frame, K, iterations=niter_PnP, reproj=5, flags=cv2.SOLVEPNP_SQPNP)
        else:
            success, R, T, score = pnp_ransac.solve(pts3d, msk, K, None, iterations=niter_PnP, reproj=5, flags=cv2.SOLVEPNP_SQPNP)
        if not success:
            continue
        if success and score > best[0]:
            best = score, R, T, f
    if not best[0]:
        return None
    _, R, T, best_focal = best
    R = cv2.Rodrigues(R)[0]  # world to cam
    return best_focal, np.linalg.inv(_rigid4x4(R, T.ravel()))  # cam to world


@torch.no_grad()
def init_minimum_spanning_tree(self, save_score_path=None, save_score_only=False, init_priors=None, niter_PnP=10):
    """Init all camera poses (image-wise and pairwise poses) given an initial set of pairwise
    estimations - reference init_im_poses.py:88-104,157-284 followed by init_from_pts3d (:107-154)."""
    ph = Phases('mst-init', self.device)
    scores = compute_edge_scores(self)
    ph.mark('scores')
    if save_score_only:                                  # reference init_im_poses.py:94-97: scores only, no initialisation
        save_edge_scores_map(scores, save_score_path)
        return
    N, E = self.n_imgs, self.n_edges
    eid = {ij: e for e, ij in enumerate(self.edges)}

    # ---- spanning tree of the score graph (host) ----
    sparse_graph = -dict_to_sparse_graph(scores)
    msp = sp.csgraph.minimum_spanning_tree(sparse_graph).tocoo()
    todo = sorted(zip(-msp.data, msp.row, msp.col))

    # ---- pass 1: traversal order (pure graph logic) ----
    key_sim, key_pose, fixed_focal = None, None, {}
    if init_priors is None:
        score, i, j = todo.pop()
    else:
        # hierarchical clip driver (reference init_im_poses.py:173-215): the tree starts at an edge touching frame 0 - the
        # key frame - whose pose and focal come from the key-frame alignment: init_priors = [pose 4x4, depth, [focal, ...]]
        key_pose = np.array(init_priors[0]).astype(np.float32).astype(np.float64)
        key_sim = (1.0, key_pose[:3, :3].copy(), key_pose[:3, 3].copy())
        fixed_focal[0] = float(init_priors[2][0])
        score, i, j = None, None, None
        while todo:
            score, i, j = todo.pop()
            if i == 0 or j == 0:
                break
            todo.insert(0, (score, i, j))
    if self.verbose:
        print(f' init edge ({i}*,{j}*) {score=}')
    tree_log = [f'init edge ({i}*,{j}*) {score=}\n']       # the reference's <save_score_path>_tree.txt (:185-190)
    k0 = eid[(int(i), int(j))]
    if key_sim is None or int(i) == 0:
        src = {int(i): 2 * k0, int(j): 2 * k0 + 1}      # frame -> slot its world pointmap is made of
    else:                                                # j is the key frame: the reference reads the mirrored edge (:207-209)
        kr = eid[(int(j), int(i))]
        src = {int(i): 2 * kr + 1, int(j): 2 * kr}
    done = {int(i), int(j)}
    first = (int(i), int(j))
    focal_slot = {} if key_sim is not None else {int(i): 2 * k0}      # frame -> slot whose Weiszfeld focal it takes
    stale_k = k0
    steps = []                                           # (anchor frame, new frame, x slot, new src slot, pose frame)
    msp_edges = [first]
    while todo:
        score, i, j = todo.pop()
        i, j = int(i), int(j)
        if i not in focal_slot and i not in fixed_focal:
            focal_slot[i] = 2 * stale_k                  # the reference's stale ``i_j`` (init_im_poses.py:223-224)
        if i in done:
            if self.verbose:
                print(f' init edge ({i},{j}*) {score=}')
            tree_log.append(f'init edge ({i},{j}*) {score=}\n')
            assert j not in done
            k = eid[(i, j)]
            stale_k = k
            steps.append((i, j, 2 * k, 2 * k + 1, i))     # register pred_i[k] onto pts3d[i]; pts3d[j] from pred_j[k]
            src[j] = 2 * k + 1
            done.add(j)
            msp_edges.append((i, j))
        elif j in done:
            if self.verbose:
                print(f' init edge ({i}*,{j}) {score=}')
            tree_log.append(f'init edge ({i}*,{j}) {score=}\n')
            assert i not in done
            k = eid[(i, j)]
            stale_k = k
            steps.append((j, i, 2 * k + 1, 2 * k, i))     # register pred_j[k] onto pts3d[j]; pts3d[i] from pred_i[k]
            src[i] = 2 * k
            done.add(i)
            msp_edges.append((i, j))
        else:
            todo.insert(0, (score, i, j))
    # leftover focals, best edge first (init_im_poses.py:266-270)
    pair_scores = list(sparse_graph.values())
    edges_from_best_to_worse = np.array(list(sparse_graph.keys()))[np.argsort(pair_scores)]
    for i, j in edges_from_best_to_worse.tolist():
        if i not in focal_slot and i not in fixed_focal:
            focal_slot[i] = 2 * eid[(i, j)]

    if save_score_path is not None and _is_rank0(self):
        save_edge_scores_map(scores, save_score_path, n_imgs=N)
        with open(save_score_path.replace(".png", "_tree.txt"), "w") as f:
            f.writelines(tree_log)
    ph.mark('tree')
    # ---- pass 2: every pointmap reduction in two batched launches ----
    pairs = []
    # tree steps: y = the slot the anchor frame's pointmap is made of *at that time* = src[anchor]
    for anchor, new, xs, new_src, pose_frame in steps:
        pairs.append((xs, src[anchor]))
    for e, (i, j) in enumerate(self.edges):               # init_from_pts3d: pred_i[e] onto pts3d[i]
        pairs.append((2 * e, src[i]))
    reg = _run_procrustes(self, pairs)
    ph.mark('procrustes')
    frames_f = sorted(focal_slot)
    focals = dict(zip(frames_f, _run_focals(self, [focal_slot[n] for n in frames_f]).tolist()))
    focals.update(fixed_focal)

    ph.mark('focals')

    def sim(row):
        return float(row[0]), row[1:10].reshape(3, 3), row[10:13]

    # ---- pass 3: compose in tree order (fp64, host) ----
    ident = (1.0, np.eye(3), np.zeros(3)) if key_sim is None else key_sim
    T = {first[0]: ident, first[1]: ident}
    im_poses = [None] * N
    if key_sim is None:
        im_poses[first[0]] = np.eye(4)
    else:
        im_poses[0] = key_pose.copy()                    # first[0] == 0 or first[1] == 0
    for t, (anchor, new, xs, new_src, pose_frame) in enumerate(steps):
        s, R, tt = _compose(T[anchor], sim(reg[t]))
        T[new] = (s, R, tt)
        if im_poses[pose_frame] is None:
            im_poses[pose_frame] = _rigid4x4(R, tt)          # sRT_to_4x4(1, R, T), init_im_poses.py:241-242,258-259
    im_focals = [focals.get(n) for n in range(N)]

    # ---- frames the tree left without a pose: RANSAC-PnP on the host (init_im_poses.py:272-279).  The calls are
    # independent (cv2 seeds its RANSAC generator per call) and release the GIL: run them on a thread pool.
    pnp = np.zeros((N, 18))          # [solved, cam-to-world (16), focal the solve settled on]
    jobs = []
    from . import pnp_ransac
    on_device = pnp_ransac.enabled()     # the points stay where they are: plane + similarity + mask on the device (PlaneFrame)
    for n in range(N):
        if im_poses[n] is None and n in self._lframe:
            h, w = self.imshapes[n]
            X = self._bufs.predw[self._lslot[src[n]], :h * w]
            msk = self.im_conf[self._lframe[n]] > self.min_conf_thr
            if not on_device:
                # the raw slot plane and the confidence map go to the host as they are; the similarity that makes the
                # world pointmap of frame n is applied by the worker (numpy), next to the cv2 call that consumes it
                X, msk = X.to('cpu', non_blocking=False), msk.cpu()
            jobs.append((n, X, msk, T[n], (h, w)))
    ph.mark(f'pnp-prep({len(jobs)})')
    if jobs:
        from concurrent.futures import ThreadPoolExecutor
        import os as _os

        def _solve(job):
            n, X, msk, (s, R, tt), (h, w) = job
            A, t32 = (s * R).astype(np.float32), tt.astype(np.float32)
            if on_device:         # the same fp32 expression, evaluated by a3r_similarity_points with numpy's rounding
                frame = pnp_ransac.PlaneFrame(X, A, t32, msk, h, w)
                return n, fast_pnp(None, im_focals[n], None, niter_PnP=niter_PnP, frame=frame)
            Xn = X.numpy()
            # plain ufuncs (no BLAS: a threaded GEMM inside every worker would oversubscribe the host)
            pts = (Xn[:, 0:1] * A[:, 0] + Xn[:, 1:2] * A[:, 1] + Xn[:, 2:3] * A[:, 2] + t32).reshape(h, w, 3)
            return n, fast_pnp(pts, im_focals[n], msk.numpy(), niter_PnP=niter_PnP)

        unposed = sum(1 for n in range(N) if im_poses[n] is None)          # the same number on every rank (the tree is global)
        with ThreadPoolExecutor(max_workers=host_workers(len(jobs), self._shard, total_jobs=unposed)) as pool:
            for n, res in pool.map(_solve, jobs):
                if res:
                    pnp[n, 0] = 1
                    pnp[n, 1:17] = res[1].ravel()
                    pnp[n, 17] = res[0]
    ph.mark('pnp-solve')
    pnp = _allreduce_np(self, pnp)
    for n in range(N):
        if im_poses[n] is None:
            im_poses[n] = pnp[n, 1:17].reshape(4, 4) if pnp[n, 0] > 0 else np.eye(4)
            if pnp[n, 0] > 0:
                # a frame that is never the first image of a pair (non-symmetrised graphs) has no Weiszfeld focal: it takes
                # the tentative focal RANSAC-PnP scored best (init_im_poses.py:276-277, 487-504)
                im_focals[n] = float(pnp[n, 17])
    im_poses = np.stack(im_poses)
    res = init_from_pts3d(self, T, src, reg[len(steps):], im_focals, im_poses)
    ph.mark('init_from_pts3d')
    ph.done()
    return res


def init_from_pts3d(self, T, src, edge_reg, im_focals, im_poses):
    """reference init_im_poses.py:107-154 on symbolic world pointmaps (T_n, src_n)."""
    from scipy.spatial.transform import Rotation
    N, E = self.n_imgs, self.n_edges
    if not self.im_poses.requires_grad:
        # every pose was preset and frozen (preset_pose): global Sim(3) alignment of the tree onto the known poses
        # (reference init_im_poses.py:109-122); world pointmaps are symbolic, so "rotate everything" = compose T_n
        if N == 1:
            raise NotImplementedError("Would be simpler to just align everything afterwards on the single known pose")
        known = self.get_im_poses().detach().cpu().double().numpy()
        s, R, t = align_multiple_poses(im_poses, known)
        trf = np.eye(4)
        trf[:3, :3], trf[:3, 3] = s * R, t
        im_poses = trf @ im_poses
        im_poses[:, :3, :3] /= s
        T = {n: _compose((s, R, t), Tn) for n, Tn in T.items()}

    # ---- pairwise poses: register(pred_i[e] -> pts3d[i]) = T_i o register(slot 2e -> src_i) ----
    Rs = np.zeros((E, 3, 3))
    ts = np.zeros((E, 3))
    ss = np.zeros(E)
    for e, (i, j) in enumerate(self.edges):
        row = edge_reg[e]
        ss[e], Rs[e], ts[e] = _compose(T[i], (float(row[0]), row[1:10].reshape(3, 3), row[10:13]))
    pw = np.zeros((E, 8), dtype=np.float32)
    pw[:, 0:4] = Rotation.from_matrix(Rs).as_quat().astype(np.float32)                  # _set_pose, base_opt.py:221
    pw[:, 4:7] = signed_log1p(torch.from_numpy((ts / ss[:, None]).astype(np.float32))).numpy()   # :223
    pw[:, 7] = np.log(ss).astype(np.float32)                                            # :227
    if self.pw_poses.requires_grad:
        self.pw_poses.data.copy_(torch.from_numpy(pw).to(self.device))

    # ---- scale normalisation (init_im_poses.py:131-135) ----
    s_factor = float(self.get_pw_norm_scale_factor()) if self.norm_pw_scale else 1.0
    im_poses = im_poses.copy()
    im_poses[:, :3, 3] *= s_factor

    # ---- depth maps: z of inv(cam2world) . (s_factor T_n X) as one launch (init_im_poses.py:138-142) ----
    plan = self._plan
    nl = plan.n_local_frames
    if nl > 0 and self.im_depthmaps.requires_grad:
        rows = np.zeros((nl, 4))
        srcs = np.zeros(nl, dtype=np.int32)
        npix = np.zeros(nl, dtype=np.int32)
        for k, n in enumerate(plan.lframe_id.tolist()):
            s, R, tt = T[n]
            w2c = np.linalg.inv(im_poses[n])
            rows[k, :3] = s_factor * (w2c[2, :3] @ (s * R))
            rows[k, 3] = s_factor * (w2c[2, :3] @ tt) + w2c[2, 3]
            srcs[k] = self._lslot[src[n]]
            npix[k] = self.imshapes[n][0] * self.imshapes[n][1]
        dev = self.device
        d_rows = torch.from_numpy(rows).to(dev)
        d_src = torch.from_numpy(srcs).to(dev)
        d_np = torch.from_numpy(npix).to(dev)
        with torch.cuda.device(dev):
            _lib.check(self._lib.a3r_init_depth(_lib.ptr(self._bufs.predw), plan.pix_stride, _lib.ptr(d_src), _lib.ptr(d_np),
                                                _lib.ptr(d_rows), nl, _lib.ptr(self.im_depthmaps.data), self._stream()),
                       "a3r_init_depth")

    # ---- image poses + focals (init_im_poses.py:143-148) ----
    po = np.zeros((N, 7), dtype=np.float32)
    po[:, 0:4] = Rotation.from_matrix(im_poses[:, :3, :3]).as_quat().astype(np.float32)
    po[:, 4:7] = signed_log1p(torch.from_numpy(im_poses[:, :3, 3].astype(np.float32))).numpy()
    if self.im_poses.requires_grad:                       # _set_pose only writes parameters that are not preset
        self.im_poses.data.copy_(torch.from_numpy(po).to(self.device))
    if self.im_focals.requires_grad:
        if self.shared_focal:
            self._set_focal(0, sum(im_focals) / N)
        else:
            for n in range(N):
                if im_focals[n] is not None:
                    self._set_focal(n, im_focals[n])
    if N > 2:
        self._set_init_depthmap()
    if self.verbose:
        print(' init loss =', float(self._evaluate(9999)))
    return None


@torch.no_grad()
def init_from_known_poses(self, niter_PnP=10, min_conf_thr=3):
    """init='known_poses' - reference init_im_poses.py:44-85.  Pairwise poses: RANSAC-PnP pose of the second view of
    every pair (on the host, as in the reference) aligned with the two known cameras; depth maps: the best-scoring
    pred_i of every frame, rescaled.  Frame-window shards: the rank owning a pair's FIRST frame solves the pair (its mask and
    its depth candidate live there); the pair's pred_j plane is fetched from the neighbour when the pair straddles a window
    boundary, and the E pose rows are summed across ranks."""
    from concurrent.futures import ThreadPoolExecutor
    from scipy.spatial.transform import Rotation
    E = self.n_edges
    sharded = self._shard is not None and int(self._shard['world']) > 1
    known = self.get_im_poses().detach().cpu().double().numpy()
    focals = self.get_focals().detach().cpu().numpy().reshape(self.n_imgs, -1).mean(axis=1)
    pps = self.get_principal_points().detach().cpu().numpy()
    mine = [e for e in range(E) if (2 * e) in self._lslot]                  # pairs whose first frame lives here
    remote = {}
    if sharded:
        remote = self._fetch_remote_planes(self._bufs.predw, [2 * e + 1 for e in mine if (2 * e + 1) not in self._lslot])
    jobs = []
    for e in mine:
        i, j = self.edges[e]
        i_j = edge_str(i, j)
        conf = self.conf_i[i_j].cpu()
        msk = conf > min(min_conf_thr, float(conf.min()) - 0.1)
        if (2 * e + 1) in self._lslot:
            pts = self.pred_j[i_j].cpu().numpy()
        else:
            h, w = self.imshapes[j]
            pts = remote[2 * e + 1][:h * w, :3].reshape(h, w, 3).cpu().numpy()
        jobs.append((e, pts, msk.numpy(), float(focals[i]), tuple(pps[i]), float(conf.mean())))

    def _solve(job):
        e, pts, msk, f, pp, _ = job
        return e, fast_pnp(pts, f, msk, pp=pp, niter_PnP=niter_PnP, device=self.device)

    pw = np.zeros((E, 9), dtype=np.float64)              # 8 pose entries + "solved" flag
    best = {}
    results = {}
    if jobs:
        with ThreadPoolExecutor(max_workers=host_workers(len(jobs), self._shard)) as pool:
            results = dict(pool.map(_solve, jobs))
    failed = [edge_str(*self.edges[e]) for e in mine if results[e] is None]
    for job in jobs:
        e = job[0]
        i, j = self.edges[e]
        res = results[e]
        if res is None:
            continue
        s, R, t = align_multiple_poses(np.stack((np.eye(4), res[1])), known[[i, j]])
        pw[e, 0:4] = Rotation.from_matrix(R).as_quat().astype(np.float32)                       # _set_pose, base_opt.py:221
        pw[e, 4:7] = signed_log1p(torch.from_numpy((t / s).astype(np.float32))).numpy()
        pw[e, 7] = np.float32(np.log(s))
        pw[e, 8] = 1.0
        score = job[5]
        if score > best.get(i, (0,))[0]:
            best[i] = (score, e, s)
    pw = _allreduce_np(self, pw)
    if failed or not np.all(pw[:, 8] == 1.0):
        bad = failed or [edge_str(*self.edges[e]) for e in np.nonzero(pw[:, 8] != 1.0)[0]]
        raise RuntimeError(f"init='known_poses': PnP failed on pair(s) {bad} (the reference fails here too)")
    if self.pw_poses.requires_grad:
        self.pw_poses.data.copy_(torch.from_numpy(pw[:, :8].astype(np.float32)).to(self.device))
    for n in range(self.n_imgs):
        if n in best:
            _, e, scale = best[n]
            i, j = self.edges[e]
            self._set_depthmap(n, self.pred_i[edge_str(i, j)][:, :, 2] * float(scale))
