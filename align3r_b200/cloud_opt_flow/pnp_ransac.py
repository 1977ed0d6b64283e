"""RANSAC-PnP with the hypothesis tests on the GPU - what ``cv2.solvePnPRansac`` computes, not an approximation of it.

The reference poses the frames its spanning tree cannot reach (``init_im_poses.py:272-279``), the second views of
``init='known_poses'`` (``:44-85``) and the two-view cameras of the self-estimated masks (``optimizer.py:156-235`` through
``pair_viewer.py``) with ``cv2.solvePnPRansac(..., iterationsCount=10, reprojectionError=5, flags=SOLVEPNP_SQPNP)`` on all
confident pixels of a frame: ~45 ms per call on a 288x512 frame, the largest single item of an end-to-end job on a B200
(12 calls on the 50-frame workload, one per pair - 235 - for the masks).  OpenCV's routine (``calib3d/src/solvepnp.cpp``,
``ptsetreg.cpp``) is, in order:

1. a RANSAC loop with OpenCV's own generator (``cv::RNG`` seeded with ``(uint64)-1``): five distinct point indices,
   an EPnP pose of the five points, the squared re-projection error of EVERY point under that pose, the inlier count; the
   best hypothesis so far shortens the loop (``RANSACUpdateNumIters``);
2. the pose re-fitted to all inliers of the best hypothesis with the requested solver (SQPNP).

Step 1's re-projection of every point for every hypothesis is the time (~8 ms each).  Here the sampling (a port of
``cv::RNG`` and of ``getSubset``), the loop control and the loop shortening are restated in Python, the five-point EPnP
solves and the final refit ARE OpenCV's (``cv2.solvePnP``), and the per-hypothesis inlier test runs on the device
(``a3r_pnp_inliers``: OpenCV's arithmetic operation by operation - fp64 without FMA contraction, fp32 projection, fp32
squared error).  The hypotheses are tested a batch at a time (``ransac_loop_batched``): OpenCV's samples come from a generator
no result feeds back into, so the next subsets are known ahead and OpenCV's decisions can be replayed on the batch's counts.  The inlier sets, and with them the refit, are identical to ``cv2.solvePnPRansac``'s: checked against it
on the CPU (``tests/test_pnp_ransac.py``: the same loop with the inlier test in numpy), on the GPU
(``tests/test_gpu_parity.py::test_device_ransac_pnp_equals_opencv``) and ONCE PER PROCESS on the first real solve
(``_self_check``): on any difference the process falls back to ``cv2.solvePnPRansac`` for good and says so.
``A3R_FAST_PNP=0`` switches the device path off.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
import warnings

import numpy as np

_MODEL_POINTS = 5           # solvepnp.cpp: every solver but P3P / AP3P draws five points and fits them with EPnP


class CvRNG:
    """``cv::RNG`` (multiply-with-carry, ``core/operations.hpp``): ``next()`` and ``uniform(int, int)``."""

    def __init__(self, state: int = 0xFFFFFFFFFFFFFFFF):
        self.state = state & 0xFFFFFFFFFFFFFFFF

    def next(self) -> int:
        s = self.state
        s = ((s & 0xFFFFFFFF) * 4164903690 + (s >> 32)) & 0xFFFFFFFFFFFFFFFF
        self.state = s
        return s & 0xFFFFFFFF

    def uniform(self, a: int, b: int) -> int:
        return a if a == b else int(self.next() % (b - a) + a)


def draw_subset(rng: CvRNG, count: int, model_points: int = _MODEL_POINTS):
    """``RANSACPointSetRegistrator::getSubset`` (ptsetreg.cpp): distinct indices, redrawn on repetition (the PnP callback
    accepts every subset)."""
    idx = []
    for _ in range(model_points):
        k = rng.uniform(0, count)
        while k in idx:
            k = rng.uniform(0, count)
        idx.append(k)
    return idx


def update_num_iters(p: float, ep: float, model_points: int, max_iters: int) -> int:
    """``cv::RANSACUpdateNumIters`` (ptsetreg.cpp)."""
    p = min(max(p, 0.0), 1.0)
    ep = min(max(ep, 0.0), 1.0)
    tiny = float(np.finfo(np.float64).tiny)
    num = max(1.0 - p, tiny)
    denom = 1.0 - (1.0 - ep) ** model_points
    if denom < tiny:
        return 0
    num, denom = float(np.log(num)), float(np.log(denom))
    if denom >= 0 or -num >= max_iters * (-denom):
        return max_iters
    return int(np.rint(num / denom))          # cvRound: to nearest, ties to even


def ransac_loop(obj, img, K, count_inliers, iterations=10, confidence=0.99, on_best=None):
    """The loop of ``RANSACPointSetRegistrator::run`` for the PnP callback.  ``obj[idx]`` (5,3) / ``img[idx]`` (5,2) fp32: rows of
    the list of points handed to the solver, in its order; ``count_inliers(rvec, tvec) -> (count, token)`` tests one hypothesis
    on all points; ``on_best(token)``: the hypothesis just became the best one (its token must stay valid).
    Returns ``(token of the best hypothesis, its (rvec, tvec))`` or ``(None, None)``."""
    import cv2
    count = len(obj)
    rng = CvRNG()
    niters = max(int(iterations), 1)
    best, best_model, max_good = None, None, 0
    it = 0
    while it < niters:
        idx = draw_subset(rng, count)
        it += 1
        ok, rvec, tvec = cv2.solvePnP(obj[idx], img[idx], K, None, flags=cv2.SOLVEPNP_EPNP)
        if not ok:
            continue
        good, token = count_inliers(rvec, tvec)
        if good > max(max_good, _MODEL_POINTS - 1):
            best, best_model, max_good = token, (rvec.copy(), tvec.copy()), good
            if on_best is not None:
                on_best(token)
            niters = update_num_iters(confidence, (count - good) / count, _MODEL_POINTS, niters)
    return best, best_model


def ransac_loop_batched(rows, K, count, count_batch, iterations=10, confidence=0.99, batch=4):
    """The same loop with the hypotheses tested ``batch`` at a time: OpenCV draws its samples from a generator of its own that no
    result feeds back into, so the next ``batch`` subsets are known ahead; their rows are fetched together (``rows(list of
    subsets) -> list of (obj (5,3), img (5,2))``), their EPnP poses solved, ``count_batch(list of (rvec, tvec)) -> counts`` tests them
    in one go, and OpenCV's decisions are replayed on the counts in order - a hypothesis past the point where the loop would have
    ended is ignored.  One device round trip per batch instead of two per hypothesis.  Returns the best ``(rvec, tvec)`` or None."""
    import cv2
    rng = CvRNG()
    niters = max(int(iterations), 1)
    best_model, max_good = None, 0
    it = 0
    while it < niters:
        nb = min(int(batch), niters - it)
        subsets = [draw_subset(rng, count) for _ in range(nb)]
        models = []
        for obj5, img5 in rows(subsets):
            ok, rvec, tvec = cv2.solvePnP(obj5, img5, K, None, flags=cv2.SOLVEPNP_EPNP)
            models.append((rvec, tvec) if ok else None)
        live = [m for m in models if m is not None]
        counts = iter(count_batch(live)) if live else iter(())
        for m in models:
            good = next(counts) if m is not None else None        # (keeps counts aligned with the live models)
            if it >= niters:
                continue
            it += 1
            if m is None:
                continue
            if good > max(max_good, _MODEL_POINTS - 1):
                best_model, max_good = (m[0].copy(), m[1].copy()), int(good)
                niters = update_num_iters(confidence, (count - good) / count, _MODEL_POINTS, niters)
    return best_model


def refit(obj_in, img_in, K, best_model, flags):
    """solvepnp.cpp after the loop: the requested solver on the inliers (as doubles); the loop's model if it fails."""
    import cv2
    ok, rvec, tvec = cv2.solvePnP(np.ascontiguousarray(obj_in, dtype=np.float64), np.ascontiguousarray(img_in, dtype=np.float64),
                                  K, None, flags=flags)
    if not ok:
        rvec, tvec = best_model
    return rvec, tvec


# ---------------------------------------------------------------------------------------------------------------------
_state = {"enabled": os.environ.get("A3R_FAST_PNP", "1") != "0", "checked": False, "checked_plane": False}
_lock = threading.Lock()


def enabled() -> bool:
    return _state["enabled"]


def _pixels(pix, W):
    out = np.empty((pix.size, 2), dtype=np.float32)
    out[:, 0] = pix % W
    out[:, 1] = pix // W
    return out


class _Lazy:
    """rows of the solver's point list without materialising it: list position -> pixel -> point"""

    def __init__(self, flat, sel):
        self.flat, self.sel = flat, sel

    def __len__(self):
        return int(self.sel.size)

    def __getitem__(self, idx):
        return self.flat[self.sel[idx]]


class _LazyPix:
    def __init__(self, sel, W):
        self.sel, self.W = sel, W

    def __getitem__(self, idx):
        return _pixels(self.sel[idx], self.W)


class HostFrame:
    """The points of one frame handed over as host arrays (``pts3d`` (H,W,3) fp32, ``msk`` (H,W) bool): uploaded once, the
    sampled rows of a hypothesis come from the host copy."""

    def __init__(self, pts3d, msk, device):
        import torch
        self.H, self.W, _ = pts3d.shape
        self.P = self.H * self.W
        self.flat = np.ascontiguousarray(pts3d, dtype=np.float32).reshape(self.P, 3)
        mflat = np.ascontiguousarray(msk, dtype=bool).reshape(self.P)
        self.sel = np.flatnonzero(mflat)                     # position in the solver's list -> pixel
        self.count = int(self.sel.size)
        self.dev = torch.device(device)
        self._mflat = mflat
        self.d_pts = self.d_msk = None

    def upload(self):
        import torch
        if self.d_pts is None:
            self.d_pts = torch.from_numpy(self.flat).to(self.dev)
            self.d_msk = torch.from_numpy(self._mflat.view(np.uint8)).to(self.dev)

    def rows(self):
        return _Lazy(self.flat, self.sel), _LazyPix(self.sel, self.W)

    def rows_batched(self, subsets):
        return [(self.flat[self.sel[idx]], _pixels(self.sel[idx], self.W)) for idx in subsets]

    def host_lists(self):
        return self.flat[self.sel], _pixels(self.sel, self.W)


class PlaneFrame:
    """The points of one frame taken from a raw slot plane that already lives on the device (``plane`` (P,4) fp32: X Y Z w) and
    the similarity that makes them world points (``A32`` (3,3), ``t32`` (3,) fp32: pts = X A[:,0] + Y A[:,1] + Z A[:,2] + t, the
    expression init_im_poses.py evaluates in numpy - a3r_similarity_points rounds every operation as numpy does), ``msk`` (P,)
    bool on the device.  Nothing is uploaded: host-to-device copies, however small, queue behind the flow planes still in flight
    from the constructor; the sampled rows of a hypothesis come back through a3r_pnp_gather (indices as kernel arguments)."""

    def __init__(self, plane, A32, t32, msk, H, W):
        import torch
        from .. import _lib
        self.H, self.W, self.P = H, W, H * W
        self.dev = plane.device
        self._plane, self._A, self._t = plane, np.ascontiguousarray(A32, dtype=np.float32), np.ascontiguousarray(t32, dtype=np.float32)
        lib = _lib.load()
        with torch.cuda.device(self.dev):
            st = C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
            self.d_pts = torch.empty((self.P, 3), dtype=torch.float32, device=self.dev)
            sim = (C.c_float * 12)(*([float(v) for v in self._A.reshape(-1)] + [float(v) for v in self._t.reshape(-1)]))
            assert plane.is_contiguous() and plane.dtype == torch.float32 and tuple(plane.shape) == (self.P, 4)
            _lib.check(lib.a3r_similarity_points(_lib.ptr(plane), self.P, sim, _lib.ptr(self.d_pts), st), "a3r_similarity_points")
            self.d_msk = msk.reshape(self.P).to(torch.uint8)
            self.d_sel = torch.nonzero(self.d_msk, as_tuple=False).flatten()        # int64, ascending = the solver's order
            self.count = int(self.d_sel.numel())
            self._out = torch.empty((64, 4), dtype=torch.float32, device=self.dev)
        self._lib, self._cache = lib, (None, None)

    def upload(self):
        pass

    def _gather(self, idx):
        import torch
        from .. import _lib
        key = tuple(idx)
        if self._cache[0] != key:
            arr = (C.c_int32 * len(idx))(*[int(k) for k in idx])
            with torch.cuda.device(self.dev):
                st = C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
                _lib.check(self._lib.a3r_pnp_gather(_lib.ptr(self.d_pts), _lib.ptr(self.d_sel), arr, len(idx), _lib.ptr(self._out), st),
                           "a3r_pnp_gather")
                rows = self._out[:len(idx)].cpu().numpy()
            self._cache = (key, (np.ascontiguousarray(rows[:, :3]), _pixels(rows[:, 3].astype(np.int64), self.W)))
        return self._cache[1]

    def rows(self):
        frame = self

        class _Obj:
            def __len__(self):
                return frame.count

            def __getitem__(self, idx):
                return frame._gather(idx)[0]

        class _Img:
            def __getitem__(self, idx):
                return frame._gather(idx)[1]

        return _Obj(), _Img()

    def rows_batched(self, subsets):
        """the sampled rows of up to 12 hypotheses in one gather + one download"""
        flat = [k for idx in subsets for k in idx]
        rows = self._gather_flat(flat)
        out = []
        for j in range(len(subsets)):
            r = rows[j * _MODEL_POINTS:(j + 1) * _MODEL_POINTS]
            out.append((np.ascontiguousarray(r[:, :3]), _pixels(r[:, 3].astype(np.int64), self.W)))
        return out

    def _gather_flat(self, flat):
        import torch
        from .. import _lib
        assert len(flat) <= 64
        arr = (C.c_int32 * len(flat))(*[int(k) for k in flat])
        with torch.cuda.device(self.dev):
            st = C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
            _lib.check(self._lib.a3r_pnp_gather(_lib.ptr(self.d_pts), _lib.ptr(self.d_sel), arr, len(flat), _lib.ptr(self._out), st),
                       "a3r_pnp_gather")
            return self._out[:len(flat)].cpu().numpy()

    def host_lists(self):
        """(self-check only) the solver's lists as the host path builds them: numpy's transform of the downloaded plane."""
        Xn = self._plane.cpu().numpy()
        A, t32 = self._A, self._t
        pts = (Xn[:, 0:1] * A[:, 0] + Xn[:, 1:2] * A[:, 1] + Xn[:, 2:3] * A[:, 2] + t32)
        sel = np.flatnonzero(self.d_msk.cpu().numpy())
        return np.ascontiguousarray(pts[sel], dtype=np.float32), _pixels(sel, self.W)


def run_on_frame(frame, K, iterations=10, reproj=5.0, confidence=0.99, flags=None):
    """``cv2.solvePnPRansac(points[msk], pixels[msk], K, None, iterationsCount, reprojectionError, flags)`` on a prepared frame
    (HostFrame / PlaneFrame).  Returns ``(success, rvec, tvec, n_inliers)``."""
    import cv2
    import torch
    from .. import _lib
    if flags is None:
        flags = cv2.SOLVEPNP_SQPNP
    if frame.count <= _MODEL_POINTS:                         # (OpenCV's special cases: not what a frame of pixels is)
        obj, img = frame.host_lists()
        ok, rvec, tvec, inl = cv2.solvePnPRansac(obj, img, K, None, iterationsCount=iterations, reprojectionError=reproj, flags=flags)
        return ok, rvec, tvec, (0 if inl is None else len(inl))
    lib = _lib.load()
    dev, P, W = frame.dev, frame.P, frame.W
    K64 = np.asarray(K, dtype=np.float64)
    frame.upload()
    batch = 4 if int(iterations) <= 16 else 8
    with torch.cuda.device(dev):
        d_flags = torch.empty(P, dtype=torch.uint8, device=dev)
        d_counts = torch.zeros(16, dtype=torch.int32, device=dev)
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        hyp = (C.c_double * 17)()
        hyp[12], hyp[13], hyp[14], hyp[15], hyp[16] = K64[0, 0], K64[1, 1], K64[0, 2], K64[1, 2], float(reproj) * float(reproj)

        def launch(model, flags_ptr, j):
            R = cv2.Rodrigues(model[0])[0]
            for k in range(9):
                hyp[k] = float(R.flat[k])
            for k in range(3):
                hyp[9 + k] = float(model[1].flat[k])
            _lib.check(lib.a3r_pnp_inliers(_lib.ptr(frame.d_pts), _lib.ptr(frame.d_msk), W, P, hyp, flags_ptr,
                                           C.c_void_p(d_counts.data_ptr() + 4 * j), stream), "a3r_pnp_inliers")

        def count_batch(models):                             # one launch per hypothesis, ONE download for the batch
            for j, m in enumerate(models):
                launch(m, None, j)
            return d_counts[:len(models)].cpu().tolist()

        best_model = ransac_loop_batched(frame.rows_batched, K, frame.count, count_batch, iterations, confidence, batch=batch)
        if best_model is None:
            return False, None, None, 0
        launch(best_model, _lib.ptr(d_flags), 0)             # the inlier set of the best hypothesis
        inl = torch.nonzero(d_flags, as_tuple=False).flatten()                # pixel order = the solver's order
        obj_in = frame.d_pts.index_select(0, inl).cpu().numpy()
        pix = inl.cpu().numpy()
    rvec, tvec = refit(obj_in, _pixels(pix, W), K, best_model, flags)
    return True, rvec, tvec, int(pix.size)


def solve_pnp_ransac_device(pts3d, msk, K, device, iterations=10, reproj=5.0, confidence=0.99, flags=None):
    """The device-assisted solve on host arrays: ``pts3d`` (H,W,3) fp32, ``msk`` (H,W) bool, pixel (u, v) observed at (u, v)."""
    return run_on_frame(HostFrame(pts3d, msk, device), K, iterations, reproj, confidence, flags)


def _reference(frame, K, iterations, reproj, flags):
    import cv2
    obj, img = frame.host_lists()
    ok, rvec, tvec, inl = cv2.solvePnPRansac(obj, img, K, None, iterationsCount=iterations, reprojectionError=reproj, flags=flags)
    return ok, rvec, tvec, (0 if inl is None else len(inl))


def solve_frame(frame, K, iterations=10, reproj=5.0, flags=None):
    """The call sites' entry for a prepared frame: the device-assisted loop once it has passed its self-check (one per frame
    kind and process), ``cv2.solvePnPRansac`` on the frame's host lists otherwise."""
    import cv2
    if flags is None:
        flags = cv2.SOLVEPNP_SQPNP
    key = "checked_plane" if isinstance(frame, PlaneFrame) else "checked"
    if _state["enabled"]:
        if not _state[key]:
            with _lock:
                if not _state[key]:
                    _self_check(frame, K, iterations, reproj, flags, key)
        if _state["enabled"]:
            return run_on_frame(frame, K, iterations, reproj, flags=flags)
    return _reference(frame, K, iterations, reproj, flags)


def solve(pts3d, msk, K, device, iterations=10, reproj=5.0, flags=None):
    """Host arrays in: the device-assisted loop when a device is given and the path is on, ``cv2.solvePnPRansac`` otherwise.
    Returns ``(success, rvec, tvec, n_inliers)``."""
    import cv2
    if flags is None:
        flags = cv2.SOLVEPNP_SQPNP
    if device is not None and _state["enabled"]:
        return solve_frame(HostFrame(pts3d, msk, device), K, iterations, reproj, flags)
    H, W, _ = pts3d.shape
    sel = np.flatnonzero(np.asarray(msk).reshape(-1))
    ok, rvec, tvec, inl = cv2.solvePnPRansac(np.ascontiguousarray(pts3d, dtype=np.float32).reshape(-1, 3)[sel], _pixels(sel, W), K, None,
                                             iterationsCount=iterations, reprojectionError=reproj, flags=flags)
    return ok, rvec, tvec, (0 if inl is None else len(inl))


def _self_check(frame, K, iterations, reproj, flags, key):
    """Once per process and frame kind, on the first real solve: the device-assisted loop against cv2.solvePnPRansac on the same
    points.  Anything but the same success flag, inlier count and pose (bit for bit) switches the device path off for the process."""
    ref = _reference(frame, K, iterations, reproj, flags)
    try:
        got = run_on_frame(frame, K, iterations, reproj, flags=flags)
        same = bool(ref[0]) == bool(got[0]) and (not ref[0] or (ref[3] == got[3] and np.array_equal(ref[1], got[1])
                                                                and np.array_equal(ref[2], got[2])))
    except Exception as exc:          # noqa: BLE001 - any failure of the fast path means: do not use it
        same = False
        warnings.warn(f"align3r_b200: device RANSAC-PnP raised {exc!r}; falling back to cv2.solvePnPRansac")
    if not same:
        warnings.warn("align3r_b200: device RANSAC-PnP does not reproduce cv2.solvePnPRansac on this installation "
                      "(another OpenCV build?); falling back to cv2.solvePnPRansac for this process")
        _state["enabled"] = False
    _state[key] = True
