"""The restated RANSAC loop of cv2.solvePnPRansac (align3r_b200/cloud_opt_flow/pnp_ransac.py: OpenCV's generator, subset draw,
loop shortening; EPnP on five points and the final refit through cv2.solvePnP) against cv2.solvePnPRansac itself, with the
per-hypothesis inlier test - the part the product runs on the GPU (a3r_pnp_inliers) - restated in numpy, operation by
operation as the kernel does it.  Equality is exact: same inlier indices, same rvec / tvec bit for bit."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from align3r_b200.cloud_opt_flow import pnp_ransac as pr  # noqa: E402


def numpy_inlier_test(flat, mflat, W, K, reproj):
    """What k_pnp_inliers computes (csrc/mst_kernels.cu): fp64 products and sums in OpenCV's order, fp32 projection, fp32 error."""
    K64 = np.asarray(K, dtype=np.float64)
    fx, fy, cx, cy = K64[0, 0], K64[1, 1], K64[0, 2], K64[1, 2]
    X, Y, Z = (flat[:, k].astype(np.float64) for k in range(3))
    pix = np.arange(flat.shape[0])
    iu, iv = (pix % W).astype(np.float32), (pix // W).astype(np.float32)
    t2 = np.float64(reproj) * np.float64(reproj)
    keep = {}

    def count(rvec, tvec):
        R = cv2.Rodrigues(rvec)[0]
        t = tvec.ravel()
        x = R[0, 0] * X + R[0, 1] * Y + R[0, 2] * Z + t[0]
        y = R[1, 0] * X + R[1, 1] * Y + R[1, 2] * Z + t[1]
        z = R[2, 0] * X + R[2, 1] * Y + R[2, 2] * Z + t[2]
        with np.errstate(divide="ignore"):
            iz = np.where(z != 0, 1.0 / z, 1.0)
        pu = ((x * iz) * fx + cx).astype(np.float32)
        pv = ((y * iz) * fy + cy).astype(np.float32)
        dx, dy = iu - pu, iv - pv
        err = dx * dx + dy * dy
        flags = mflat & (err.astype(np.float64) <= t2)
        token = len(keep)
        keep[token] = flags
        return int(flags.sum()), token
    return count, keep


def restated(pts3d, msk, K, iterations, reproj, flags):
    H, W, _ = pts3d.shape
    flat = np.ascontiguousarray(pts3d, dtype=np.float32).reshape(-1, 3)
    mflat = np.asarray(msk).reshape(-1)
    sel = np.flatnonzero(mflat)
    count, keep = numpy_inlier_test(flat, mflat, W, K, reproj)
    best, model = pr.ransac_loop(pr._Lazy(flat, sel), pr._LazyPix(sel, W), K, count, iterations)
    if best is None:
        return False, None, None, None
    pix = np.flatnonzero(keep[best])
    rvec, tvec = pr.refit(flat[pix], pr._pixels(pix, W), K, model, flags)
    return True, rvec, tvec, np.searchsorted(sel, pix)           # indices in the solver's list, as cv2 returns them


def restated_batched(pts3d, msk, K, iterations, reproj, flags, batch):
    """The batched form of the loop (what run_on_frame drives on the device), with the numpy inlier test."""
    H, W, _ = pts3d.shape
    frame = pr.HostFrame(pts3d, msk, "cpu")
    mflat = np.asarray(msk).reshape(-1)
    count, keep = numpy_inlier_test(frame.flat, mflat, W, K, reproj)

    def count_batch(models):
        return [count(rvec, tvec)[0] for rvec, tvec in models]
    model = pr.ransac_loop_batched(frame.rows_batched, K, frame.count, count_batch, iterations, batch=batch)
    if model is None:
        return False, None, None, None
    n, token = count(*model)
    pix = np.flatnonzero(keep[token])
    rvec, tvec = pr.refit(frame.flat[pix], pr._pixels(pix, W), K, model, flags)
    return True, rvec, tvec, np.searchsorted(frame.sel, pix)


def scene(H, W, noise, outliers, seed):
    r = np.random.default_rng(seed)
    u, v = np.meshgrid(np.arange(W), np.arange(H))
    f = 0.6 * W + 0.5 * W * r.random()
    z = 2 + 0.5 * np.sin(u / (0.1 * W) + r.random()) + 0.3 * np.cos(v / (0.14 * H))
    X = np.stack([(u - W / 2) / f * z, (v - H / 2) / f * z, z], -1).astype(np.float32)
    R = cv2.Rodrigues(r.normal(0, 0.3, 3))[0].astype(np.float32)
    X = (X @ R.T + r.normal(0, 1, 3).astype(np.float32)).astype(np.float32)
    X += r.normal(0, noise, X.shape).astype(np.float32)
    bad = r.random((H, W)) < outliers
    X[bad] += r.normal(0, 1.0, (int(bad.sum()), 3)).astype(np.float32)
    msk = r.random((H, W)) > 0.1
    fk = f * (0.8 + 0.4 * r.random())
    K = np.float32([(fk, 0, W / 2), (0, fk, H / 2), (0, 0, 1)])
    return X, msk, K


@pytest.mark.parametrize("seed", range(12))
def test_restated_loop_equals_opencv(seed):
    H, W = ((96, 160), (48, 64), (40, 56))[seed % 3]
    X, msk, K = scene(H, W, 0.01 * (1 + seed % 3), 0.06 * (seed % 5), seed)
    iterations = 10 if seed % 2 == 0 else 100        # init_im_poses.py:498 / pair_viewer.py:52
    pix = np.mgrid[:W, :H].T.astype(np.float32)
    ok, rvec, tvec, inl = cv2.solvePnPRansac(X[msk], pix[msk], K, None, iterationsCount=iterations, reprojectionError=5,
                                             flags=cv2.SOLVEPNP_SQPNP)
    ok2, rvec2, tvec2, inl2 = restated(X, msk, K, iterations, 5, cv2.SOLVEPNP_SQPNP)
    assert bool(ok) == bool(ok2)
    if ok:
        assert np.array_equal(inl.ravel(), inl2)
        assert np.array_equal(rvec, rvec2) and np.array_equal(tvec, tvec2)
    for batch in (1, 4, 8):          # hypotheses tested several at a time, OpenCV's decisions replayed on the counts
        ok3, rvec3, tvec3, inl3 = restated_batched(X, msk, K, iterations, 5, cv2.SOLVEPNP_SQPNP, batch)
        assert bool(ok) == bool(ok3)
        if ok:
            assert np.array_equal(inl.ravel(), inl3), batch
            assert np.array_equal(rvec, rvec3) and np.array_equal(tvec, tvec3), batch


def test_generator_and_loop_shortening():
    rng = pr.CvRNG()
    first = [rng.next() for _ in range(3)]
    # cv::RNG((uint64)-1): state' = (uint32)state * 4164903690 + (state >> 32)
    s = 0xFFFFFFFFFFFFFFFF
    want = []
    for _ in range(3):
        s = ((s & 0xFFFFFFFF) * 4164903690 + (s >> 32)) & 0xFFFFFFFFFFFFFFFF
        want.append(s & 0xFFFFFFFF)
    assert first == want
    idx = pr.draw_subset(pr.CvRNG(), 7)
    assert len(set(idx)) == 5 and all(0 <= k < 7 for k in idx)
    assert pr.update_num_iters(0.99, 0.0, 5, 10) == 0             # every point an inlier: the loop ends
    assert pr.update_num_iters(0.99, 0.5, 5, 10) == 10
    assert pr.update_num_iters(0.99, 0.2, 5, 100) == 12           # log(0.01) / log(1 - 0.8^5) = 11.6


def test_solve_without_a_device_is_opencv():
    X, msk, K = scene(48, 64, 0.01, 0.1, 3)
    pix = np.mgrid[:64, :48].T.astype(np.float32)
    ok, rvec, tvec, inl = cv2.solvePnPRansac(X[msk], pix[msk], K, None, iterationsCount=10, reprojectionError=5, flags=cv2.SOLVEPNP_SQPNP)
    ok2, rvec2, tvec2, n2 = pr.solve(X, msk, K, None, iterations=10, reproj=5)
    assert ok and ok2 and n2 == len(inl) and np.array_equal(rvec, rvec2) and np.array_equal(tvec, tvec2)
