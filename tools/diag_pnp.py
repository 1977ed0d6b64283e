"""Diagnostic: cv2.solvePnPRansac against the device-assisted RANSAC-PnP (cloud_opt_flow/pnp_ransac.py) on one 288x512 frame."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, cv2, torch
from align3r_b200.cloud_opt_flow import pnp_ransac as pr
from tests.test_pnp_ransac import scene
for seed, outl in ((0, 0.0), (1, 0.1), (2, 0.3)):
    X, msk, K = scene(288, 512, 0.01, outl, seed)
    pix = np.mgrid[:512, :288].T.astype(np.float32)
    for it in (10, 100):
        t = time.perf_counter(); ref = cv2.solvePnPRansac(X[msk], pix[msk], K, None, iterationsCount=it, reprojectionError=5, flags=cv2.SOLVEPNP_SQPNP); ta = time.perf_counter() - t
        pr.solve_pnp_ransac_device(X, msk, K, "cuda", iterations=it)
        torch.cuda.synchronize(); t = time.perf_counter(); got = pr.solve_pnp_ransac_device(X, msk, K, "cuda", iterations=it); tb = time.perf_counter() - t
        same = len(ref[3]) == got[3] and np.array_equal(ref[1], got[1]) and np.array_equal(ref[2], got[2])
        print(f"outliers {outl:.1f} iterations<={it:3d}: cv2 {ta * 1e3:6.1f} ms   device-assisted {tb * 1e3:6.1f} ms   inliers {got[3]}/{int(msk.sum())}  identical={same}", flush=True)
