"""ORACLE -- TEST INFRASTRUCTURE ONLY (imported by tests/ only; nothing under spann3r_b200/ touches it).

CPU restatement (numpy) of the camera-pose step of demo.py:166-180: `cv2.solvePnPRansac(pts3d.reshape(-1, 3),
pixel grid, intrinsic, zeros(4))`.  OpenCV is a third-party dependency the reference leaves unpinned
(requirements.txt; 4.13.0 in this image); its published algorithm (calib3d/solvepnp.cpp) is RANSAC over minimal pose
hypotheses scored by the reprojection error (threshold 8 px, <= 100 iterations), followed by a non-linear least-squares
refinement (SOLVEPNP_ITERATIVE, Levenberg-Marquardt on the reprojection error) of the best model on its inliers.  This
restatement deliberately uses a DIFFERENT minimal solver than the CUDA path (6-point DLT + SVD here, Grunert P3P there):
both must land on the same least-squares optimum.

Parity pin: `tests/golden/pnp.json` holds rvec / tvec / inlier counts produced by the REAL cv2 call on seeded synthetic
pointmaps (`tools/make_golden_pnp.py`); `tests/test_pnp.py` holds this oracle to those values (<= 3e-4 absolute: two RANSAC
runs differ by a handful of threshold-straddling inliers; <= 1e-6 on the outlier-free case).
"""
from __future__ import annotations

import numpy as np


def rodrigues(rvec):
    r = np.asarray(rvec, np.float64)
    th = np.linalg.norm(r)
    if th < 1e-12:
        return np.eye(3)
    k = r / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K


def rodrigues_inv(R):
    th = np.arccos(np.clip((np.trace(R) - 1) / 2, -1, 1))
    a = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    if th < 1e-8:
        return 0.5 * a
    return th / (2 * np.sin(th)) * a


def project(R, t, K, X):
    xc = X @ R.T + t
    z = xc[:, 2]
    with np.errstate(divide="ignore", invalid="ignore"):
        uv = np.stack((K[0, 0] * xc[:, 0] / z + K[0, 2], K[1, 1] * xc[:, 1] / z + K[1, 2]), -1)
    return uv, z


def dlt_pose(X, uv, K):
    """6+ point DLT: P = [R|t] up to scale from normalised image coordinates, R projected onto SO(3)."""
    xn = (uv[:, 0] - K[0, 2]) / K[0, 0]
    yn = (uv[:, 1] - K[1, 2]) / K[1, 1]
    n = len(X)
    A = np.zeros((2 * n, 12))
    Xh = np.concatenate((X, np.ones((n, 1))), 1)
    A[0::2, 0:4] = Xh
    A[0::2, 8:12] = -xn[:, None] * Xh
    A[1::2, 4:8] = Xh
    A[1::2, 8:12] = -yn[:, None] * Xh
    _, _, vt = np.linalg.svd(A)
    P = vt[-1].reshape(3, 4)
    if np.linalg.det(P[:, :3]) < 0:
        P = -P
    U, S, Vt = np.linalg.svd(P[:, :3])
    R = U @ Vt
    if np.linalg.det(R) < 0:
        return None
    return R, P[:, 3] / S.mean()


def refine(R, t, K, X, uv, iters=15):
    """Levenberg-Marquardt on the reprojection error, left-multiplicative rotation update."""
    lam, best = 1e-4, None
    for _ in range(iters + 1):
        xc = X @ R.T + t
        iz = 1.0 / xc[:, 2]
        x, y = xc[:, 0] * iz, xc[:, 1] * iz
        fx, fy = K[0, 0], K[1, 1]
        r = np.stack((fx * x + K[0, 2] - uv[:, 0], fy * y + K[1, 2] - uv[:, 1]), -1)
        cost = float((r ** 2).sum())
        if best is None or cost <= best[0]:
            Ju = np.stack((-fx * x * y, fx * (1 + x * x), -fx * y, fx * iz, 0 * x, -fx * x * iz), -1)
            Jv = np.stack((-fy * (1 + y * y), fy * x * y, fy * x, 0 * x, fy * iz, -fy * y * iz), -1)
            H = Ju.T @ Ju + Jv.T @ Jv
            g = Ju.T @ r[:, 0] + Jv.T @ r[:, 1]
            best = (cost, R, t, H, g)
            lam = max(lam * 0.1, 1e-9)
        else:
            lam = min(lam * 10, 1e6)
        _, R0, t0, H, g = best
        d = np.linalg.solve(H + lam * np.diag(np.diag(H)), -g)
        E = rodrigues(d[:3])
        R, t = E @ R0, E @ t0 + d[3:]
    return best[1], best[2], np.sqrt(best[0] / len(X))


def solve_pnp_ransac(pts3d, K, image_points=None, reproj_err=8.0, n_samples=300, seed=0, score_stride=1):
    """pts3d [H, W, 3] (pixel-grid correspondences, demo.py:166-168) or [n, 3] with image_points [n, 2].
    Returns (success, rvec, tvec, inlier_mask [n])."""
    pts3d = np.asarray(pts3d, np.float64)
    if image_points is None:
        H, W = pts3d.shape[:2]
        u, v = np.meshgrid(np.arange(W), np.arange(H))
        image_points = np.stack((u, v), -1).reshape(-1, 2)
    X = pts3d.reshape(-1, 3)
    uv = np.asarray(image_points, np.float64).reshape(-1, 2)
    rng = np.random.default_rng(seed)
    Xs, uvs = X[::score_stride], uv[::score_stride]
    best = (-1, None)
    for _ in range(n_samples):
        idx = rng.choice(len(X), 6, replace=False)
        pose = dlt_pose(X[idx], uv[idx], K)
        if pose is None:
            continue
        p, z = project(pose[0], pose[1], K, Xs)
        cnt = int(((((p - uvs) ** 2).sum(-1) < reproj_err ** 2) & (z > 0)).sum())
        if cnt > best[0]:
            best = (cnt, pose)
    if best[1] is None or best[0] < 6:
        return False, np.zeros(3), np.zeros(3), np.zeros(len(X), bool)
    R, t = best[1]
    p, z = project(R, t, K, X)
    mask = (((p - uv) ** 2).sum(-1) < reproj_err ** 2) & (z > 0)
    # DLT on 6 noisy points is a rough model: re-estimate on its inliers until the inlier set stops growing, as RANSAC
    # implementations do (local optimisation); the final refinement is the same LSQ problem cv2 solves
    for _ in range(6):
        R, t, _ = refine(R, t, K, X[mask], uv[mask], iters=5)
        p, z = project(R, t, K, X)
        new = (((p - uv) ** 2).sum(-1) < reproj_err ** 2) & (z > 0)
        grown = new.sum() > mask.sum()
        mask = new
        if not grown:
            break
    R, t, _ = refine(R, t, K, X[mask], uv[mask])
    return True, rodrigues_inv(R), t, mask
