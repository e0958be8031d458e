"""Camera pose from a pointmap (second step): `cv2.solvePnPRansac` of demo.py:166-180.

Three implementations are held to the golden poses the REAL cv2 call produced on seeded synthetic pointmaps
(tools/make_golden_pnp.py -> tests/golden/pnp.json):
  * the oracle (oracle/pnp_oracle.py, numpy, 6-point DLT RANSAC + LM)                                    -- CPU
  * the product's device math header (spann3r_b200/csrc/pnp_math.cuh) compiled for the host by g++ and driven
    sequentially (tests/native/pnp_host_check.cpp): pins P3P / Gauss-Newton / exp-log arithmetic without a GPU  -- CPU
  * the CUDA path through the C ABI (`s3r_pnp_ransac`), which must also reproduce the host run of the same header for
    the same seed                                                                                              -- GPU
Tolerances: 1e-6 absolute on the outlier-free case (same least-squares optimum), 3e-4 otherwise (two RANSAC runs differ
by a handful of inliers that straddle the 8 px threshold); the translation scale of the cases is ~1.
"""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from spann3r_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(GOLDEN, "pnp.json")))


def _tol(case):
    return 1e-6 if case[6] == 0.0 else 3e-4


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("pnp") / "pnp_host_check.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-std=c++17", "-x", "c++",
                           os.path.join(HERE, "native", "pnp_host_check.cpp"), "-o", so])
    L = C.CDLL(so)
    L.pnp_host_check.restype = C.c_int
    L.pnp_host_check.argtypes = ([C.c_void_p, C.c_void_p, C.c_longlong, C.c_int] + [C.c_double] * 5 +
                                 [C.c_int, C.c_int, C.c_ulonglong, C.c_void_p, C.c_void_p])
    L.p3p_host.restype = C.c_int
    L.p3p_host.argtypes = [C.c_void_p] * 3
    return L


def _host_run(L, pts, K, n_samples=100, iters=15, seed=0):
    pts = np.ascontiguousarray(pts.reshape(-1, 3), np.float32)
    out = np.zeros(18)
    mask = np.zeros(len(pts), np.uint8)
    ok = L.pnp_host_check(pts.ctypes.data, None, len(pts), K_w(pts, K), K[0, 0], K[1, 1], K[0, 2], K[1, 2], 8.0, n_samples,
                          iters, seed, out.ctypes.data, mask.ctypes.data)
    return ok, out, mask


def K_w(pts, K):   # image width of the synthetic cases: the principal point sits at the centre
    return int(round(2 * K[0, 2]))


def _rot(rv):
    th = np.linalg.norm(rv)
    if th < 1e-12:
        return np.eye(3)
    k = rv / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx


def test_p3p_recovers_the_pose(host_lib):
    """Grunert P3P of pnp_math.cuh: the true pose is among the (<= 4) solutions for random non-degenerate triangles."""
    rng = np.random.default_rng(0)
    done = 0
    for _ in range(600):
        R = _rot(rng.normal(0, 0.5, 3))
        t = rng.normal(0, 0.5, 3) + np.array([0, 0, 4.0])
        P = rng.normal(0, 1, (3, 3))
        Y = P @ R.T + t
        if (Y[:, 2] < 0.1).any():
            continue
        f = np.ascontiguousarray(Y / np.linalg.norm(Y, axis=1, keepdims=True))
        out = np.zeros(48)
        n = host_lib.p3p_host(np.ascontiguousarray(P).ctypes.data, f.ctypes.data, out.ctypes.data)
        assert 1 <= n <= 4
        best = min(np.abs(out[12 * k: 12 * k + 9].reshape(3, 3) - R).max() + np.abs(out[12 * k + 9: 12 * k + 12] - t).max()
                   for k in range(n))
        assert best < 1e-6, best
        done += 1
    assert done > 500


@pytest.mark.parametrize("i", range(len(synth.PNP_CASES)))
def test_device_math_on_host_matches_cv2_golden(host_lib, i):
    case, g = synth.PNP_CASES[i], GOLD["cases"][i]
    pts, K = synth.make_pointmap_case(*case)
    ok, out, mask = _host_run(host_lib, pts, K)
    assert ok == 1 and g["success"]
    assert np.abs(out[12:15] - g["rvec"]).max() < _tol(case), (out[12:15], g["rvec"])
    assert np.abs(out[9:12] - g["tvec"]).max() < _tol(case), (out[9:12], g["tvec"])
    assert abs(int(mask.sum()) - g["n_inliers"]) <= max(3, 0.002 * g["n_inliers"])
    assert np.abs(_rot(out[12:15]) - out[:9].reshape(3, 3)).max() < 1e-12      # rvec == log(R)


@pytest.mark.parametrize("i", range(len(synth.PNP_CASES)))
def test_oracle_matches_cv2_golden(i):
    from oracle import pnp_oracle as po
    case, g = synth.PNP_CASES[i], GOLD["cases"][i]
    pts, K = synth.make_pointmap_case(*case)
    ok, rvec, tvec, mask = po.solve_pnp_ransac(pts, K, seed=0, score_stride=8)
    assert ok
    assert np.abs(rvec - g["rvec"]).max() < _tol(case) and np.abs(tvec - g["tvec"]).max() < _tol(case)
    assert abs(int(mask.sum()) - g["n_inliers"]) <= max(3, 0.002 * g["n_inliers"])


def test_explicit_image_points_and_degenerate_input(host_lib):
    """Sparse correspondences (img_pts given) and an unsolvable input (all points identical -> no hypothesis)."""
    pts, K = synth.make_pointmap_case(*synth.PNP_CASES[0])
    H, W = pts.shape[:2]
    u, v = np.meshgrid(np.arange(W), np.arange(H))
    sel = np.random.default_rng(1).choice(H * W, 5000, replace=False)
    p = np.ascontiguousarray(pts.reshape(-1, 3)[sel], np.float32)
    im = np.ascontiguousarray(np.stack((u, v), -1).reshape(-1, 2)[sel], np.float32)
    out = np.zeros(18)
    mask = np.zeros(len(p), np.uint8)
    ok = host_lib.pnp_host_check(p.ctypes.data, im.ctypes.data, len(p), 0, K[0, 0], K[1, 1], K[0, 2], K[1, 2], 8.0, 100, 15, 3,
                                 out.ctypes.data, mask.ctypes.data)
    assert ok == 1 and np.abs(out[12:15] - GOLD["cases"][0]["rvec"]).max() < 2e-4
    p[:] = 1.0
    ok = host_lib.pnp_host_check(p.ctypes.data, im.ctypes.data, len(p), 0, K[0, 0], K[1, 1], K[0, 2], K[1, 2], 8.0, 100, 15, 3,
                                 out.ctypes.data, mask.ctypes.data)
    assert ok == 0 and out[17] == 0.0


# ------------------------------------------------------------------------------------------------------------------
# GPU: the CUDA path through the C ABI
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_cuda_pnp_matches_cv2_golden_and_host_math(host_lib):
    from spann3r_b200.postprocess import solve_pnp_ransac
    for i, (case, g) in enumerate(zip(synth.PNP_CASES, GOLD["cases"])):
        pts, K = synth.make_pointmap_case(*case)
        ok, rvec, tvec, inl = solve_pnp_ransac(torch.from_numpy(pts)[None].cuda(), K, seed=7)
        torch.cuda.synchronize()
        rvec, tvec = rvec[0].cpu().numpy(), tvec[0].cpu().numpy()
        assert bool(ok[0]) and inl.shape == (1,) + pts.shape[:2] and inl.dtype == torch.bool
        assert np.abs(rvec - g["rvec"]).max() < _tol(case), (i, rvec, g["rvec"])
        assert np.abs(tvec - g["tvec"]).max() < _tol(case), (i, tvec, g["tvec"])
        # same header, same seed, same sample count on the host: same hypotheses, hence the same inlier set and optimum
        # up to libm / FMA-contraction differences between g++ and nvcc (a threshold-straddling point may flip)
        hok, hout, hmask = _host_run(host_lib, pts, K, seed=7)
        assert hok == 1
        flips = int((inl[0].cpu().numpy().reshape(-1) != hmask.astype(bool)).sum())
        dpose = max(np.abs(rvec - hout[12:15]).max(), np.abs(tvec - hout[9:12]).max())
        print(f"case {i}: inliers {int(inl.sum())} (cv2 {g['n_inliers']}), mask flips vs host {flips}, pose diff vs host {dpose:.1e}")
        assert flips <= 3 and dpose < 1e-6, (i, flips, dpose)


@pytest.mark.gpu
def test_cuda_pnp_batched_sparse_and_degenerate():
    """A batch of frames in one call == each frame alone; explicit image points; an unsolvable frame reports failure
    without disturbing its neighbours."""
    from spann3r_b200.postprocess import solve_pnp_ransac
    cases = [synth.PNP_CASES[0], synth.PNP_CASES[1], synth.PNP_CASES[3]]
    maps = [synth.make_pointmap_case(*c) for c in cases]
    K = maps[0][1]                               # shared intrinsics: re-synthesise every frame with the same focal
    maps = [synth.make_pointmap_case(c[0], c[1], cases[0][2], *c[3:]) for c in cases]
    batch = torch.stack([torch.from_numpy(m[0]) for m in maps]).cuda()
    bad = batch.clone()
    bad[1] = 1.0
    ok, rvec, tvec, inl = solve_pnp_ransac(batch, K, seed=11)
    okb, rvecb, tvecb, inlb = solve_pnp_ransac(bad, K, seed=11)
    for j in range(3):
        o1, r1, t1, m1 = solve_pnp_ransac(batch[j: j + 1], K, seed=11)
        assert bool(o1[0]) and bool(ok[j])
        assert torch.equal(r1[0], rvec[j]) and torch.equal(t1[0], tvec[j]) and torch.equal(m1[0], inl[j])
        assert np.abs(rvec[j].cpu().numpy() - np.array(cases[j][3])).max() < 2e-3
    assert okb.tolist() == [True, False, True] and int(inlb[1].sum()) == 0
    assert torch.equal(rvecb[0], rvec[0]) and torch.equal(tvecb[2], tvec[2])
    # sparse correspondences
    H, W = batch.shape[1:3]
    u, v = torch.meshgrid(torch.arange(W), torch.arange(H), indexing="xy")
    grid = torch.stack((u, v), -1).reshape(-1, 2).float()
    sel = torch.randperm(H * W, generator=torch.Generator().manual_seed(0))[:8000]
    oks, rs, ts, ms = solve_pnp_ransac(batch[:1].reshape(1, -1, 3)[:, sel].contiguous(), K,
                                       image_points=grid[sel][None].cuda().contiguous(), seed=5)
    assert bool(oks[0]) and ms.shape == (1, 8000)
    assert np.abs(rs[0].cpu().numpy() - np.array(cases[0][3])).max() < 1e-3
