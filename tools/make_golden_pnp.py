#!/usr/bin/env python
"""Golden poses for the GPU PnP (second step): runs the call demo.py:166-180 makes --
`cv2.solvePnPRansac(pts.reshape(-1, 3), pixel grid, intrinsic, zeros(4))`, OpenCV 4.13.0 here (a third-party dependency the
reference leaves unpinned) -- on the seeded synthetic pointmaps of `spann3r_b200.synth.PNP_CASES` and commits
rvec / tvec / inlier count to tests/golden/pnp.json.  Authoring-container tool: nothing under tests/ or bench.py imports it."""
import json
import os
import sys
import time

import cv2
import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from spann3r_b200 import synth  # noqa: E402

out = {"cv2": cv2.__version__, "cases": []}
for case in synth.PNP_CASES:
    H, W = case[0], case[1]
    pts, K = synth.make_pointmap_case(*case)
    u, v = np.meshgrid(np.arange(W), np.arange(H))
    points_2d = np.stack((u, v), axis=-1)
    cv2.setRNGSeed(0)
    t0 = time.time()
    ok, rvec, tvec, inl = cv2.solvePnPRansac(pts.reshape(-1, 3).astype(np.float32), points_2d.reshape(-1, 2).astype(np.float32),
                                             K.astype(np.float32), np.zeros(4).astype(np.float32))
    dt = time.time() - t0
    out["cases"].append({"args": list(case[:3]) + [list(case[3]), list(case[4])] + list(case[5:]), "success": bool(ok),
                         "rvec": rvec.ravel().tolist(), "tvec": tvec.ravel().tolist(), "n_inliers": int(len(inl)),
                         "cv2_seconds": round(dt, 3)})
    print(case, ok, rvec.ravel(), tvec.ravel(), len(inl), "%.2fs" % dt)
with open(os.path.join(REPO, "tests", "golden", "pnp.json"), "w") as f:
    json.dump(out, f, indent=1)
