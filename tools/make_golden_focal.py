#!/usr/bin/env python
"""Golden values of the reference's focal estimate: the REAL dust3r.post_process.estimate_focal_knowing_depth (CPU) on
seeded synthetic pointmaps -> tests/golden/focal.json.  The pointmaps are regenerated from the seeds by the tests."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, "/root/reference")
from dust3r.post_process import estimate_focal_knowing_depth  # noqa: E402


def pointmap(seed, B, H, W, f_true):
    """A noisy pinhole pointmap with focal f_true, some invalid depths (z = 0 / negative) and outliers."""
    g = torch.Generator().manual_seed(seed)
    jj, ii = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    z = 1.0 + 2.0 * torch.rand(B, H, W, generator=g)
    x = (ii - W / 2) * z / f_true + 0.01 * torch.randn(B, H, W, generator=g)
    y = (jj - H / 2) * z / f_true + 0.01 * torch.randn(B, H, W, generator=g)
    pts = torch.stack((x, y, z), dim=-1)
    bad = torch.rand(B, H, W, generator=g) < 0.01
    pts[bad] = 5.0 * torch.randn(int(bad.sum()), 3, generator=g)
    pts[:, 0, 0, 2] = 0.0
    return pts.float()


if __name__ == "__main__":
    cases = []
    for seed, B, H, W, f_true in [(1, 1, 384, 512, 300.0), (2, 2, 224, 224, 180.0), (3, 1, 336, 512, 700.0), (4, 1, 64, 80, 55.0)]:
        pts = pointmap(seed, B, H, W, f_true)
        pp = torch.tensor((W / 2, H / 2))
        f = estimate_focal_knowing_depth(pts, pp, focal_mode="weiszfeld")
        fm = estimate_focal_knowing_depth(pts, pp, focal_mode="median")      # an element of the vote set: exact in fp32
        cases.append(dict(seed=seed, B=B, H=H, W=W, f_true=f_true, focal=[float(v) for v in f],
                          focal_median=[float(v) for v in fm]))
        print(cases[-1])
    json.dump({"cases": cases}, open(os.path.join(ROOT, "tests", "golden", "focal.json"), "w"), indent=1)
