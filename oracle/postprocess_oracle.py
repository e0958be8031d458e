"""ORACLE (test infrastructure, not product code): torch restatement of the reference's focal estimate
`estimate_focal_knowing_depth(pts3d, pp, focal_mode='weiszfeld')` (dust3r/post_process.py:12-60; demo.py:148-150).
PINNED: tests/test_postprocess.py holds it to the values the REAL reference function returns on seeded synthetic
pointmaps (tools/make_golden_focal.py -> tests/golden/focal.json)."""
import math

import torch


def focal_weiszfeld(pts3d: torch.Tensor, pp, min_focal: float = 0.0, max_focal: float = float("inf"), iters: int = 10):
    B, H, W, _ = pts3d.shape
    jj, ii = torch.meshgrid(torch.arange(H, device=pts3d.device), torch.arange(W, device=pts3d.device), indexing="ij")
    pix = torch.stack((ii, jj), dim=-1).view(1, H * W, 2) - torch.as_tensor(pp, dtype=torch.float32, device=pts3d.device).view(-1, 1, 2)
    p = pts3d.reshape(B, H * W, 3)
    a = (p[..., :2] / p[..., 2:3]).nan_to_num(posinf=0, neginf=0)            # :37
    d_px = (a * pix).sum(-1)
    d_xx = a.square().sum(-1)
    f = d_px.mean(1) / d_xx.mean(1)                                          # :42 closed-form l2 start
    for _ in range(iters):                                                   # :45-51 re-weighted least squares
        w = (pix - f.view(-1, 1, 1) * a).norm(dim=-1).clip(min=1e-8).reciprocal()
        f = (w * d_px).mean(1) / (w * d_xx).mean(1)
    base = max(H, W) / (2 * math.tan(math.radians(60) / 2))                  # :55
    return f.clip(min=min_focal * base, max=max_focal * base)


def focal_median(pts3d: torch.Tensor, pp, min_focal: float = 0.0, max_focal: float = float("inf")):
    """focal_mode='median' (dust3r/post_process.py:26-36): nanmedian over the fx votes u z / x and fy votes v z / y."""
    B, H, W, _ = pts3d.shape
    jj, ii = torch.meshgrid(torch.arange(H, device=pts3d.device), torch.arange(W, device=pts3d.device), indexing="ij")
    pix = torch.stack((ii, jj), dim=-1).view(1, H * W, 2) - torch.as_tensor(pp, dtype=torch.float32, device=pts3d.device).view(-1, 1, 2)
    p = pts3d.reshape(B, H * W, 3)
    u, v = pix.unbind(dim=-1)
    x, y, z = p.unbind(dim=-1)
    votes = torch.cat((((u * z) / x).view(B, -1), ((v * z) / y).view(B, -1)), dim=-1)      # :30-34
    f = torch.nanmedian(votes, dim=-1).values                                              # :35
    base = max(H, W) / (2 * math.tan(math.radians(60) / 2))
    return f.clip(min=min_focal * base, max=max_focal * base)
