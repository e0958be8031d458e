"""Post-path geometry on the GPU (first step): the reference's focal estimate.

`demo.py:148-150` copies `preds[0]['pts3d']` to the CPU and runs `estimate_focal_knowing_depth(pts3d, pp,
focal_mode='weiszfeld')` (dust3r/post_process.py:12-60).  Same function name, arguments and clipping here, computed on the
device by libspann3r_b200.so (csrc/geometry.cu); no CPU fallback.  The PnP-RANSAC that follows in demo.py (cv2, random
sampling) stays with the caller."""
from __future__ import annotations

import math

import torch

from . import _lib


def estimate_focal_knowing_depth(pts3d: torch.Tensor, pp, focal_mode: str = "weiszfeld", min_focal: float = 0.0,
                                 max_focal: float = float("inf")) -> torch.Tensor:
    """pts3d [B, H, W, 3] fp32 on the device, pp = (cx, cy) (tensor or pair) -> focal [B] on the device."""
    if focal_mode != "weiszfeld":
        raise NotImplementedError("only focal_mode='weiszfeld' (what demo.py uses) is built")
    _lib.require_device()
    if not (pts3d.is_cuda and pts3d.dtype == torch.float32 and pts3d.dim() == 4 and pts3d.shape[-1] == 3):
        raise ValueError("expected pts3d [B, H, W, 3] float32 on the GPU")
    pts3d = pts3d.contiguous()
    B, H, W, _ = pts3d.shape
    ppx, ppy = (float(v) for v in (pp.flatten().tolist() if torch.is_tensor(pp) else pp))
    base = max(H, W) / (2 * math.tan(math.radians(60) / 2))
    lo = min_focal * base
    hi = max_focal * base if math.isfinite(max_focal) else 3.0e38
    scratch = torch.empty(B * 148 * 2, dtype=torch.float32, device=pts3d.device)
    focal = torch.empty(B, dtype=torch.float32, device=pts3d.device)
    _lib.check(_lib.lib().s3r_focal_weiszfeld(_lib.ptr(pts3d), B, H, W, ppx, ppy, 10, lo, hi, _lib.ptr(scratch),
                                              _lib.ptr(focal), _lib.stream_ptr()), "s3r_focal_weiszfeld")
    return focal
