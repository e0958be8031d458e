"""ORACLE (test infrastructure, not product code): CPU restatement of the reference's input adapter -- the step right
before the hot path.

What the reference does per frame (spann3r/datasets/demo.py:57-86 -> dust3r/datasets/base/base_stereo_view_dataset.py:143-194
`_crop_resize_if_necessary` -> dust3r/datasets/utils/cropping.py:55-124 -> dust3r/utils/image.py:23 `ImgNorm`):

  1. centre crop on the principal point (pseudo intrinsics: cx = W // 2, cy = H // 2)          base_stereo...:155-168
  2. transpose the target resolution for portrait images                                       :170-178
  3. Lanczos down-scale with PIL so that the image CONTAINS the target (floor(size * scale))   cropping.py:55-84
  4. centred crop to the target resolution                                                     base_stereo...:190-192
  5. ToTensor (uint8 -> float32 / 255, HWC -> CHW) and Normalize(0.5, 0.5)                      image.py:23

Step 3 lives in a third-party dependency that is NOT vendored in the reference: Pillow (`requirements.txt` leaves it
unpinned; this image has Pillow 12.2.0).  Its published algorithm (src/libImaging/Resample.c: `precompute_coeffs`,
`normalize_coeffs_8bpc`, `ImagingResampleHorizontal_8bpc`, `ImagingResampleVertical_8bpc`) is restated below in numpy:
separable, horizontal pass first, uint8 intermediate, coefficients in 22-bit fixed point, round-half-up, clip to
[0, 255].  PINNED: tests/test_input_adapter.py holds this restatement bit-exact against Pillow itself (and the whole
pipeline against the reference's own functions imported from /root/reference where that is present) on random images.
"""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2      # Resample.c
LANCZOS_SUPPORT = 3.0


def _sinc(x: float) -> float:
    if x == 0.0:
        return 1.0
    x = x * math.pi
    return math.sin(x) / x


def lanczos_filter(x: float) -> float:
    """Resample.c: lanczos_filter -- truncated sinc, -3 <= x < 3."""
    if -3.0 <= x < 3.0:
        return _sinc(x) * _sinc(x / 3.0)
    return 0.0


def precompute_coeffs(in_size: int, in0: float, in1: float, out_size: int):
    """Resample.c: precompute_coeffs + normalize_coeffs_8bpc -> (bounds [out, 2] int32 (first, count),
    coeffs [out, ksize] int32 fixed point, ksize)."""
    scale = filterscale = (in1 - in0) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = LANCZOS_SUPPORT * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.float64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = in0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        ww = 0.0
        for x in range(xmax):
            w = lanczos_filter((x + xmin - center + 0.5) * ss)
            kk[xx, x] = w
            ww += w
        if ww != 0.0:
            kk[xx, :xmax] /= ww
        bounds[xx] = (xmin, xmax)
    fixed = np.where(kk < 0, np.trunc(-0.5 + kk * (1 << PRECISION_BITS)), np.trunc(0.5 + kk * (1 << PRECISION_BITS)))
    return bounds, fixed.astype(np.int32), ksize


def _clip8(ss):
    return np.clip(ss >> PRECISION_BITS, 0, 255).astype(np.uint8)


def resample_horizontal(img: np.ndarray, out_w: int, bounds, kk) -> np.ndarray:
    """ImagingResampleHorizontal_8bpc on an [H, W, C] uint8 array."""
    h, _, c = img.shape
    out = np.empty((h, out_w, c), dtype=np.uint8)
    src = img.astype(np.int64)
    for xx in range(out_w):
        x0, n = bounds[xx]
        acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(src[:, x0:x0 + n, :], kk[xx, :n].astype(np.int64), axes=([1], [0]))
        out[:, xx, :] = _clip8(acc)
    return out


def resample_vertical(img: np.ndarray, out_h: int, bounds, kk) -> np.ndarray:
    """ImagingResampleVertical_8bpc on an [H, W, C] uint8 array."""
    _, w, c = img.shape
    out = np.empty((out_h, w, c), dtype=np.uint8)
    src = img.astype(np.int64)
    for yy in range(out_h):
        y0, n = bounds[yy]
        acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(kk[yy, :n].astype(np.int64), src[y0:y0 + n], axes=([0], [0]))
        out[yy] = _clip8(acc)
    return out


def pil_resize_lanczos(img: np.ndarray, out_w: int, out_h: int) -> np.ndarray:
    """PIL.Image.resize((out_w, out_h), LANCZOS) for an RGB uint8 array (Resample.c: ImagingResample -- horizontal pass
    over the rows the vertical pass needs, then vertical pass)."""
    h, w, _ = img.shape
    bh, kh, _ = precompute_coeffs(w, 0.0, float(w), out_w)
    bv, kv, _ = precompute_coeffs(h, 0.0, float(h), out_h)
    if out_w != w:
        ybox_first = int(bv[0, 0])
        ybox_last = int(bv[out_h - 1, 0] + bv[out_h - 1, 1])
        tmp = resample_horizontal(img[ybox_first:ybox_last], out_w, bh, kh)
        bv = bv.copy()
        bv[:, 0] -= ybox_first
    else:
        tmp = img
    if out_h != h:
        return resample_vertical(tmp, out_h, bv, kv)
    return tmp


def plan_frame(h: int, w: int, resolution=(512, 384), square_flip: bool = False):
    """Geometry of steps 1-4 for an [h, w] image with the demo's pseudo intrinsics, as integers:
    returns dict(crop1=(l, t, r, b), scaled=(W2, H2), crop2=(l, t, r, b), out=(W_out, H_out)).
    `square_flip`: the reference draws rng.integers(2) for (nearly) square images; pass that draw."""
    cx, cy = w // 2, h // 2                                   # demo.py:73-74 (then .round().astype(int): integers already)
    mx, my = min(cx, w - cx), min(cy, h - cy)
    assert mx > w / 5 and my > h / 5
    l, t, r, b = cx - mx, cy - my, cx + mx, cy + my
    W, H = r - l, b - t
    res = tuple(resolution)
    assert res[0] >= res[1]
    if H > 1.1 * W:
        res = res[::-1]
    elif 0.9 < H / W < 1.1 and res[0] != res[1] and square_flip:
        res = res[::-1]
    # rescale_image_depthmap (cropping.py:55-84)
    scale_final = max(res[0] / W, res[1] / H) + 1e-8
    W2, H2 = int(np.floor(W * scale_final)), int(np.floor(H * scale_final))
    # intrinsics after crop 1 / rescale (colmap convention +0.5), then camera_matrix_of_crop + bbox_from_intrinsics_in_out
    cx1, cy1 = float(np.float32(cx - l)), float(np.float32(cy - t))
    cx2 = np.float32((np.float32(cx1) + np.float32(0.5)) * np.float32(scale_final) - np.float32(0.5))
    cy2 = np.float32((np.float32(cy1) + np.float32(0.5)) * np.float32(scale_final) - np.float32(0.5))
    margins = np.asarray((W2, H2)) * 1.0 - np.asarray(res)
    assert np.all(margins >= 0.0)
    off = 0.5 * margins
    cx3 = np.float32(np.float32(cx2 + np.float32(0.5)) - off[0]) - np.float32(0.5)
    cy3 = np.float32(np.float32(cy2 + np.float32(0.5)) - off[1]) - np.float32(0.5)
    l2 = int(np.int32(np.round(cx2 - cx3)))
    t2 = int(np.int32(np.round(cy2 - cy3)))
    return dict(crop1=(l, t, r, b), scaled=(W2, H2), crop2=(l2, t2, l2 + res[0], t2 + res[1]), out=res)


def preprocess_frame(rgb: np.ndarray, resolution=(512, 384), square_flip: bool = False) -> np.ndarray:
    """uint8 [H, W, 3] -> float32 [3, H_out, W_out] exactly as the reference's Demo dataset + ImgNorm produce it."""
    p = plan_frame(rgb.shape[0], rgb.shape[1], resolution, square_flip)
    l, t, r, b = p["crop1"]
    img = rgb[t:b, l:r]
    W2, H2 = p["scaled"]
    img = pil_resize_lanczos(np.ascontiguousarray(img), W2, H2)
    l2, t2, r2, b2 = p["crop2"]
    img = img[t2:b2, l2:r2]
    x = img.astype(np.float32) / np.float32(255.0)
    x = (x - np.float32(0.5)) / np.float32(0.5)
    return np.ascontiguousarray(x.transpose(2, 0, 1))
