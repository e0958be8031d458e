#!/usr/bin/env python
"""Throughput of the GPU input adapter (spann3r_b200/preprocess.py) on 1080p -> 512x384 frames, next to the CPU path it
replaces (PIL Lanczos + numpy normalise, what the reference's Demo dataset does per frame), and the achieved fraction of
the HBM roofline (algorithmic bytes = source crop read once + fp32 output written once).  One JSON line."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spann3r_b200 import preprocess as P  # noqa: E402

H, W, RES, N = 1080, 1920, (512, 384), 64
rng = np.random.default_rng(0)
frames = [rng.integers(0, 256, (H, W, 3), dtype=np.uint8) for _ in range(4)]
ad = P.FrameAdapter(RES)
dev = [torch.from_numpy(f).cuda() for f in frames]
pinned = [torch.from_numpy(f).pin_memory() for f in frames]
for f in dev:
    ad(f)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(N):
    ad(dev[i % 4])
e1.record()
torch.cuda.synchronize()
ms_dev = e0.elapsed_time(e1) / N
e0.record()
for i in range(N):
    ad(pinned[i % 4])
e1.record()
torch.cuda.synchronize()
ms_e2e = e0.elapsed_time(e1) / N
g = P.plan_frame(H, W, RES)
l, t, r, b = g["crop1"]
alg_bytes = (r - l) * (b - t) * 3 + 3 * RES[0] * RES[1] * 4
peaks = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
hbm = float(json.load(open(peaks)).get("hbm_gbs", 6578.3)) if os.path.exists(peaks) else 6650.0
# CPU path (Pillow + numpy), bounded sample
import PIL.Image  # noqa: E402
t0 = time.time()
n_cpu = 8
for i in range(n_cpu):
    im = PIL.Image.fromarray(frames[i % 4]).crop(g["crop1"]).resize(g["scaled"], resample=PIL.Image.Resampling.LANCZOS).crop(g["crop2"])
    x = (np.asarray(im).astype(np.float32) / 255.0 - 0.5) / 0.5
cpu_ms = (time.time() - t0) / n_cpu * 1e3
print(json.dumps({"metric": "input-adapter frames/s (1080p RGB -> 512x384 fp32 CHW, Pillow-exact Lanczos)",
                  "value": 1e3 / ms_dev, "unit": "frames/s", "ms_per_frame": ms_dev,
                  "e2e": {"value": 1e3 / ms_e2e, "unit": "frames/s", "h2d_bytes_per_frame": H * W * 3},
                  "roofline": {"bound": "hbm", "achieved": alg_bytes / (ms_dev * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                               "frac": alg_bytes / (ms_dev * 1e-3) / 1e9 / hbm, "algorithmic_bytes_per_frame": alg_bytes},
                  "cpu_baseline": {"value": 1e3 / cpu_ms, "unit": "frames/s", "cores": 1, "kind": "port",
                                   "sample": f"{n_cpu} frames, PIL crop + LANCZOS resize + crop + numpy normalise"}}))
