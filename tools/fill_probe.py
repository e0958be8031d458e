#!/usr/bin/env python
"""Is the GEMM main loop bound per SM (ingest port / latency x smem in flight) or by the shared L2 / NoC?
One 128 x 64 tile per CTA (bn 64, 48 KB per k-block), K long enough that the fixed cost is small, and 1..148 CTAs:
if the time per k-block does not change with the CTA count the limit is per SM."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spann3r_b200 import _lib as L  # noqa: E402


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


for bn in (64, 128, 256):
    for ncta in (1, 8, 37, 74, 148):
        res = []
        for K in (2048, 8192):
            rows, N = 128, bn * ncta
            x = torch.randn(rows, K, device="cuda")
            w = torch.randn(N, K, device="cuda") * K ** -0.5
            xp, wp = L.split(x), L.split(w)
            out = torch.empty(rows, N, device="cuda")
            d = L.GemmDesc()
            d.a_hi, d.a_lo, d.b_hi, d.b_lo = xp[0].data_ptr(), xp[1].data_ptr(), wp[0].data_ptr(), wp[1].data_ptr()
            d.groups, d.nb, d.h, d.w, d.kc, d.taps, d.n = 1, 1, 1, rows, K, 1, N
            d.force_bn = bn
            d.out_f32, d.ldo = out.data_ptr(), N
            res.append(timeit(lambda: L.gemm(d)))
        per_kb = (res[1] - res[0]) / ((8192 - 2048) / 64)
        stage_bytes = (128 + bn) * 64 * 4
        print(f"bn={bn:3d} ctas={ncta:3d}: K=2048 {res[0]:7.1f} us  K=8192 {res[1]:7.1f} us  -> {per_kb * 1e3:6.0f} ns / k-block "
              f"= {stage_bytes / (per_kb * 1e-6) / 1e9:6.1f} GB/s per SM (MMA floor {6 * bn / 1.9:5.0f} ns)", flush=True)


print("12 x 12 tiles (144 CTAs), operands L2-resident (every A tile shared by 12 CTAs, every B tile by 12):")
for bn in (64, 128, 256):
    res = []
    for K in (1024, 4096):
        rows, N = 128 * 12, bn * 12
        x = torch.randn(rows, K, device="cuda")
        w = torch.randn(N, K, device="cuda") * K ** -0.5
        xp, wp = L.split(x), L.split(w)
        out = torch.empty(rows, N, device="cuda")
        d = L.GemmDesc()
        d.a_hi, d.a_lo, d.b_hi, d.b_lo = xp[0].data_ptr(), xp[1].data_ptr(), wp[0].data_ptr(), wp[1].data_ptr()
        d.groups, d.nb, d.h, d.w, d.kc, d.taps, d.n = 1, 1, 1, rows, K, 1, N
        d.force_bn = bn
        d.out_f32, d.ldo = out.data_ptr(), N
        res.append(timeit(lambda: L.gemm(d)))
    per_kb = (res[1] - res[0]) / ((4096 - 1024) / 64)
    stage_bytes = (128 + bn) * 64 * 4
    print(f"bn={bn:3d}: K=1024 {res[0]:7.1f} us  K=4096 {res[1]:7.1f} us  -> {per_kb * 1e3:6.0f} ns / k-block "
          f"= {stage_bytes / (per_kb * 1e-6) / 1e9:6.1f} GB/s per SM, {144 * stage_bytes / (per_kb * 1e-6) / 1e12:5.2f} TB/s aggregate "
          f"(MMA floor {6 * bn / 1.9:5.0f} ns)", flush=True)
