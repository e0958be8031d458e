#!/usr/bin/env python
"""Fixed vs per-block cost of the attention and GEMM kernels (CUDA events, back-to-back launches with PDL)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spann3r_b200 import _lib as L  # noqa: E402


def timeit(fn, iters=40):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


print("attention: BH, nq, nk -> us")
for BH, nq in ((24, 768), (160, 768), (16, 768)):
    for nk in (128, 256, 384, 768):
        q = torch.randn(BH, nq, 64, device="cuda")
        k = torch.randn(BH, nk, 64, device="cuda")
        vt = torch.randn(BH, 64, nk, device="cuda")
        heads = 8
        oh = torch.empty(BH // heads * nq, heads * 64, dtype=torch.bfloat16, device="cuda")
        ol = torch.empty_like(oh)
        fn = lambda: L.check(L.lib().s3r_attention(L.ptr(q), L.ptr(k), L.ptr(vt), BH, heads, nq, nk, nk, L.ptr(oh), L.ptr(ol),  # noqa
                                                   None, heads * 64, L.stream_ptr()), "attn")
        print(f"  BH={BH:4d} nq={nq} nk={nk:4d}: {timeit(fn):7.1f} us", flush=True)

print("gemm G=2 M=768 N=768 (bn64, 144 CTAs): K -> us")
for K in (64, 128, 256, 512, 768, 1536, 3072):
    G, rows, N = 2, 768, 768
    x = torch.randn(G * rows, K, device="cuda")
    w = torch.randn(G * N, K, device="cuda")
    xp, wp = L.split(x), L.split(w)
    out = torch.empty(G * rows, N, device="cuda")
    r = torch.randn(G * rows, N, device="cuda")
    b = torch.randn(G * N, device="cuda")
    d = L.GemmDesc()
    d.a_hi, d.a_lo, d.b_hi, d.b_lo = xp[0].data_ptr(), xp[1].data_ptr(), wp[0].data_ptr(), wp[1].data_ptr()
    d.groups, d.nb, d.h, d.w, d.kc, d.taps, d.n = G, 1, 1, rows, K, 1, N
    d.force_bn = 64
    d.bias = b.data_ptr()
    d.res1, d.ldr1 = r.data_ptr(), N
    d.out_f32, d.ldo = out.data_ptr(), N
    t_full = timeit(lambda: L.gemm(d))
    d.bias = None
    d.res1 = None
    t_nores = timeit(lambda: L.gemm(d))
    print(f"  K={K:5d}: {t_full:6.1f} us (bias+residual epilogue)   {t_nores:6.1f} us (plain store)", flush=True)

print("layernorm 1536x768 planes:", timeit(lambda: L.layernorm(torch.empty(1536, 768, device='cuda'), torch.ones(768, device='cuda'), torch.zeros(768, device='cuda'), 1e-6, False, True)))
x = torch.randn(1536, 768, device="cuda"); w1 = torch.ones(768, device="cuda"); b1 = torch.zeros(768, device="cuda")
hi = torch.empty(1536, 768, dtype=torch.bfloat16, device="cuda"); lo = torch.empty_like(hi)
fn = lambda: L.check(L.lib().s3r_layernorm(L.ptr(x), 768, L.ptr(w1), L.ptr(b1), 0, 0, 1e-6, 1536, 768, None, 768, L.ptr(hi), L.ptr(lo), 768, 0, 0, L.stream_ptr()), "ln")  # noqa
print("layernorm kernel only 1536x768 -> planes: %.1f us" % timeit(fn))
