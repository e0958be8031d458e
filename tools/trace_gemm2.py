#!/usr/bin/env python
"""Timeline of CTA 0 of the CTA-pair GEMM kernel (gemm2.cu) on the decoder's shapes at B = 1 (2 groups x 768 rows), in a
steady-state run of back-to-back launches: where do the 17 / 35 / 43 us of proj / fc1 / qkv (tools/launch_times.py) go?
Stamps (%globaltimer, ns): entry, prologue done, dependency wait done, exit; per tile: first operands landed, all MMAs issued,
accumulator seen by the epilogue, first chunk done, epilogue done, producer issued the last k-block."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spann3r_b200 import _lib as L  # noqa: E402

G, rows = 2, 768
CASES = [("proj  K=768  N=768  bn 64 ", 768, 768, 2064, L.ACT_NONE), ("fc2   K=3072 N=768  bn 64 ", 3072, 768, 2064, L.ACT_NONE),
         ("fc1   K=768  N=3072 bn 128 gelu", 768, 3072, 2128, L.ACT_GELU), ("qkv~  K=768  N=3840 bn 128 plain", 768, 3840, 2128, L.ACT_NONE)]
for name, K, N, fbn, act in CASES:
    n_launch = 16
    w = torch.randn(G * N, K, device="cuda") * K ** -0.5
    wp = L.split(w)
    a = [L.split(torch.randn(G * rows, K, device="cuda")) for _ in range(2)]
    outs = [(torch.empty(G * rows, N, dtype=torch.bfloat16, device="cuda"), torch.empty(G * rows, N, dtype=torch.bfloat16, device="cuda"))
            for _ in range(2)]
    trace = torch.zeros(n_launch, 32, dtype=torch.int64, device="cuda")
    b = torch.randn(G * N, device="cuda")
    descs = []
    for i in range(n_launch):
        d = L.GemmDesc()
        d.a_hi, d.a_lo, d.b_hi, d.b_lo = a[i % 2][0].data_ptr(), a[i % 2][1].data_ptr(), wp[0].data_ptr(), wp[1].data_ptr()
        d.groups, d.nb, d.h, d.w, d.kc, d.taps, d.n = G, 1, 1, rows, K, 1, N
        d.force_bn, d.act = fbn, act
        d.bias = b.data_ptr()
        d.out_hi, d.out_lo, d.ldp = outs[i % 2][0].data_ptr(), outs[i % 2][1].data_ptr(), N
        d.trace = trace[i].data_ptr()
        descs.append(d)
    for rep in range(3):
        for d in descs:
            L.gemm(d)
    torch.cuda.synchronize()
    t = trace.cpu().double()
    print(f"{name}: CTA 0, ns relative to this launch's dependency-wait release (launches 8..11); tiles of CTA 0: "
          f"{int((t[8, 5::6] > 0).sum())}")
    for i in range(8, 12):
        base = t[i, 2]
        head = f"   prev-exit->wait {t[i, 2] - t[i - 1, 3]:6.0f}  entry {t[i, 0] - base:7.0f}  exit {t[i, 3] - base:7.0f} |"
        tiles = []
        for it in range(4):
            s = t[i, 4 + 6 * it: 10 + 6 * it]
            if s[1] == 0:
                break
            tiles.append(f" T{it}: ops {s[0] - base:6.0f} mma_done {s[1] - base:6.0f} epi_start {s[2] - base:6.0f} chunk0 {s[3] - base:6.0f} "
                         f"epi_end {s[4] - base:6.0f} loads_out {s[5] - base:6.0f}")
        print(head + "".join(tiles))
    per = (t[8:14, 3] - t[7:13, 3]).mean()
    print(f"   mean exit-to-exit period {per:.0f} ns")
