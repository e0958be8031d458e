#!/usr/bin/env python
"""Warm, pipelined per-stage times of the engine at B=1, 512x384 (CUDA events around repeated stage calls;
unlike the ncu launch list the caches are warm and PDL overlap is on).  Diagnostic only: bench.py is the number."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spann3r_b200 import Spann3R, synth  # noqa: E402
from spann3r_b200.engine import MemoryBank  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
H, W, F_ = 384, 512, 10
sd = synth.make_state_dict(sharpen=True)
m = Spann3R(dus3r_name=None)
m.load_state_dict(sd, strict=True)
m = m.cuda().eval()
frames = [{"img": f["img"].cuda()} for f in synth.make_frames(F_, H, W, batch=B)]
m(frames)
eng = m._engine_for(B, H, W, n_frames=F_)
torch.cuda.synchronize()


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


imgs = torch.cat([f["img"] for f in frames], dim=0).contiguous()
out = {}
out["encode_10"] = timeit(lambda: eng.encode(imgs), 5)
feats = eng.encode(imgs).view(F_, B, eng.N, 1024)
f1, f2 = feats[0].contiguous(), feats[1].contiguous()
out["decode"] = timeit(lambda: eng.decode(f1, f2))
eng.decode(f1, f2)
out["keyheads"] = timeit(lambda: eng.keyheads(f1, f2))
k1, k2 = eng.keyheads(f1, f2)
out["heads"] = timeit(lambda: eng.heads())
pts, conf = eng.heads()
out["value"] = timeit(lambda: eng.value(pts[0], k1))
v = eng.value(pts[0], k1)
for nfr in (1, 4, 8):
    bank = MemoryBank(B, 4000 + 8 * eng.N, eng.device)
    for _ in range(nfr):
        eng.memory_append(bank, k1, v)
    out[f"memory_read_M{bank.len}"] = timeit(lambda: eng.memory_read(bank, k2, 5e-4))
bank = MemoryBank(B, 4000 + 8 * eng.N, eng.device)


def app():
    bank.len = 0
    eng.memory_append(bank, k1, v)


out["memory_append"] = timeit(app)
out["full_forward"] = timeit(lambda: m(frames), 3)
steps = F_ - 1
out["sum_of_stages"] = out["encode_10"] + steps * (out["decode"] + out["keyheads"] + out["heads"] + out["value"] +
                                                   out["memory_append"]) + (steps - 1) * out["memory_read_M3072"]
print(json.dumps({k: round(v, 4) for k, v in out.items()}))
