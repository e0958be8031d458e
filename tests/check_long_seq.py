#!/usr/bin/env python
"""BASELINE config[3]: 100-frame 512x384 sequence (bank saw-tooth 4000..7840, prunes) -- CUDA path vs the oracle port
run on the same GPU in strict fp32, plus timing of both."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import spann3r_oracle as orc  # noqa: E402
from spann3r_b200 import Spann3R, synth  # noqa: E402

F_ = int(sys.argv[1]) if len(sys.argv) > 1 else 100
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
sdc = synth.make_state_dict(sharpen=True)
m = Spann3R(dus3r_name=None)
m.load_state_dict(sdc, strict=True)
m = m.cuda().eval()
frames = [{"img": f["img"].cuda()} for f in synth.make_frames(F_, 384, 512)]
import io, contextlib  # noqa
buf = io.StringIO()
with contextlib.redirect_stdout(buf):
    m(frames[:12])
    torch.cuda.synchronize()
    t0 = time.time()
    preds, _, mem = m(frames, return_memory=True)
    torch.cuda.synchronize()
    t1 = time.time()
print("cuda path: %d frames in %.3fs = %.1f frames/s; prunes: %d; bank len %d wm %d lm %d" %
      (F_, t1 - t0, F_ / (t1 - t0), buf.getvalue().count("Memory pruned"), mem.bank.len, mem.wm, mem.lm))
keep = [{k: v.clone() for k, v in p.items()} for p in preds]
sd = {k: v.cuda() for k, v in sdc.items()}
t0 = time.time()
ref, _, omem = orc.forward(sd, frames, return_memory=True)
torch.cuda.synchronize()
print("oracle eager fp32 on GPU: %.1f frames/s; bank len %d wm %d lm %d" % (F_ / (time.time() - t0), omem.mem_k.shape[1], omem.wm, omem.lm))
errs = []
for p, r in zip(keep, ref):
    k = "pts3d" if "pts3d" in r else "pts3d_in_other_view"
    e = float((p[k].double() - r[k].double()).norm() / r[k].double().norm())
    errs.append(e)
finite = all(torch.isfinite(r[k]).all().item() for r in ref for k in r)
print("reference finite:", finite)
print("rel-L2 per frame: max %.2e  median %.2e  first-10 %s  last-5 %s" %
      (max(errs), sorted(errs)[len(errs) // 2], ["%.1e" % e for e in errs[:10]], ["%.1e" % e for e in errs[-5:]]))
