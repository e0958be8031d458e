#!/usr/bin/env python
"""Per-launch CUDA-event times of one 10-frame encoder call and one decode call (engine profile mode: launches are
bracketed by events, so no PDL overlap -- compare launches with each other, not with bench.py)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spann3r_b200 import Spann3R, synth  # noqa: E402

sd = synth.make_state_dict(sharpen=True)
m = Spann3R(dus3r_name=None)
m.load_state_dict(sd, strict=True)
m = m.cuda().eval()
frames = [{"img": f["img"].cuda()} for f in synth.make_frames(10, 384, 512)]
m(frames)
eng = m._engine_for(1, 384, 512, n_frames=10)
imgs = torch.cat([f["img"] for f in frames], dim=0).contiguous()
for what in ("encode", "decode", "value", "heads"):
    feats = eng.encode(imgs).view(10, 1, eng.N, 1024)
    f1, f2 = feats[0].contiguous(), feats[1].contiguous()
    eng.decode(f1, f2)
    k1, k2 = eng.keyheads(f1, f2)
    pts, conf = eng.heads()
    torch.cuda.synchronize()
    for rep in range(2):
        eng.profile(True)
        if what == "encode":
            eng.encode(imgs)
        elif what == "decode":
            eng.decode(f1, f2)
        elif what == "value":
            eng.value(pts[0], k1)
        else:
            eng.heads()
        if hasattr(eng, "profile_list"):
            lst = eng.profile_list()
        else:
            lst = None
        tot = eng.profile_read()
        eng.profile(False)
    print(what, {k: round(v, 3) if isinstance(v, float) else v for k, v in tot.items()})
    if lst:
        n = {"encode": 5, "decode": 9, "value": 5, "heads": 28}[what]
        start = 1 if what in ("encode", "decode", "value") else 0   # skip patch-embed / decoder_embed
        per = lst[start + n * 2: start + n * 3] if what != "heads" else lst   # third block / layer
        print("   ", " ".join(f"{'A' if k else 'G'}{ms * 1e3:.1f}" for ms, fl, k in per))
