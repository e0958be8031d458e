#!/usr/bin/env python
"""A/B of the encoder-overlap experiment (Spann3R.overlap_encoder, S3R_ENC_OVERLAP=1): same 10-frame
512x384 sequences with the encoder batched up front (default) vs frame i+2 encoded on a low-priority side stream while
step i's decode / heads / value chain runs.  Prints ms per sequence for both and the worst rel-L2 between their outputs.
Run on the GPU box:  python tools/ab_overlap.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spann3r_b200 import Spann3R, synth  # noqa: E402

m = Spann3R(dus3r_name=None)
m.load_state_dict(synth.make_state_dict(sharpen=True), strict=True)
m = m.cuda().eval()
seqs = [[{"img": f["img"].cuda()} for f in synth.make_frames(10, 384, 512, seed0=1 + 100 * s)] for s in range(2)]
res, outs = {}, {}
for mode in (False, True, False, True):
    m.overlap_encoder = mode
    for i in range(3):
        m(seqs[i % 2])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(6):
        preds, _ = m(seqs[i % 2])
    e1.record()
    torch.cuda.synchronize()
    res.setdefault("overlap" if mode else "batched", []).append(e0.elapsed_time(e1) / 6)
    outs[mode] = [{k: v.clone() for k, v in p.items()} for p in m(seqs[0])[0]]
worst = max(float((a[k].double() - b[k].double()).norm() / b[k].double().norm()) for a, b in zip(outs[True], outs[False]) for k in b)
print(json.dumps({"ms_per_sequence": res, "worst_rel_l2_between_modes": worst}))
