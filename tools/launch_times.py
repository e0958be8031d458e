#!/usr/bin/env python
"""Per-launch CUDA-event times of ONE decode / value / heads call (engine profiling mode: launches are bracketed by events, so
PDL overlap is off and each duration stands alone), folded by position inside a layer.  Diagnostic: where a decoder layer's
~190 us go at B = 1.   python tools/launch_times.py"""
import json
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spann3r_b200 import Spann3R, synth  # noqa: E402

H, W = 384, 512
m = Spann3R(dus3r_name=None)
m.load_state_dict(synth.make_state_dict(sharpen=True), strict=True)
m = m.cuda().eval()
frames = [{"img": f["img"].cuda()} for f in synth.make_frames(3, H, W)]
m(frames)
eng = m._engine_for(1, H, W, n_frames=3)
imgs = torch.cat([f["img"] for f in frames[:2]]).contiguous()
feats = eng.encode(imgs)
f1, f2 = feats[:1].contiguous(), feats[1:2].contiguous()
out = {}


def run(name, fn, reps=5):
    fn()
    torch.cuda.synchronize()
    lists = []
    for _ in range(reps):
        eng.profile(True)
        fn()
        lst = eng.profile_list()
        eng.profile_read()
        eng.profile(False)
        lists.append(lst)
    n = len(lists[0])
    med = [statistics.median(l[i][0] for l in lists) * 1e3 for i in range(n)]     # us
    kinds = [lists[0][i][2] for i in range(n)]
    gfl = [lists[0][i][1] / 1e9 for i in range(n)]
    out[name] = {"launches": n, "sum_us": round(sum(med), 1)}
    return med, kinds, gfl


med, kinds, gfl = run("decode", lambda: eng.decode(f1, f2))
# decode: decoder_embed, qkv(0), then 12 x [attn, proj, q, attn, cproj, fc1, fc2, qkv(l+1)] (the last layer has no next qkv)
names = ["attn_self", "proj", "q", "attn_cross", "cproj", "fc1", "fc2", "qkv_next"]
pos = {k: [] for k in names}
i = 2
for l in range(12):
    for k in names:
        if k == "qkv_next" and l == 11:
            continue
        pos[k].append(med[i])
        i += 1
out["decode"]["embed_us"] = round(med[0], 1)
out["decode"]["qkv0_us"] = round(med[1], 1)
out["decode"]["per_position_median_us"] = {k: round(statistics.median(v), 1) for k, v in pos.items()}
out["decode"]["per_layer_sum_us"] = round(sum(statistics.median(v) for v in pos.values()), 1)
k1, k2 = eng.keyheads(f1, f2)
med, _, _ = run("heads", lambda: eng.heads())
out["heads"]["top10_us"] = sorted((round(x, 1) for x in med), reverse=True)[:10]
pts, conf = eng.heads()
med, _, _ = run("value", lambda: eng.value(pts[0], k1))
out["value"]["per_launch_us"] = [round(x, 1) for x in med]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, fn in (("decode", lambda: eng.decode(f1, f2)), ("heads", lambda: eng.heads()), ("value", lambda: eng.value(pts[0], k1))):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    out[name]["pipelined_us"] = round(e0.elapsed_time(e1) * 100, 1)      # per call, PDL overlap on
print(json.dumps(out))
