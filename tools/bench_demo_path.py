#!/usr/bin/env python
"""The whole `demo.py` per-sequence pipeline on the GPU, stage by stage: decoded 1080p uint8 frames (pinned host) ->
input adapter (Pillow-exact Lanczos + ImgNorm) -> Spann3R.forward -> Weiszfeld focal -> PnP-RANSAC pose of every frame ->
poses / focal read back to the host.  Next to it: what the reference does on the CPU around the network for the same
frames (PIL preprocessing per frame, demo.py:57-86; cv2.solvePnPRansac per frame, demo.py:166-180), bounded sample.
One JSON line.  Run on the GPU box:  python tools/bench_demo_path.py [--frames 10]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spann3r_b200 import Spann3R, synth  # noqa: E402
from spann3r_b200 import preprocess as P  # noqa: E402
from spann3r_b200.postprocess import estimate_focal_knowing_depth, solve_pnp_ransac  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=10)
ap.add_argument("--reps", type=int, default=5)
args = ap.parse_args()
F_, H0, W0, RES = args.frames, 1080, 1920, (512, 384)
rng = np.random.default_rng(0)
raw = [torch.from_numpy(rng.integers(0, 256, (H0, W0, 3), dtype=np.uint8)).pin_memory() for _ in range(F_)]
model = Spann3R(dus3r_name=None)
model.load_state_dict(synth.make_state_dict(sharpen=True), strict=True)
model = model.cuda().eval()
adapter = P.FrameAdapter(RES)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]


def run():
    ev[0].record()
    views = P.load_frames(raw, RES, adapter=adapter)
    ev[1].record()
    preds, _ = model(views)
    ev[2].record()
    _, H, W, _ = preds[0]["pts3d"].shape
    focal = estimate_focal_knowing_depth(preds[0]["pts3d"], (W / 2, H / 2), focal_mode="weiszfeld")
    f = float(focal[0])                      # demo.py needs the value to build the intrinsic matrix (one small sync)
    if not (f == f and 1.0 < f < 1e6):       # random-init pointmaps can give a meaningless focal; keep the pipeline going
        f = 400.0
    ev[3].record()
    K = [[f, 0, W / 2], [0, f, H / 2], [0, 0, 1]]
    pts = torch.cat([preds[j]["pts3d" if j == 0 else "pts3d_in_other_view"] for j in range(len(preds))])
    ok, rvec, tvec, inl = solve_pnp_ransac(pts, K)
    poses = torch.cat((rvec, tvec), 1).cpu()
    ev[4].record()
    torch.cuda.synchronize()
    return [ev[i].elapsed_time(ev[i + 1]) for i in range(4)], poses, ok


for _ in range(3):
    run()
acc = np.zeros(4)
t0 = time.time()
for _ in range(args.reps):
    ms, poses, ok = run()
    acc += ms
wall = (time.time() - t0) / args.reps
acc /= args.reps
res = {"metric": "demo.py pipeline frames/s (1080p frames -> adapter -> Spann3R.forward -> focal -> PnP poses on the host)",
       "value": F_ / wall, "unit": "frames/s", "frames": F_, "wall_ms_per_sequence": wall * 1e3,
       "stage_ms": {"adapter": acc[0], "forward": acc[1], "focal": acc[2], "pnp_and_readback": acc[3]},
       "note": "random-init weights give pointmaps a camera cannot explain, so PnP reports few inliers here; the work per "
               "frame (400 hypotheses x all points, 16 refinement passes) is the same as for a real reconstruction",
       "pnp_success": [bool(v) for v in ok.tolist()]}
# the reference's CPU work around the network, bounded sample (2 frames each)
try:
    import cv2
    import PIL.Image
    g = P.plan_frame(H0, W0, RES)
    t0 = time.time()
    for i in range(2):
        im = PIL.Image.fromarray(raw[i].numpy()).crop(g["crop1"]).resize(g["scaled"], resample=PIL.Image.Resampling.LANCZOS).crop(g["crop2"])
        _ = (np.asarray(im).astype(np.float32) / 255.0 - 0.5) / 0.5
    pre_ms = (time.time() - t0) / 2 * 1e3
    case = synth.PNP_CASES[1]
    pts, K = synth.make_pointmap_case(*case)
    u, v = np.meshgrid(np.arange(case[1]), np.arange(case[0]))
    p2 = np.stack((u, v), -1).reshape(-1, 2).astype(np.float32)
    t0 = time.time()
    for i in range(2):
        cv2.solvePnPRansac(pts.reshape(-1, 3), p2, K.astype(np.float32), np.zeros(4, np.float32))
    pnp_ms = (time.time() - t0) / 2 * 1e3
    res["reference_cpu_ms_per_frame"] = {"pil_preprocess": pre_ms, "cv2_solvePnPRansac": pnp_ms,
                                         "cores": len(os.sched_getaffinity(0)), "sample": "2 frames each"}
except Exception as ex:
    res["reference_cpu_ms_per_frame"] = {"unavailable": repr(ex)[:120]}
print(json.dumps(res))
