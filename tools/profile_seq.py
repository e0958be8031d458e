#!/usr/bin/env python
"""Run ONE 10-frame 512x384 sequence between cudaProfilerStart/Stop (use with `ncu --profile-from-start off`).
Never a source of benchmark numbers (bench.py is); it only gives ncu a warm, bounded region."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spann3r_b200 import Spann3R, synth  # noqa: E402

F_ = int(sys.argv[1]) if len(sys.argv) > 1 else 10
sd = synth.make_state_dict(sharpen=True)
m = Spann3R(dus3r_name=None)
m.load_state_dict(sd, strict=True)
m = m.cuda().eval()
frames = [{"img": f["img"].cuda()} for f in synth.make_frames(F_, 384, 512)]
m(frames)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
m(frames)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
