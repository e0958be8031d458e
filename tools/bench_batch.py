#!/usr/bin/env python
"""Run bench.py at several per-GPU batch sizes (BASELINE config[2]-style throughput) and print one line each."""
import json
import subprocess
import sys

for b in [int(x) for x in (sys.argv[1:] or ["2", "4", "8"])]:
    r = subprocess.run([sys.executable, "bench.py", "--steps", "3", "--warmup", "3", "--no-cpu-baseline", "--batch", str(b)],
                       capture_output=True, text=True)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        print(json.dumps({"batch": b, "frames_per_s": round(d["value"], 1), "e2e": round(d["e2e"]["value"], 1),
                          "ms_per_step": round(d["ms_per_step"], 2), "gemm_tflops": round(d["roofline"]["achieved"], 1),
                          "whole_path_frac": round(d["roofline"]["whole_path_frac"], 3)}), flush=True)
    except Exception as e:  # noqa
        print("batch", b, "failed:", r.stderr[-800:], flush=True)
