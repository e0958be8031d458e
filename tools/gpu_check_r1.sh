#!/bin/bash
# One gpurun call: the GPU tests added late in round 1 (portrait, mem_pos_enc, PnP) + the golden model tests, then the
# PnP / demo-pipeline / headline benches.  Logs under gpurun_out/.
mkdir -p gpurun_out
timeout 330 python -m pytest tests/test_pnp.py tests/test_model_gpu.py -q -m gpu -s -k "pnp or golden or portrait or mem_pos or stagewise" \
  > gpurun_out/r1_new_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/r1_new_tests.log
tail -5 gpurun_out/r1_new_tests.log
timeout 100 python tools/bench_pnp.py > gpurun_out/r1_bench_pnp.json 2> gpurun_out/bench_pnp.err; tail -2 gpurun_out/bench_pnp.err; cat gpurun_out/r1_bench_pnp.json
timeout 150 python bench.py --steps 5 --warmup 3 > gpurun_out/r1_bench_last.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err; cut -c1-600 gpurun_out/r1_bench_last.json
timeout 120 python tools/bench_demo_path.py > gpurun_out/r1_bench_demo_path.json 2> gpurun_out/demo.err; tail -2 gpurun_out/demo.err; cat gpurun_out/r1_bench_demo_path.json
