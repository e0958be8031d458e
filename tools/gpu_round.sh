#!/bin/bash
# One gpurun call that produces everything a round's evidence needs, each step under its own timeout so a hang in one does
# not eat the call.  Usage (from the repo root, through gpurun):
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh r2 tests bench ab ncu'
# steps: tests = full `pytest -m gpu` (+ durations), bench = bench.py (N=1), ab = tools/ab_overlap.py,
#        ncu = launch list of one sequence (+ DRAM bytes) summarised by tools/summarize_ncu.py, demo = tools/bench_demo_path.py,
#        ab64 = bench.py with S3R_GEMM2_64=0/1/0 (256 x 64 CTA-pair tiles where the planner picks 1-CTA 128 x 64),
#        unverified = the tests marked gpu_unverified (kernels written without a GPU), sanitize = compute-sanitizer memcheck /
#        racecheck / synccheck over the op-level tests (slow: minutes per tool)
# Outputs: gpurun_out/<tag>_*.  Copy what should be judged into profiles/ afterwards.  Round-1 timings for budgeting: the
# 11 golden / portrait / PnP tests 147 s, bench.py 60-70 s (incl. the eager-GPU and CPU-baseline legs), bench_demo_path 40 s,
# first `import torch` on a fresh box up to 60 s, ncu launch list of one sequence ~3 min.
tag=${1:-rX}; shift
mkdir -p gpurun_out
for step in "$@"; do
  case $step in
    tests) timeout 900 python -m pytest tests -q -m gpu -x --durations=15 > gpurun_out/${tag}_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_tests.log; tail -25 gpurun_out/${tag}_tests.log ;;
    bench) timeout 240 python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -3 gpurun_out/${tag}_bench.err; cut -c1-900 gpurun_out/${tag}_bench.json ;;
    ab)    timeout 240 python tools/ab_overlap.py > gpurun_out/${tag}_ab_overlap.json 2> gpurun_out/${tag}_ab_overlap.err; tail -3 gpurun_out/${tag}_ab_overlap.err; cat gpurun_out/${tag}_ab_overlap.json ;;
    ab64)  # in-situ A/B of the 256 x 64 pair-tile experiment (env read once per process): default, variant, default again
           for v in 0 1 0; do S3R_GEMM2_64=$v timeout 200 python bench.py --steps 8 --warmup 3 --no-eager-gpu --no-cpu-baseline 2>/dev/null \
             | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('S3R_GEMM2_64=$v', round(d['value'],1), 'frames/s', round(d['ms_per_step'],2), 'ms  gemm', round(d['roofline']['gemm_ms_per_seq'],2), 'ms')" \
             | tee -a gpurun_out/${tag}_ab64.txt; done ;;
    demo)  timeout 200 python tools/bench_demo_path.py > gpurun_out/${tag}_bench_demo_path.json 2> gpurun_out/${tag}_demo.err; tail -3 gpurun_out/${tag}_demo.err; cat gpurun_out/${tag}_bench_demo_path.json ;;
    ncu)   timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
             --profile-from-start off --csv --log-file gpurun_out/${tag}_launches.csv python tools/profile_seq.py > gpurun_out/${tag}_ncu.log 2>&1
           python tools/summarize_ncu.py gpurun_out/${tag}_launches.csv --title "${tag}: ncu launch list of ONE 10-frame 512x384 sequence (B=1)" \
             > gpurun_out/${tag}_launches.md 2>> gpurun_out/${tag}_ncu.log; head -20 gpurun_out/${tag}_launches.md ;;
    sanitize) # memcheck + racecheck + synccheck on the small op-level tests (tcgen05 / TMA / mbarrier kernels) and one 224x224 sequence
           for tool in memcheck racecheck synccheck; do
             timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_ops_gpu.py tests/test_pnp.py -q -m gpu -x -k "not 7680" \
               > gpurun_out/${tag}_sanitizer_${tool}.log 2>&1; echo "$tool exit $?" | tee -a gpurun_out/${tag}_sanitizer_${tool}.log
             grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/${tag}_sanitizer_${tool}.log | tail -3
           done ;;
    unverified) timeout 300 python -m pytest tests -q -m gpu_unverified > gpurun_out/${tag}_unverified.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_unverified.log; tail -8 gpurun_out/${tag}_unverified.log ;;
    *) echo "unknown step $step" ;;
  esac
done
