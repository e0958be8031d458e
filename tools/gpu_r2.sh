#!/bin/bash
# Round-2 evidence script (one gpurun call, every step under its own timeout).  Usage, from the repo root through gpurun:
#   gpurun --timeout 2400 -- 'bash tools/gpu_r2.sh r2a unverified tests ab bench ncu'
# steps: unverified = tests still marked gpu_unverified; tests = full `pytest -m gpu`; ab = in-situ A/B of the switches below
#        (quick bench legs, same box); bench = full bench.py line (real-reference CPU + eager-GPU legs); ncu = launch list of
#        one sequence; smoke = __graft_entry__.smoke() under an ncu launch list (what the driver records); multi = 2-GPU tests
tag=${1:-r2}; shift
mkdir -p gpurun_out
quick() {   # quick NAME [ENV=VAL ...]: one short bench leg, prints value / ms / gemm ms
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 8 --warmup 3 --no-eager-gpu --no-cpu-baseline --no-raw 2> gpurun_out/${tag}_ab_${name}.err \
    | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('$name', round(d['value'],1), 'frames/s', round(d['ms_per_step'],2), 'ms  e2e', round(d['e2e']['value'],1), ' gemm', round(r['gemm_ms_per_seq'],2), 'ms attn', round(r['attention_ms_per_seq'],2), 'ms launches', d['gpu_launches'])" \
    | tee -a gpurun_out/${tag}_ab.txt
}
for step in "$@"; do
  case $step in
    unverified) timeout 400 python -m pytest tests -q -m gpu_unverified > gpurun_out/${tag}_unverified.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_unverified.log; tail -12 gpurun_out/${tag}_unverified.log ;;
    tests) timeout 1200 python -m pytest tests -q -m gpu --durations=12 > gpurun_out/${tag}_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_tests.log; tail -30 gpurun_out/${tag}_tests.log ;;
    ab)    quick base
           quick prefetch_off S3R_PREFETCH_B=0
           quick pair64 S3R_GEMM2_64=1
           quick base_again ;;
    e2e)   quick plain
           quick omp64 OMP_NUM_THREADS=64
           quick omp1 OMP_NUM_THREADS=1
           quick plain_again ;;
    abin)  timeout 600 python tools/ab_inproc.py gemm2_64=0,1 prefetch_b=0,1 attn_pair=0,1 --rounds 12 > gpurun_out/${tag}_ab_inproc.jsonl 2> gpurun_out/${tag}_ab_inproc.err
           cat gpurun_out/${tag}_ab_inproc.jsonl; tail -2 gpurun_out/${tag}_ab_inproc.err ;;
    chain) timeout 600 python -m pytest tests/test_chain_gpu.py -q -m gpu -x -s > gpurun_out/${tag}_chain_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_chain_tests.log; tail -30 gpurun_out/${tag}_chain_tests.log | cut -c1-400 ;;
    stages) for c in 0 1; do echo "S3R_CHAIN=$c"; S3R_CHAIN=$c timeout 300 python tools/stage_times.py 2>/dev/null | tee -a gpurun_out/${tag}_stage_times_chain$c.json; done ;;
    ab256) timeout 600 python tools/ab_inproc.py gemm2_256=0,1 --rounds 12 > gpurun_out/${tag}_ab_256.jsonl 2> gpurun_out/${tag}_ab_256.err; cat gpurun_out/${tag}_ab_256.jsonl; tail -2 gpurun_out/${tag}_ab_256.err ;;
    abchain) timeout 600 python tools/ab_inproc.py chain=0,1 --rounds 12 > gpurun_out/${tag}_ab_chain.jsonl 2> gpurun_out/${tag}_ab_chain.err; cat gpurun_out/${tag}_ab_chain.jsonl; tail -2 gpurun_out/${tag}_ab_chain.err ;;
    train) timeout 900 python -m pytest tests/test_train_gpu.py tests/test_model_gpu.py -q -m gpu -k "train or raw or dropout or backward" -s > gpurun_out/${tag}_train_tests.log 2>&1
           echo "pytest exit $?" >> gpurun_out/${tag}_train_tests.log; tail -25 gpurun_out/${tag}_train_tests.log ;;
    nativelin) timeout 600 python -m pytest tests/test_train_gpu.py -q -m gpu -k "native_linear" -s > gpurun_out/${tag}_native_linear_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_native_linear_tests.log; tail -12 gpurun_out/${tag}_native_linear_tests.log | cut -c1-300
           for v in 0 1; do S3R_TRAIN_NATIVE_LINEAR=$v timeout 400 python tools/train_step_bench.py --impl ours --steps 3 --warmup 2 2> gpurun_out/${tag}_train1_native$v.err | grep '"what"' | tee gpurun_out/${tag}_train1_native$v.json | cut -c1-700; done ;;
    trainbench) # config 5 on the GPUs of this call (N = CUDA device count): ours, then the reference, NCCL algorithm from its log
           N=$(python -c "import torch; print(torch.cuda.device_count())")
           for impl in ours reference; do
             NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL,TUNING timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
               tools/train_step_bench.py --impl $impl --steps 4 --warmup 2 > gpurun_out/${tag}_train_${impl}_n${N}.log 2>&1
             grep '"what"' gpurun_out/${tag}_train_${impl}_n${N}.log | tee gpurun_out/${tag}_train_${impl}_n${N}.json
             grep -E "NVLS|Ring|Tree|Algo|algo" gpurun_out/${tag}_train_${impl}_n${N}.log | sort | uniq -c | sort -rn | head -8 > gpurun_out/${tag}_train_${impl}_n${N}_nccl.txt
             head -4 gpurun_out/${tag}_train_${impl}_n${N}_nccl.txt; tail -3 gpurun_out/${tag}_train_${impl}_n${N}.log | cut -c1-300
           done ;;
    config3) N=$(python -c "import torch; print(torch.cuda.device_count())")
           timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 5 --warmup 3 --config3 \
             > gpurun_out/${tag}_bench_config3_n${N}.json 2> gpurun_out/${tag}_bench_config3_n${N}.err; tail -2 gpurun_out/${tag}_bench_config3_n${N}.err | cut -c1-300
           python -c "import json; d=json.loads(open('gpurun_out/${tag}_bench_config3_n${N}.json').read().strip().splitlines()[-1]); print(d['n_gpus'], round(d['value'],1), 'frames/s', d['config3'])" ;;
    bench) timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -3 gpurun_out/${tag}_bench.err; cut -c1-1500 gpurun_out/${tag}_bench.json ;;
    ncu)   timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
             --profile-from-start off --csv --log-file gpurun_out/${tag}_launches.csv python tools/profile_seq.py > gpurun_out/${tag}_ncu.log 2>&1
           python tools/summarize_ncu.py gpurun_out/${tag}_launches.csv --title "${tag}: ncu launch list of ONE 10-frame 512x384 sequence (B=1)" \
             > gpurun_out/${tag}_launches.md 2>> gpurun_out/${tag}_ncu.log; head -30 gpurun_out/${tag}_launches.md ;;
    ncufull) # `--set full` of a run of step-phase launches (decoder layer: qkv, attention, proj, q, attention, cproj, fc1, fc2 ...)
           timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:'gemm2?_bf16x3_kernel|attention_kernel|upsample2x' \
             --launch-skip ${NCU_SKIP:-150} --launch-count ${NCU_COUNT:-18} -f -o gpurun_out/${tag}_full python tools/profile_seq.py > gpurun_out/${tag}_ncufull.log 2>&1
           ncu -i gpurun_out/${tag}_full.ncu-rep --page raw --csv > gpurun_out/${tag}_full_raw.csv 2>> gpurun_out/${tag}_ncufull.log
           python tools/ncu_extract.py gpurun_out/${tag}_full_raw.csv --title "${tag}: ncu --set full, ${NCU_COUNT:-18} launches of the step phase (B=1, 512x384)" > gpurun_out/${tag}_ncu_full.md 2>> gpurun_out/${tag}_ncufull.log
           ls -la gpurun_out/${tag}_full.ncu-rep | cut -c20-80; head -12 gpurun_out/${tag}_ncu_full.md | cut -c1-260 ;;
    trace2) timeout 300 python tools/trace_gemm2.py > gpurun_out/${tag}_trace_gemm2.txt 2> gpurun_out/${tag}_trace_gemm2.err; cat gpurun_out/${tag}_trace_gemm2.txt | cut -c1-400; tail -2 gpurun_out/${tag}_trace_gemm2.err ;;
    ltimes) timeout 300 python tools/launch_times.py > gpurun_out/${tag}_launch_times.json 2> gpurun_out/${tag}_launch_times.err; cat gpurun_out/${tag}_launch_times.json | cut -c1-1500; tail -2 gpurun_out/${tag}_launch_times.err ;;
    cfg4)  timeout 300 python bench.py --frames 100 --steps 2 --warmup 3 --no-eager-gpu --no-cpu-baseline --no-raw > gpurun_out/${tag}_bench_config4.json 2> gpurun_out/${tag}_bench_config4.err
           python -c "import json; d=json.loads(open('gpurun_out/${tag}_bench_config4.json').read().strip().splitlines()[-1]); print('config 4 (100 frames):', round(d['value'],1), 'frames/s', round(d['ms_per_step'],1), 'ms per sequence, e2e', round(d['e2e']['value'],1))" ;;
    ncusrc) # source-level stall profile of ONE fc1-sized launch (gemm2<128, EPI_PLAIN>) of the decoder: where do the epilogue warps wait?
           timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off --kernel-name-base demangled -k regex:'gemm2_bf16x3_kernel<.int.128, .int.0>' \
             --launch-skip ${NCU_SKIP:-8} --launch-count 1 -f -o gpurun_out/${tag}_src python tools/profile_seq.py 3 > gpurun_out/${tag}_ncusrc.log 2>&1
           ncu -i gpurun_out/${tag}_src.ncu-rep --page source --csv > gpurun_out/${tag}_src_page.csv 2>> gpurun_out/${tag}_ncusrc.log
           ncu -i gpurun_out/${tag}_src.ncu-rep --page raw --csv > gpurun_out/${tag}_src_raw.csv 2>> gpurun_out/${tag}_ncusrc.log
           ls -la gpurun_out/${tag}_src* | cut -c25-100; head -3 gpurun_out/${tag}_src_page.csv | cut -c1-400 ;;
    b8)    timeout 400 python bench.py --batch 8 --steps 4 --warmup 3 --no-eager-gpu --no-cpu-baseline --no-raw > gpurun_out/${tag}_bench_b8.json 2> gpurun_out/${tag}_bench_b8.err
           python -c "import json; d=json.loads(open('gpurun_out/${tag}_bench_b8.json').read().strip().splitlines()[-1]); r=d['roofline']; print('B=8 lockstep, 1 GPU:', round(d['value'],1), 'frames/s', round(d['ms_per_step'],1), 'ms per 8 sequences; e2e', round(d['e2e']['value'],1), '; GEMM engine', round(r['achieved'],1), 'TFLOP/s frac', round(r['frac'],3), 'whole path', round(r['whole_path_frac'],3))" ;;
    smoke) timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/${tag}_smoke_launches.csv \
             python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -2 gpurun_out/${tag}_smoke.log
           python - <<PY
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/${tag}_smoke_launches.csv")) if len(r) > 5 and r[0].isdigit()]
c = collections.Counter(r[4].split("(")[0][:70] for r in rows)
print("first 1000 launches of smoke():", len(rows)); [print(" ", n, k) for k, n in c.most_common(14)]
PY
           ;;
    multi) timeout 600 python -m pytest tests/test_multi_gpu.py -q -m gpu > gpurun_out/${tag}_multi.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_multi.log; tail -8 gpurun_out/${tag}_multi.log ;;
    *) echo "unknown step $step" ;;
  esac
done
