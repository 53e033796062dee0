#!/bin/bash
# One GPU-box visit: bring-up diagnostics, the GPU test-suite, a short bench and (NCU=1) the ncu evidence.
# Everything lands in gpurun_out/ (merged back by gpurun).
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.csv 2>&1
if [ "${DEBUGCONV:-1}" = "1" ]; then
  timeout 900 python tools/gpu_debug_conv.py > gpurun_out/debug_conv.log 2>&1
  echo "debug_conv rc=$?"
fi
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest.log 2>&1
echo "pytest rc=$?"
tail -3 gpurun_out/pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --profile-ops --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?"
cat gpurun_out/bench.json | head -c 600
timeout 300 python bench.py --steps 20 --warmup 3 --precision fp16 --no-cpu-baseline > gpurun_out/bench_fp16.json 2> gpurun_out/bench_fp16.err
timeout 600 python bench.py --steps 20 --warmup 3 --precision fp16f8 --profile-ops --no-cpu-baseline > gpurun_out/bench_fp16f8.json 2> gpurun_out/bench_fp16f8.err
echo "bench fp16f8 rc=$?"
cat gpurun_out/bench_fp16f8.json | head -c 400
timeout 300 python tools/bench_layout.py 4 > gpurun_out/bench_layout.json 2> gpurun_out/bench_layout.err
if [ "${ATTN:-0}" = "1" ]; then
  (B200_FA_FFMA=1 python tools/gpu_bench_attn.py; python tools/gpu_bench_attn.py) > gpurun_out/attn_bench.txt 2>&1
  cat gpurun_out/attn_bench.txt
fi
if [ "${SPLIT:-0}" = "1" ]; then
  timeout 600 python tools/exp_split.py fp16f8 8 > gpurun_out/exp_split.txt 2>&1
  cat gpurun_out/exp_split.txt | tail -4
fi
if [ "${NCU:-0}" = "1" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --profiler-range > gpurun_out/ncu_bench.log 2>&1
  echo "ncu launches rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_tc -s 4 -c 4 -f -o gpurun_out/prof_conv \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --profiler-range > gpurun_out/ncu_conv.log 2>&1
  echo "ncu conv rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'gn_act|in_conv|out_conv|fir_|flash_attn|sampler' -c 12 -f -o gpurun_out/prof_elem \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --profiler-range > gpurun_out/ncu_elem.log 2>&1
  echo "ncu elem rc=$?"
fi
