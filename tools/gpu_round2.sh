#!/bin/bash
# One GPU-box visit (session 2 of round 1): GPU test-suite, the default bench (with cpu_baseline), the reference arm,
# per-shape profiles; NCU=1 adds the ncu launch list + full captures.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.csv 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -s > gpurun_out/pytest.log 2>&1
echo "pytest rc=$?"
tail -5 gpurun_out/pytest.log
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench default rc=$?"
head -c 1500 gpurun_out/bench_default.json; echo
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
echo "bench reference rc=$?"
head -c 700 gpurun_out/bench_reference.json; echo
timeout 600 python bench.py --steps 20 --warmup 3 --profile-ops --no-cpu-baseline > gpurun_out/bench_ops.json 2> gpurun_out/bench_ops.err
echo "bench ops rc=$?"
timeout 600 python bench.py --steps 20 --warmup 3 --precision fp16f8 --profile-ops --no-cpu-baseline > gpurun_out/bench_fp16f8.json 2> gpurun_out/bench_fp16f8.err
echo "bench fp16f8 rc=$?"; head -c 300 gpurun_out/bench_fp16f8.json; echo
timeout 300 python tools/bench_layout.py 4 > gpurun_out/bench_layout.json 2> gpurun_out/bench_layout.err
echo "layout rc=$?"; head -c 400 gpurun_out/bench_layout.json; echo
if [ -n "${EXTRA:-}" ]; then
  timeout 900 bash -c "$EXTRA" > gpurun_out/extra.log 2>&1
  echo "extra rc=$?"; tail -20 gpurun_out/extra.log
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
