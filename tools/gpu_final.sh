#!/bin/bash
# Round-end evidence: GPU test-suite, default bench (with cpu_baseline), reference arm, per-kernel profile, layout bench,
# ncu launch list of the bench command + one full capture of the dominant conv launches.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.csv 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -s > gpurun_out/pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench default rc=$?"; head -c 900 gpurun_out/bench_default.json; echo
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
echo "bench reference rc=$?"; head -c 300 gpurun_out/bench_reference.json; echo
timeout 600 python bench.py --steps 20 --warmup 3 --profile-ops --no-cpu-baseline > gpurun_out/bench_ops.json 2> gpurun_out/bench_ops.err
echo "bench ops rc=$?"
timeout 600 python bench.py --steps 20 --warmup 3 --precision fp16f8 --no-cpu-baseline > gpurun_out/bench_fp16f8.json 2> gpurun_out/bench_fp16f8.err
echo "bench fp16f8 rc=$?"; head -c 200 gpurun_out/bench_fp16f8.json; echo
timeout 300 python tools/bench_layout.py 4 > gpurun_out/bench_layout.json 2> gpurun_out/bench_layout.err
echo "layout rc=$?"; head -c 300 gpurun_out/bench_layout.json; echo
if [ "${NCU:-1}" = "1" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --profiler-range > gpurun_out/ncu_bench.log 2>&1
  echo "ncu launches rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_tc -s 4 -c 4 -f -o gpurun_out/prof_conv \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --profiler-range > gpurun_out/ncu_conv.log 2>&1
  echo "ncu conv rc=$?"
fi
