#!/bin/bash
# One GPU-box visit: bring-up diagnostics, the GPU test-suite, a short bench and the ncu evidence.
# Everything lands in gpurun_out/ (merged back by gpurun).
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.csv 2>&1
timeout 900 python tools/gpu_debug_conv.py > gpurun_out/debug_conv.log 2>&1
echo "debug_conv rc=$?"
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest.log 2>&1
echo "pytest rc=$?"
tail -5 gpurun_out/pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --profile-ops --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?"
cat gpurun_out/bench.json | head -c 1500
cat gpurun_out/bench.err | head -20
timeout 300 python bench.py --steps 20 --warmup 3 --precision fp16 --profile-ops --no-cpu-baseline > gpurun_out/bench_fp16.json 2> gpurun_out/bench_fp16.err
if [ "${NCU:-1}" = "1" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
  echo "ncu launches rc=$?"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 2 -c 2 -f -o gpurun_out/prof_conv \
      python tools/gpu_debug_conv.py '{"taps": 9, "Cin": 64, "Cout": 64, "bn": 64, "rows": 2, "parts": 2, "B": 8, "H": 32, "W": 1024}' > gpurun_out/ncu_conv.log 2>&1
  echo "ncu full rc=$?"
fi
timeout 300 python tools/bench_layout.py 4 > gpurun_out/bench_layout.json 2> gpurun_out/bench_layout.err
