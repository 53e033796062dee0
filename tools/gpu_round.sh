#!/bin/bash
# One GPU-box visit: bring-up diagnostics, the GPU test-suite, a short bench and the ncu launch list.
# Everything lands in gpurun_out/ (merged back by gpurun).
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.csv 2>&1
timeout 900 python tools/gpu_debug_conv.py > gpurun_out/debug_conv.log 2>&1
echo "debug_conv rc=$?"
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest.log 2>&1
echo "pytest rc=$?"
tail -5 gpurun_out/pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --profile-ops > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?"
cat gpurun_out/bench.json | head -c 3000
