#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'out_conv|flash_attn' -c 3 -f -o gpurun_out/prof_out_attn \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --profiler-range > gpurun_out/ncu_out_attn.log 2>&1
echo "ncu rc=$?"
