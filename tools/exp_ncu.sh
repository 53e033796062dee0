#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --profiler-range > gpurun_out/ncu_bench.log 2>&1
echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_tc -s 4 -c 6 -f -o gpurun_out/prof_conv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --profiler-range > gpurun_out/ncu_conv.log 2>&1
echo "ncu conv rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'gn_act|in_conv|out_conv|fir_|flash_attn|sampler' -c 30 -f -o gpurun_out/prof_elem \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --profiler-range > gpurun_out/ncu_elem.log 2>&1
echo "ncu elem rc=$?"
ls -la gpurun_out/
