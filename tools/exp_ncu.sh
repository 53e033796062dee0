#!/bin/bash
# ncu evidence of the bench command.  The tile autotune is measured once by a plain bench run and persisted
# (B200_TUNE_FILE), so the profiled processes run exactly the tiles the bench measured.
mkdir -p gpurun_out
export B200_TUNE_FILE=gpurun_out/tune.json
rm -f $B200_TUNE_FILE
timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --profile-ops > gpurun_out/bench_tuned.json 2> gpurun_out/bench_tuned.err
echo "bench rc=$?"; python -c "import json; d=json.load(open('gpurun_out/bench_tuned.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'])"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --profiler-range > gpurun_out/ncu_bench.log 2>&1
echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_tc -s 4 -c 4 -f -o gpurun_out/prof_conv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --profiler-range > gpurun_out/ncu_conv.log 2>&1
echo "ncu conv rc=$?"
