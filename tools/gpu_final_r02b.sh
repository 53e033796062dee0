#!/bin/bash
# Round 2, final evidence after the column-walk conv: GPU test-suite, the bench lines the driver runs (default + reference arm),
# per-kernel profile, clip / rollout lines, ncu launch list + full captures (column-walk conv, tile-walk convs).  -> gpurun_out/
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r02b_pytest_gpu.txt 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/r02b_pytest_gpu.txt
export B200_TUNE_FILE=gpurun_out/r02b_tune.json
rm -f $B200_TUNE_FILE gpurun_out/r02b_col_traffic.json
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > /dev/null 2>&1     # measures + persists the tile / front-end choices
if [ "${NCU:-1}" = "1" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_col -c 4 -f -o gpurun_out/r02b_prof_col \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --profiler-range > gpurun_out/ncu_col.log 2>&1
  echo "ncu col rc=$?"
  python tools/ncu_traffic.py gpurun_out/r02b_prof_col.ncu-rep conv_col gpurun_out/r02b_col_traffic.json
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/r02b_ncu_launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --profiler-range > gpurun_out/ncu_bench.log 2>&1
  echo "ncu launches rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_tc -s 8 -c 6 -f -o gpurun_out/r02b_prof_conv \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --profiler-range > gpurun_out/ncu_conv.log 2>&1
  echo "ncu conv rc=$?"
fi
timeout 600 python bench.py > gpurun_out/r02b_bench_default.json 2> gpurun_out/r02b_bench_default.err
echo "bench default rc=$?"; head -c 700 gpurun_out/r02b_bench_default.json; echo
timeout 600 python bench.py --impl reference --steps 50 --warmup 3 > gpurun_out/r02b_bench_reference.json 2> gpurun_out/r02b_bench_reference.err
echo "bench reference rc=$?"; head -c 300 gpurun_out/r02b_bench_reference.json; echo
timeout 600 python bench.py --steps 20 --warmup 3 --profile-ops --no-cpu-baseline > gpurun_out/r02b_bench_ops.json 2> gpurun_out/r02b_bench_per_kernel.txt
echo "bench ops rc=$?"
timeout 600 python bench.py --steps 20 --warmup 3 --batch-per-gpu 1 --profile-ops --no-cpu-baseline > gpurun_out/r02b_bench_b1.json 2> gpurun_out/r02b_bench_per_kernel_b1.txt
echo "bench b1 rc=$?"; python -c "import json; d=json.load(open('gpurun_out/r02b_bench_b1.json')); print(d['ms_per_step'], d['value'])"
timeout 300 python tools/bench_layout.py 4 > gpurun_out/r02b_bench_layout_b4.json 2> /dev/null
echo "layout rc=$?"; head -c 300 gpurun_out/r02b_bench_layout_b4.json; echo
timeout 600 python bench.py --workload clip --steps 20 > gpurun_out/r02b_bench_clip.json 2> gpurun_out/r02b_bench_clip.err
echo "clip rc=$?"; head -c 300 gpurun_out/r02b_bench_clip.json; echo
timeout 900 python bench.py --workload rollout --steps 20 > gpurun_out/r02b_bench_rollout.json 2> gpurun_out/r02b_bench_rollout.err
echo "rollout rc=$?"; head -c 300 gpurun_out/r02b_bench_rollout.json; echo
