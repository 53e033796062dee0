#!/bin/bash
# Round-2 evidence in one GPU-box visit: GPU test-suite, the bench lines the driver runs (default + reference arm), per-kernel
# profile, the other precision, the layout / clip / rollout workloads, and the ncu passes (launch list of the bench command +
# full captures of the conv and attention kernels).  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.csv 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r02_pytest_gpu.txt 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/r02_pytest_gpu.txt
export B200_TUNE_FILE=gpurun_out/r02_tune.json
rm -f $B200_TUNE_FILE
timeout 600 python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err
echo "bench default rc=$?"; head -c 600 gpurun_out/r02_bench_default.json; echo
timeout 600 python bench.py --impl reference --steps 50 --warmup 3 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
echo "bench reference rc=$?"; head -c 300 gpurun_out/r02_bench_reference.json; echo
timeout 600 python bench.py --steps 20 --warmup 3 --profile-ops --no-cpu-baseline > gpurun_out/r02_bench_ops.json 2> gpurun_out/r02_bench_per_kernel.txt
echo "bench ops rc=$?"
timeout 600 python bench.py --steps 50 --warmup 3 --precision fp16x3 --no-cpu-baseline > gpurun_out/r02_bench_fp16x3.json 2> /dev/null
echo "bench fp16x3 rc=$?"; head -c 200 gpurun_out/r02_bench_fp16x3.json; echo
timeout 300 python tools/bench_layout.py 4 > gpurun_out/r02_bench_layout_b4.json 2> /dev/null
echo "layout rc=$?"; head -c 400 gpurun_out/r02_bench_layout_b4.json; echo
timeout 600 python bench.py --workload clip > gpurun_out/r02_bench_clip.json 2> gpurun_out/r02_bench_clip.err
echo "clip rc=$?"; head -c 300 gpurun_out/r02_bench_clip.json; echo
timeout 900 python bench.py --workload rollout > gpurun_out/r02_bench_rollout.json 2> gpurun_out/r02_bench_rollout.err
echo "rollout rc=$?"; head -c 300 gpurun_out/r02_bench_rollout.json; echo
COUNTERS=1 timeout 120 python tools/gpu_bench_attn.py > gpurun_out/r02_attn_bench.txt 2>&1
B200_FA_IMPL=mma timeout 120 python tools/gpu_bench_attn.py > gpurun_out/r02_attn_bench_mma.txt 2>&1
if [ "${NCU:-1}" = "1" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/r02_ncu_launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --profiler-range > gpurun_out/ncu_bench.log 2>&1
  echo "ncu launches rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_tc -s 8 -c 6 -f -o gpurun_out/r02_prof_conv \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --profiler-range > gpurun_out/ncu_conv.log 2>&1
  echo "ncu conv rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'flash_attn_tc|attn_pack' -s 8 -c 4 -f -o gpurun_out/r02_prof_attn \
      python tools/gpu_bench_attn.py > gpurun_out/ncu_attn.log 2>&1
  echo "ncu attn rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'gn_act' -s 4 -c 3 -f -o gpurun_out/r02_prof_gn \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --profiler-range > gpurun_out/ncu_gn.log 2>&1
  echo "ncu gn rc=$?"
fi
