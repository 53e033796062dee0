#!/bin/bash
# round 2, last visit: the bench lines with the per-kernel timing behind a spin kernel (no host launch latency in the interval),
# clip / rollout at their default 50 steps
set -u
mkdir -p gpurun_out
export B200_TUNE_FILE=gpurun_out/r02b_tune.json
timeout 600 python bench.py > gpurun_out/r02b_bench_default.json 2> gpurun_out/r02b_bench_default.err
echo "bench default rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02b_bench_default.json')); print('ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], d['clocks']); r=d['roofline']; print(r['kernel'], r['frac'], r['per_launch']['ms']); print(r.get('resblock'))"
timeout 600 python bench.py --steps 20 --warmup 3 --profile-ops --no-cpu-baseline > gpurun_out/r02b_bench_ops.json 2> gpurun_out/r02b_bench_per_kernel.txt
echo "bench ops rc=$?"; tail -9 gpurun_out/r02b_bench_per_kernel.txt
timeout 600 python bench.py --workload clip > gpurun_out/r02b_bench_clip.json 2> gpurun_out/r02b_bench_clip.err
echo "clip rc=$?"; head -c 300 gpurun_out/r02b_bench_clip.json; echo
timeout 900 python bench.py --workload rollout > gpurun_out/r02b_bench_rollout.json 2> gpurun_out/r02b_bench_rollout.err
echo "rollout rc=$?"; head -c 300 gpurun_out/r02b_bench_rollout.json; echo
