#!/bin/bash
# round 2, session 3: GPU test-suite + default bench + per-kernel profile after the column-walk conv
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -x > gpurun_out/r02b_pytest_gpu.txt 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r02b_pytest_gpu.txt
B200_TUNE_FILE=gpurun_out/r02b_tune.json timeout 600 python bench.py --steps 30 --warmup 3 --profile-ops --no-cpu-baseline > gpurun_out/r02b_bench_ops.json 2> gpurun_out/r02b_bench_per_kernel.txt
echo "bench ops rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02b_bench_ops.json')); print('ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], d['clocks']); r=d['roofline']; print(r['kernel'], r['frac'], r['per_launch']['ms']); print(r.get('resblock'))"
tail -12 gpurun_out/r02b_bench_per_kernel.txt
