#!/bin/bash
# quick GPU check: a subset of the GPU tests (PYTEST_ARGS) + one bench line
mkdir -p gpurun_out
timeout 900 python -m pytest ${PYTEST_ARGS:-tests/test_gpu_unet.py} -m gpu -q --timeout 600 -p no:cacheprovider -x > gpurun_out/pytest_quick.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_quick.log
timeout 600 python bench.py --steps 30 --warmup 3 ${BENCH_ARGS:---no-cpu-baseline} > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_quick.json')); print('ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], d['clocks'])"
tail -12 gpurun_out/bench_quick.err
