#!/bin/bash
# A/B: conv micro-bench (ablation build) then the conv parity tests + bench with whatever .so is in tree
timeout 300 python tools/exp_conv_ablate.py
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout 300 -p no:cacheprovider -k conv 2>&1 | tail -3
timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], d['clocks'])"
