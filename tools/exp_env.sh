timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout 300 -p no:cacheprovider -k "in_conv or out_conv" 2>&1 | tail -2
timeout 600 python -m pytest tests/test_gpu_unet.py tests/test_gpu_layout.py -m gpu -q -x --timeout 300 -p no:cacheprovider 2>&1 | tail -2
python bench.py --steps 30 --warmup 3 --no-cpu-baseline --profile-ops 2>/tmp/err.txt | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks'])"; grep "fir_resample\|in_conv \|out_conv " /tmp/err.txt
