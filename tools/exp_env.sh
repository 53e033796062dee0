for v in "B200_PDL=2" "B200_PDL=0" "B200_PDL=2" "B200_PDL=0"; do
  echo "== $v"; env $v python bench.py --steps 30 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks'])"
done
timeout 900 python -m pytest tests/test_gpu_unet.py tests/test_gpu_kernels.py tests/test_gpu_layout.py -m gpu -q -x --timeout 600 -p no:cacheprovider 2>&1 | tail -3
