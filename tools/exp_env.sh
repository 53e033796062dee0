timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout 300 -p no:cacheprovider -k "out_conv or in_conv" 2>&1 | tail -3
for v in "B200_OUT_CONV=gather" "B200_OUT_CONV=rows" "B200_OUT_CONV=gather"; do
  echo "== $v"; env $v python bench.py --steps 30 --warmup 3 --no-cpu-baseline --profile-ops 2>/tmp/err.txt | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks'])"; grep "out_conv\|in_conv " /tmp/err.txt
done
