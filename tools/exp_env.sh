for v in "B200_FUSE_GN_MAX_PIX=0" "B200_FUSE_GN_MAX_PIX=4096" "B200_FUSE_GN_MAX_PIX=16384" "B200_FUSE_GN_MAX_PIX=65536" "B200_FUSE_GN_MAX_PIX=0"; do
  echo "== $v"; env $v python bench.py --steps 30 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['kernels_per_step'], d['clocks'])"
done
