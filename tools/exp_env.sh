for v in "B200_PDL=1" "B200_FUSE_GN=1" "B200_PDL=1 B200_FUSE_GN=1" "X=1"; do
  echo "== $v"; env $v python bench.py --steps 30 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks'])"
done
