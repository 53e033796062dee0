for b in 64 1 2 16; do
  echo "== batch $b"; python bench.py --steps 20 --warmup 3 --no-cpu-baseline --batch-per-gpu $b 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks'])"
done
