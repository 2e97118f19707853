for m in 3 4; do
  FMCW_CHAIN_MINB=$m python bench.py --workload c2 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-extra 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c2 minb $m', {k: round(v,4) for k,v in d['roofline']['stage_ms'].items()})"
  FMCW_CHAIN_MINB=$m python bench.py --workload c3 --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --no-extra 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c3 minb $m', round(d['ms_per_step'],3), {k: round(v,4) for k,v in d['roofline']['stage_ms'].items()})"
done
