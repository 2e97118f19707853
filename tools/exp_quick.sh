B="python bench.py --workload c2 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extra"
for d in 0 2048 68 76; do
  echo "== dbg $d"
  FMCW_TC_PROF=1 FMCW_TC_DEBUG=$d timeout 120 $B 2>&1 | grep "PROF cta  77" | sort | uniq | awk '{k=$5 $6 $7; if (!(k in seen)) {seen[k]=1; print}}' | grep -v "lane [123]"
done
