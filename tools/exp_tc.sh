#!/bin/bash
# tools/exp_tc.sh : STFT-kernel experiment run on one B200: parity tests, C2 stage times under the FMCW_TC_DEBUG bits, per-warp phase times
B="python bench.py --workload c2 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-extra"
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for d in ${DBG_LIST:-0 1 8}; do
  FMCW_TC_DEBUG=$d timeout 120 $B 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('dbg $d', {k: round(v,4) for k,v in d['roofline']['stage_ms'].items()})"
done
FMCW_TC_PROF=1 timeout 120 python bench.py --workload c2 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extra 2>/dev/null | grep PROF | sort | tail -16
