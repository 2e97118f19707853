#!/bin/bash
# tools/ab.sh [tag] : build, run GPU tests + C2/C3 device-resident benches on one B200, print the stage summary
cd /root/repo
tag=${1:-x}
python -m fmcw_radar_processing_b200.build 2>&1 | grep -v deprecated | tail -1
tools/gpu.sh --timeout 900 -- "python -m pytest tests -m gpu -x -q 2>&1 | tail -3; python bench.py --workload c2 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b_c2_$tag.json 2> gpurun_out/b_$tag.err; python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b_c3_$tag.json 2>> gpurun_out/b_$tag.err; tail -c 300 gpurun_out/b_$tag.err" > gpurun_out/call_$tag.log 2>&1
tail -5 gpurun_out/call_$tag.log
python - <<PY
import json
for f in ("gpurun_out/b_c2_$tag.json","gpurun_out/b_c3_$tag.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        st=d["roofline"]["stage_ms"]
        print(f, "%.3fM f/s" % (d["value"]/1e6), "ms %.4f" % d["ms_per_step"], {k: round(v,4) for k,v in st.items()}, "chain_frac %.3f" % d["roofline"]["chain_frac_of_peak"], "stft_frac %.3f" % d["roofline"]["frac"])
    except Exception as e: print(f, "ERR", e)
PY
