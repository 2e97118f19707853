#!/bin/bash
# tools/final_bench.sh : the three bench lines of the final build on one B200
python bench.py > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err
python bench.py --workload c2 > gpurun_out/bench_r2_c2.json 2>> gpurun_out/bench_r2.err
tail -c 300 gpurun_out/bench_r2.err
