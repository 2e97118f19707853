#!/bin/bash
# tools/final_evidence.sh : one B200: bench lines, precision probe, ncu launch list and full captures of the two top kernels
python bench.py > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err
python bench.py --workload c2 > gpurun_out/bench_r2_c2.json 2>> gpurun_out/bench_r2.err
python bench.py --impl reference > gpurun_out/bench_reference_r2.json 2>> gpurun_out/bench_r2.err
python tests/precision_probe.py > gpurun_out/precision_r2.txt 2>&1
CMD="python bench.py --workload c2 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extra"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv $CMD > gpurun_out/ncu1.log 2>&1
$CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"stft_tc_kernel|frame_chain_warp" -s 2 -c 2 -o gpurun_out/prof_r2_final -f $CMD > gpurun_out/ncu2.log 2>&1
tail -c 300 gpurun_out/bench_r2.err; tail -2 gpurun_out/ncu2.log
