#!/bin/bash
# tools/scale_run.sh N : on N GPUs of one box: weak-scaling C3 bench, strong-scaling point, C4 stream, C5 fleet, multi-GPU check
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29521 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/scale_c3_weak_n$N.json 2> gpurun_out/scale_n$N.err
$TR --master-port 29522 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --scaling strong --frames 200000 > gpurun_out/scale_c3_strong_n$N.json 2>> gpurun_out/scale_n$N.err
$TR --master-port 29523 bench.py --gpus $N --workload c2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/scale_c2_weak_n$N.json 2>> gpurun_out/scale_n$N.err
$TR --master-port 29524 profiles/c4_stream.py > gpurun_out/c4_n$N.json 2>> gpurun_out/scale_n$N.err
$TR --master-port 29525 profiles/c5_fleet.py > gpurun_out/c5_n$N.json 2>> gpurun_out/scale_n$N.err
$TR --master-port 29526 tests/run_multi_gpu_check.py > gpurun_out/multi_gpu_check_n$N.log 2>&1
$TR --master-port 29527 tools/pcie_probe.py > gpurun_out/pcie_n$N.json 2>> gpurun_out/scale_n$N.err
free -g | head -2 > gpurun_out/host_n$N.txt; lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)" >> gpurun_out/host_n$N.txt; nvidia-smi topo -m >> gpurun_out/host_n$N.txt 2>&1
tail -c 600 gpurun_out/scale_n$N.err
