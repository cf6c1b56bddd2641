#!/bin/bash
# One GPU-box visit: parity tests, the default bench line (all five configs), launch list.
# Usage (under gpurun): bash tools/gpu_round.sh <tag> [steps]
tag=${1:-run}; steps=${2:-5}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$tag.txt 2>&1
nproc > gpurun_out/host_$tag.txt; lscpu | grep -E "Model name|Socket|NUMA" >> gpurun_out/host_$tag.txt; df -h /dev/shm >> gpurun_out/host_$tag.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$tag.log
tail -5 gpurun_out/pytest_$tag.log
timeout 900 python bench.py --steps $steps --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/bench_$tag.err
head -c 3000 gpurun_out/bench_$tag.json
