#!/bin/bash
# GPU visit: parity tests, c2 + c4 lines, ncu captures of the compress kernels.
tag=${1:-run}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; rc=$?; echo "pytest rc=$rc" >> gpurun_out/pytest_$tag.log
tail -15 gpurun_out/pytest_$tag.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-extras --no-cpu > gpurun_out/bench_c2_$tag.json 2> gpurun_out/bench_c2_$tag.err; echo "c2 rc=$?"
tail -c 600 gpurun_out/bench_c2_$tag.err; head -c 1800 gpurun_out/bench_c2_$tag.json; echo
timeout 600 python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_c4_$tag.json 2> gpurun_out/bench_c4_$tag.err; echo "c4 rc=$?"
tail -c 600 gpurun_out/bench_c4_$tag.err; head -c 1500 gpurun_out/bench_c4_$tag.json; echo
if [ "$2" != "noncu" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_lz2 -s 2 -c 1 -o gpurun_out/prof_lz2_$tag -f python bench.py --workload c4 --tiles 2048 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_lz2_$tag.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_pack|k_lz" -s 6 -c 2 -o gpurun_out/prof_lzpack_$tag -f python bench.py --blocks 16384 --steps 1 --warmup 3 --no-e2e --no-cpu --no-extras > gpurun_out/ncu_lzpack_$tag.log 2>&1
tail -3 gpurun_out/ncu_lz2_$tag.log gpurun_out/ncu_lzpack_$tag.log
fi
