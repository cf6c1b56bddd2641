#!/bin/bash
# Profiles for the round: launch list of the default bench line, full captures of the dominant kernels.
tag=${1:-r2}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-extras > gpurun_out/bench_under_ncu_$tag.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches_$tag.csv
# compress kernels at the bench's own launch size (65536 chunks): one launch each
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_lz|k_pack|k_huff" -s 6 -c 3 -o gpurun_out/prof_compress_$tag -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-extras > gpurun_out/ncu_compress_$tag.log 2>&1; tail -1 gpurun_out/ncu_compress_$tag.log
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_inflate|k_piece_checksum" -c 2 -o gpurun_out/prof_inflate_$tag -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-extras > gpurun_out/ncu_inflate_$tag.log 2>&1; tail -1 gpurun_out/ncu_inflate_$tag.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_lz2 -s 2 -c 1 -o gpurun_out/prof_lz2_$tag -f python bench.py --workload c4 --tiles 2048 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_lz2_$tag.log 2>&1; tail -1 gpurun_out/ncu_lz2_$tag.log
# standalone checksums (tools/bench_extra.py --what crc: five crc32 launches, then five adler32 launches)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_piece_checksum -s 4 -c 1 -o gpurun_out/prof_crc_$tag -f python tools/bench_extra.py --what crc > gpurun_out/ncu_crc_$tag.log 2>&1; tail -1 gpurun_out/ncu_crc_$tag.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_piece_checksum -s 9 -c 1 -o gpurun_out/prof_adler_$tag -f python tools/bench_extra.py --what crc > gpurun_out/ncu_adler_$tag.log 2>&1; tail -1 gpurun_out/ncu_adler_$tag.log
# the cubins the reports' SASS belongs to (tools/sass_lines.py)
mkdir -p gpurun_out/cubin_$tag && (cd gpurun_out/cubin_$tag && cuobjdump -xelf all ../../zippy_b200/libzippy_b200.so > /dev/null 2>&1; ls | head)
