#!/bin/bash
tag=${1:-run}
mkdir -p gpurun_out
python tools/diag_determinism.py 65536 2>&1 | tail -8
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$tag.log
grep -E "passed|failed|FAILED|Error|assert " gpurun_out/pytest_$tag.log | head -30
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
tail -c 800 gpurun_out/bench_$tag.err
TAG=$tag python - <<'PY'
import json,os
f="gpurun_out/bench_%s.json"%os.environ["TAG"]
try:
    d=json.load(open(f))
    print(f, "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "pageable", (d.get("e2e_pageable") or {}).get("value"), "kern", {k:round(v,1) for k,v in d["roofline"]["kernel_ms_all"].items()})
    u=d["uncompress"]; print("  unc dev out", round(u["out_gibs"],1), "e2e", u.get("e2e"))
    print("  cpu", d.get("cpu_baselines",{}).get("checksums_1_thread"))
    for k,v in d.get("extras",{}).items():
        if isinstance(v,dict): print("  ",k, json.dumps({kk:(round(vv,2) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ("value","unit","error","compress_ms","uncompress_ms","size_vs_reference","h2d_gbs","d2h_gbs","duplex_each_gbs","in_gibs","out_gibs","own_1GiB_member","foreign_zlib6_64MiB_member")})[:900], "e2e", (v.get("e2e") or {}).get("value"), (v.get("e2e") or {}).get("out_gibs"))
except Exception as ex:
    print("unreadable", ex)
PY
python tools/bench_extra.py --what crc 2>&1 | tail -5
for gb in 268435456 536870912 1073741824; do ZB200_UNC_GROUP_BYTES=$gb python bench.py --steps 3 --warmup 3 --no-extras --no-cpu 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('unc group bytes', $gb, d['uncompress']['e2e'])"; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_lz2 -s 2 -c 1 -o gpurun_out/prof_lz2_$tag -f python bench.py --workload c4 --tiles 2048 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_lz2_$tag.log 2>&1
tail -2 gpurun_out/ncu_lz2_$tag.log
