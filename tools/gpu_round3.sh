#!/bin/bash
tag=${1:-run}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$tag.log
tail -12 gpurun_out/pytest_$tag.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
tail -c 800 gpurun_out/bench_$tag.err
ZB200_SERIAL_COPIES=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-extras --no-cpu > gpurun_out/bench_serial_$tag.json 2> gpurun_out/bench_serial_$tag.err; echo "serial rc=$?"
TAG=$tag python - <<'PY'
import json
for f in ("gpurun_out/bench_TAG.json","gpurun_out/bench_serial_TAG.json"):
    try:
        d=json.load(open(f.replace("TAG", __import__("os").environ["TAG"])))
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f, "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "kern", {k:round(v,1) for k,v in d["roofline"]["kernel_ms_all"].items()})
    u=d["uncompress"]; print("  unc dev out", round(u["out_gibs"],1), "e2e", u.get("e2e"))
    for k,v in d.get("extras",{}).items():
        if isinstance(v,dict): print("  ",k, {kk:(round(vv,2) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ("value","unit","error","compress_ms","uncompress_ms","size_vs_reference","h2d_gbs","d2h_gbs","duplex_each_gbs","in_gibs","out_gibs")}, "e2e", (v.get("e2e") or {}).get("value"), (v.get("e2e") or {}).get("out_gibs"))
PY
