#!/bin/bash
# usage: tools/run_variants.sh name...   (variants/<name>.so, built locally with -D flags)
for v in "$@"; do
  cp variants/$v.so zippy_b200/libzippy_b200.so
  c2=$(timeout 300 python bench.py --steps 2 --warmup 3 --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c2_ms=%.1f value=%.1f' % (d['uncompress']['ms'], d['value']))")
  c3=$(timeout 300 python tools/bench_extra.py --what c3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c3_inflate_ms=%.1f out_gibs=%.1f' % (d['inflate_ms'], d['out_gibs']))")
  echo "$v: $c2 $c3"
done
