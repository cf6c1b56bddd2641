#!/bin/bash
# debug visit: the C++ zip test with and without the gated uncompress
set -x
cd "$(dirname "$0")/.."
g++ -std=c++17 -O1 -g -o /tmp/cpp_zip_test tests/native/cpp_zip_test.cpp -Lzippy_b200 -l:libzippy_b200.so -Wl,-rpath,$PWD/zippy_b200
python - <<'PY'
import sys
sys.path.insert(0, ".")
from tests import test_ziparchives as t
t._zipfile_made_archive("/tmp/in.zip")
PY
for g in 0 1; do
  echo "== gated=$g"
  ( time ZB200_UNC_GATED=$g ZB200_DEBUG_SEGV=1 timeout 120 /tmp/cpp_zip_test /tmp/in.zip /tmp/out$g.zip ) 2>&1 | tail -8
done
