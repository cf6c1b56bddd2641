#!/usr/bin/env python3
"""Diagnostic: is level-1 output identical run to run at full size?  If not: which chunks, how often,
and does a single differing chunk vary when compressed alone?"""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import zippy_b200 as z
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
class E: pass
e = E(); e.torch = torch; e.dev = torch.device("cuda", 0)
d_src, T = bench.gen_c2(e, n, 0)
ctx = z.Context(0)
offs = np.arange(n + 1, dtype=np.uint64) * 65536
cap = n * (65536 + 64) + 4096
outs = []
for r in range(4):
    d = torch.full((cap,), r * 17 + 1, dtype=torch.uint8, device="cuda")
    oo = ctx.compress_batch_device(d_src.data_ptr(), offs, 1, z.dfGzip, d.data_ptr(), cap)
    outs.append((oo.copy(), d))
base = np.diff(outs[0][0].astype(np.int64))
for r in range(1, 4):
    sz = np.diff(outs[r][0].astype(np.int64))
    diff = np.nonzero(sz != base)[0]
    print("run", r, "members with a different size:", len(diff), diff[:10], (sz - base)[diff[:10]])
    if len(diff) == 0:
        same = torch.equal(outs[0][1][:int(outs[0][0][n])], outs[r][1][:int(outs[0][0][n])])
        print("   bytes identical:", same)
sz1 = np.diff(outs[1][0].astype(np.int64))
diff = np.nonzero(sz1 != base)[0]
if len(diff):
    i = int(diff[0])
    one = d_src[i * 65536:(i + 1) * 65536].contiguous()
    o1 = np.array([0, 65536], dtype=np.uint64)
    sizes = []
    d = torch.empty(70000, dtype=torch.uint8, device="cuda")
    for r in range(20):
        oo = ctx.compress_batch_device(one.data_ptr(), o1, 1, z.dfGzip, d.data_ptr(), 69996)
        sizes.append(int(oo[1]))
    print("chunk", i, "alone, 20 runs:", sorted(set(sizes)), "in batch:", int(base[i]), int(sz1[i]))
    # which sub-chunk? compare per-window masks is not exposed; compare the raw streams' first difference
    a = outs[0][1][int(outs[0][0][i]):int(outs[0][0][i + 1])].cpu().numpy()
    b = outs[1][1][int(outs[1][0][i]):int(outs[1][0][i + 1])].cpu().numpy()
    m = min(len(a), len(b))
    fd = int(np.nonzero(a[:m] != b[:m])[0][0]) if (a[:m] != b[:m]).any() else m
    print("first differing byte of the member at", fd, "of", len(a), len(b))
    import zlib
    print("both inflate to the input:", zlib.decompress(a.tobytes(), 31) == zlib.decompress(b.tobytes(), 31))
