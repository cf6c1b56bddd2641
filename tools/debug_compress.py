"""GPU-side debugging aid: compress inputs and report where zlib stops agreeing."""
import sys, zlib
sys.path.insert(0, ".")
import zippy_b200 as z
from tests import util

corpus = util.load_corpus()
T = util.text_corpus(corpus)
cases = {"rfctest1": corpus["rfctest1.gold"], "alice29": corpus["alice29.txt"], "alice64k": corpus["alice29.txt"][:65536],
         "alice64k+1": corpus["alice29.txt"][:65537], "alice100k": corpus["alice29.txt"][:100000],
         "alice8k": corpus["alice29.txt"][:8192], "alice20k": corpus["alice29.txt"][:20000],
         "c2_0": util.c2_block(T, 0), "zeros1M": b"\0" * (1 << 20), "html_x_4": corpus["html_x_4"],
         "urls": corpus["urls.10K"], "kppkn": corpus["kppkn.gtb"], "geo": corpus["geo.protodata"]}
for name, raw in cases.items():
    for level in (1,):
        d = z.deflate(raw, level)
        do = zlib.decompressobj(-15)
        out = b""
        err = None
        try:
            out = do.decompress(d)
        except zlib.error as e:
            err = str(e)
            # feed byte by byte to find how far it gets
            do = zlib.decompressobj(-15)
            out = b""
            for i in range(len(d)):
                try:
                    out += do.decompress(d[i:i + 1])
                except zlib.error:
                    break
        mism = next((i for i in range(min(len(out), len(raw))) if out[i] != raw[i]), None)
        print("%-12s len=%7d comp=%7d ratio=%.3f decoded=%7d firstmismatch=%s err=%s" %
              (name, len(raw), len(d), len(d) / max(1, len(raw)), len(out), mism, err))

print("---- default level vs oracle")
from oracle import oracle as o
for name in ("urls", "alice29", "html_x_4", "kppkn", "geo", "zeros1M"):
    raw = cases[name]
    d = z.deflate(raw, -1)
    ok = zlib.decompress(d, -15) == raw
    ref = len(o.deflate(raw, -1))
    print("%-10s gpu=%8d oracle(-1)=%8d ratio_to_ref=%.4f ok=%s lvl1=%d" % (name, len(d), ref, len(d) / ref, ok, len(z.deflate(raw, 1))))
