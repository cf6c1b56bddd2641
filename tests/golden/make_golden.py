#!/usr/bin/env python3
"""Regenerates tests/golden/ from the reference checkout (run in the build
container only; /root/reference does not exist on the GPU box).

Copies the reference's compressed fixtures (tests/data/*.gz, fixed.z,
empty.gzip -- the golden INPUTS of tests/test.nim:41-60,
tests/test_known_bad.nim:3, tests/bench.nim:5-11) and records, for each, the
sha256 / length / crc32 / adler32 of the expected OUTPUT taken from the
reference's own .gold / original file (never from our code).
tor-list.gz (7.3 MB, commented out of tests/test.nim:11) is not copied; it is
checked against the oracle only when /root/reference is present.
"""
import hashlib
import json
import os
import shutil
import sys
import zlib

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/tests/data"
HERE = os.path.dirname(os.path.abspath(__file__))

PAIRS = [("%s.gz" % n, "%s.gold" % n) for n in
         ("randtest1", "randtest2", "randtest3", "rfctest1", "rfctest2", "rfctest3",
          "zerotest1", "zerotest2", "zerotest3")]
PAIRS += [("empty.gz", "empty.gold"), ("empty.gzip", "empty.gold"),
          ("gzipfiletest.txt.gz", "gzipfiletest.txt"), ("fixed.z", "urls.10K")]
PAIRS += [(n + ".gz", n) for n in
          ("alice29.txt", "asyoulik.txt", "fireworks.jpg", "geo.protodata", "html", "html_x_4",
           "kppkn.gtb", "lcet10.txt", "paper-100k.pdf", "plrabn12.txt", "urls.10K")]

manifest = {}
for comp, gold in PAIRS:
    shutil.copyfile(os.path.join(REF, comp), os.path.join(HERE, comp))
    g = open(os.path.join(REF, gold), "rb").read()
    manifest[comp] = {
        "gold": gold, "len": len(g), "sha256": hashlib.sha256(g).hexdigest(),
        "crc32": zlib.crc32(g), "adler32": zlib.adler32(g),
        "ref_test": "tests/test.nim:41-60 / tests/bench.nim:5-11",
    }
# tests/test_known_bad.nim:3 pins only the output length (574); the bytes are whatever zlib says.
shutil.copyfile(os.path.join(REF, "known_bad_nitter.json.gz"), os.path.join(HERE, "known_bad_nitter.json.gz"))
g = zlib.decompress(open(os.path.join(REF, "known_bad_nitter.json.gz"), "rb").read(), 31)
assert len(g) == 574
manifest["known_bad_nitter.json.gz"] = {
    "gold": None, "len": 574, "sha256": hashlib.sha256(g).hexdigest(),
    "crc32": zlib.crc32(g), "adler32": zlib.adler32(g), "ref_test": "tests/test_known_bad.nim:3 (len only)"}
json.dump(manifest, open(os.path.join(HERE, "manifest.json"), "w"), indent=1, sort_keys=True)
print("wrote", len(manifest), "fixtures")
