"""CPU-only: pins the oracle (oracle/zippy_oracle.c) against the reference's own
fixtures and against system zlib.  Mirrors tests/test.nim, tests/test_levels.nim,
tests/fuzz.nim, tests/stress.nim, tests/stress2.nim of the reference."""
import random
import zlib

import pytest

from oracle import oracle as o
from tests import util

ALL_LEVELS = list(range(-2, 10))


def test_inflate_kats(golden):
    # tests/test.nim:41-60, tests/test_known_bad.nim:3, tests/bench.nim:5-11
    for name, (comp, meta) in golden.items():
        out = o.uncompress(comp)
        assert len(out) == meta["len"], name
        assert util.sha(out) == meta["sha256"], name


def test_checksum_kats(golden, corpus):
    # trailers of every fixture pin crc32/adler32
    for name, (comp, meta) in golden.items():
        raw = corpus[meta["gold"] or name]
        assert o.crc32(raw) == meta["crc32"] == zlib.crc32(raw)
        assert o.adler32(raw) == meta["adler32"] == zlib.adler32(raw)
    assert o.crc32(b"") == 0 and o.adler32(b"") == 1
    rng = random.Random(5)
    for n in (1, 7, 8, 9, 63, 64, 65, 5551, 5552, 5553, 100003):
        x = bytes(rng.randrange(256) for _ in range(n))
        assert o.crc32(x) == zlib.crc32(x) and o.adler32(x) == zlib.adler32(x)


@pytest.mark.parametrize("fmt", [o.dfDeflate, o.dfZlib, o.dfGzip])
def test_roundtrip_formats(corpus, fmt):
    # tests/test.nim:62-85 (default level, all formats, detect path for zlib/gzip)
    wb = {o.dfDeflate: -15, o.dfZlib: 15, o.dfGzip: 31}[fmt]
    for name, raw in corpus.items():
        c = o.compress(raw, o.DefaultCompression, fmt)
        assert o.uncompress(c, o.dfDeflate if fmt == o.dfDeflate else o.dfDetect) == raw, name
        assert zlib.decompress(c, wb) == raw, name  # third-party inflater (tests/stress.nim:50)


@pytest.mark.parametrize("level", ALL_LEVELS)
def test_roundtrip_levels(corpus, level):
    # tests/test_levels.nim:18-25
    for name in ("randtest1.gold", "rfctest1.gold", "zerotest1.gold", "empty.gold", "alice29.txt",
                 "asyoulik.txt", "fireworks.jpg", "geo.protodata", "html", "kppkn.gtb", "paper-100k.pdf"):
        raw = corpus[name]
        c = o.compress(raw, level)
        assert o.uncompress(c) == raw
        assert zlib.decompress(c, 31) == raw


def test_roundtrip_edges():
    for x in util.edge_inputs():
        for level in (-2, -1, 0, 1, 9):
            c = o.deflate(x, level)
            assert o.inflate(c) == x
            assert zlib.decompress(c, -15) == x


def test_gzip_fname_lengths(corpus):
    # zippy.nim:28-42: FNAME of 0..25 letters + NUL; the reader skips it (gzip.nim:49-50)
    raw = corpus["html"]
    for k in (0, 1, 25):
        c = o.compress(raw, 1, o.dfGzip, fname_len=k)
        assert c[3] == 8 and c[10:10 + k + 1] == bytes(range(97, 97 + k)) + b"\0"
        assert o.uncompress(c) == raw and zlib.decompress(c, 31) == raw


def test_invalid_args():
    with pytest.raises(o.ZippyError):
        o.deflate(b"x", 10)
    with pytest.raises(o.ZippyError):
        o.deflate(b"x", -3)
    with pytest.raises(o.ZippyError):
        o.compress(b"x", 1, o.dfDetect)
    with pytest.raises(o.ZippyError):
        o.uncompress(b"not compressed data at all....")


def test_oracle_inflates_zlib_streams(corpus):
    # tests/stress2.nim:8-20: zlib-compressed tilings of rfctest3.gold
    base = corpus["rfctest3.gold"]
    for reps in (1, 2, 5, 17):
        data = base * reps
        for lvl in (1, 6, 9):
            assert o.uncompress(zlib.compress(data, lvl)) == data
        co = zlib.compressobj(6, zlib.DEFLATED, 31)
        assert o.uncompress(co.compress(data) + co.flush()) == data
        co = zlib.compressobj(6, zlib.DEFLATED, -15, 9, zlib.Z_FIXED)
        assert o.inflate(co.compress(data) + co.flush()) == data
        co = zlib.compressobj(6, zlib.DEFLATED, -15, 9, zlib.Z_HUFFMAN_ONLY)
        assert o.inflate(co.compress(data) + co.flush()) == data
        assert o.inflate(zlib.compress(data, 0)[2:-4]) == data


def test_fuzz_never_crashes(golden):
    # tests/fuzz.nim:16-33: flip one byte, then truncate; only ZippyError may escape
    rng = random.Random(77)
    names = ["randtest1.gz", "randtest2.gz", "randtest3.gz", "rfctest1.gz", "rfctest2.gz", "rfctest3.gz",
             "zerotest1.gz", "zerotest2.gz"]
    ok = bad = 0
    for _ in range(3000):
        comp = bytearray(golden[rng.choice(names)][0])
        pos = rng.randrange(len(comp))
        comp[pos] = rng.randrange(256)
        for data in (bytes(comp), bytes(comp[:pos])):
            try:
                assert len(o.uncompress(data)) > 0
                ok += 1
            except o.ZippyError:
                bad += 1
    assert ok > 0 and bad > 0


def test_stress_run_length_blobs():
    # tests/stress.nim:10-58
    rng = random.Random(99)
    for _ in range(150):
        x = util.run_length_blob(rng)
        c = o.compress(x, o.DefaultCompression, o.dfZlib)
        assert o.uncompress(c) == x and zlib.decompress(c) == x
        y = bytearray(x)
        rng.shuffle(y)
        y = bytes(y)
        c = o.compress(y, 1, o.dfGzip)
        assert o.uncompress(c) == y and zlib.decompress(c, 31) == y


def test_batch_threads_match_single(corpus):
    import numpy as np
    T = util.text_corpus(corpus)
    blocks = [util.c2_block(T, i) for i in range(12)]
    base = np.frombuffer(b"".join(blocks), dtype=np.uint8)
    offs = np.arange(13, dtype=np.uint64) * 65536
    t1, l1, s1 = o.compress_batch(base, offs, 1, o.dfGzip, threads=1)
    t4, l4, s4 = o.compress_batch(base, offs, 1, o.dfGzip, threads=4)
    assert t1 == t4 and (l1 == l4).all() and not s1.any() and not s4.any()
    assert int(l1[0]) == len(o.compress(blocks[0], 1, o.dfGzip))


def test_simd_checksums_match_scalar_and_zlib():
    """The CPU arm runs what the reference runs on amd64 (PCLMUL CRC-32 crc32_simd.nim:39-144, SSSE3
    Adler-32 adler32_simd.nim:45-120); both are checked against the scalar restatement and zlib."""
    import random
    from oracle import oracle as o
    L = o.lib()
    rng = random.Random(7)
    blob = bytes(rng.randrange(256) for _ in range(70001))
    for n in list(range(0, 200)) + [1000, 5551, 5552, 5553, 65535, 65536, 70001]:
        b = blob[:n]
        assert o.crc32(b) == zlib.crc32(b) == L.zo_crc32_scalar(b, n), n
        assert o.adler32(b) == zlib.adler32(b) == L.zo_adler32_scalar(b, n), n
    big = blob * 40
    assert o.crc32(big) == zlib.crc32(big) and o.adler32(big) == zlib.adler32(big)
