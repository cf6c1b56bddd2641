"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI
(zippy_b200 -> libzippy_b200.so), against the CPU oracle, the reference's golden fixtures
and system zlib.  Mirrors tests/test.nim, test_levels.nim, test_known_bad.nim, fuzz.nim,
stress.nim and stress2.nim of the reference.  Integer/byte work: every comparison is
bit-exact."""
import os
import random
import zlib

import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu

ALL_LEVELS = list(range(-2, 10))


@pytest.fixture(scope="module")
def z():
    import zippy_b200
    return zippy_b200


@pytest.fixture(scope="module")
def o():
    from oracle import oracle
    return oracle


# ------------------------------------------------------------------ inflate (pinned)
def test_inflate_golden_fixtures(z, golden):
    # tests/test.nim:41-60, tests/test_known_bad.nim:3, tests/bench.nim:5-11
    names = list(golden)
    outs = z.uncompress_batch([golden[n][0] for n in names])
    for n, out in zip(names, outs):
        meta = golden[n][1]
        assert not isinstance(out, Exception), (n, out)
        assert len(out) == meta["len"], n
        assert util.sha(out) == meta["sha256"], n
    for n in ("alice29.txt.gz", "fixed.z", "empty.gz", "zerotest3.gz", "gzipfiletest.txt.gz"):
        assert util.sha(z.uncompress(golden[n][0])) == golden[n][1]["sha256"]


def test_inflate_matches_oracle_on_oracle_streams(z, o, corpus):
    items, want = [], []
    for name, raw in corpus.items():
        for level in (-2, -1, 0, 1, 9):
            for fmt in (o.dfGzip, o.dfZlib):
                items.append(o.compress(raw, level, fmt, fname_len=level % 26))
                want.append(raw)
    for x in util.edge_inputs():
        items.append(o.compress(x, 1, o.dfGzip))
        want.append(x)
        items.append(o.compress(x, -1, o.dfZlib))
        want.append(x)
    outs = z.uncompress_batch(items)
    for i, (out, w) in enumerate(zip(outs, want)):
        assert not isinstance(out, Exception), (i, out)
        assert out == w, i
    raws = [o.deflate(x, 6) for x in util.edge_inputs()[:20]]
    outs = z.uncompress_batch(raws, z.dfDeflate)
    for out, x in zip(outs, util.edge_inputs()[:20]):
        assert out == x


def test_inflate_zlib_streams(z, corpus):
    # tests/stress2.nim:8-20 + block types zlib can be forced into
    base = corpus["rfctest3.gold"]
    items, want = [], []
    for reps in (1, 2, 5, 17, 40):
        data = base * reps
        for lvl in (0, 1, 6, 9):
            items.append(zlib.compress(data, lvl))
            want.append(data)
        co = zlib.compressobj(6, zlib.DEFLATED, 31)
        items.append(co.compress(data) + co.flush())
        want.append(data)
        for strat in (zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FILTERED):
            co = zlib.compressobj(6, zlib.DEFLATED, 15, 9, strat)
            items.append(co.compress(data) + co.flush())
            want.append(data)
        co = zlib.compressobj(9, zlib.DEFLATED, 15, 9)
        parts = b"".join(co.compress(data[i:i + 5000]) + co.flush(zlib.Z_FULL_FLUSH) for i in range(0, len(data), 5000))
        items.append(parts + co.flush())
        want.append(data)
    outs = z.uncompress_batch(items)
    for i, (out, w) in enumerate(zip(outs, want)):
        assert out == w, i


def test_inflate_error_contract(z, o, golden):
    # tests/fuzz.nim:16-33: flip a byte / truncate; each outcome must agree with the oracle:
    # either both succeed with identical bytes or both report an error.
    rng = random.Random(2024)
    names = ["randtest1.gz", "randtest2.gz", "randtest3.gz", "rfctest1.gz", "rfctest2.gz", "rfctest3.gz",
             "zerotest1.gz", "zerotest2.gz"]
    items = []
    for _ in range(1500):
        comp = bytearray(golden[rng.choice(names)][0])
        pos = rng.randrange(len(comp))
        comp[pos] = rng.randrange(256)
        items.append(bytes(comp))
        items.append(bytes(comp[:pos]))
    items += [b"", b"\x1f", b"\x1f\x8b\x08" + b"\0" * 20, b"\x78\x01", b"x" * 40, b"\x78\x9c\x03\x00\x00\x00\x00\x01"]
    outs = z.uncompress_batch(items)
    n_ok = n_bad = 0
    for data, out in zip(items, outs):
        try:
            want = o.uncompress(data)
        except o.ZippyError as e:
            assert isinstance(out, z.ZippyError), (len(data), e)
            n_bad += 1
            continue
        assert not isinstance(out, Exception), (len(data), out)
        assert out == want
        n_ok += 1
    assert n_ok > 0 and n_bad > 0


def test_inflate_wrapper_errors(z, o, corpus):
    raw = corpus["html"]
    good = o.compress(raw, 1, o.dfGzip)
    cases = {
        "bad id": bytes([0]) + good[1:],
        "bad method": good[:2] + b"\x07" + good[3:],
        "reserved flag": good[:3] + bytes([good[3] | 0x80]) + good[4:],
        "fextra": good[:3] + bytes([good[3] | 0x04]) + good[4:],
        "bad crc": good[:-8] + bytes([good[-8] ^ 1]) + good[-7:],
        "bad isize": good[:-1] + bytes([good[-1] ^ 1]),
    }
    for name, data in cases.items():
        with pytest.raises(o.ZippyError):
            o.uncompress(data, o.dfGzip)
        with pytest.raises(z.ZippyError):
            z.uncompress(data, z.dfGzip)
    zl = o.compress(raw, 1, o.dfZlib)
    with pytest.raises(z.ZippyError):
        z.uncompress(zl[:-1] + bytes([zl[-1] ^ 1]))
    with pytest.raises(z.ZippyError):
        z.uncompress(bytes([0x78, 0x20 | (31 - (0x7820 % 31))]) + zl[2:], z.dfZlib)  # FDICT


# ------------------------------------------------------------------ deflate (round trip)
@pytest.mark.parametrize("fmt_name", ["dfDeflate", "dfZlib", "dfGzip"])
def test_compress_roundtrip_formats(z, o, corpus, fmt_name):
    # tests/test.nim:62-85
    fmt = getattr(z, fmt_name)
    wb = {z.dfDeflate: -15, z.dfZlib: 15, z.dfGzip: 31}[fmt]
    names = list(corpus)
    comp = z.compress_batch([corpus[n] for n in names], z.DefaultCompression, fmt)
    for n, c in zip(names, comp):
        assert o.uncompress(c, o.dfDeflate if fmt == z.dfDeflate else o.dfDetect) == corpus[n], n
        assert zlib.decompress(c, wb) == corpus[n], n
    back = z.uncompress_batch(comp, z.dfDeflate if fmt == z.dfDeflate else z.dfDetect)
    for n, b in zip(names, back):
        assert b == corpus[n], n


@pytest.mark.parametrize("level", ALL_LEVELS)
def test_compress_roundtrip_levels(z, o, corpus, level):
    # tests/test_levels.nim:18-25
    names = ["randtest1.gold", "rfctest1.gold", "zerotest1.gold", "empty.gold", "alice29.txt", "asyoulik.txt",
             "fireworks.jpg", "geo.protodata", "html", "kppkn.gtb", "paper-100k.pdf"]
    for n in names:
        c = z.compress(corpus[n], level)
        assert c[:4] == b"\x1f\x8b\x08\x08"          # zippy.nim:23-26
        assert o.uncompress(c) == corpus[n]
        assert zlib.decompress(c, 31) == corpus[n]
        assert z.uncompress(c) == corpus[n]


def test_compress_edges(z, o):
    xs = util.edge_inputs()
    for level in (1, -2, 0):
        comp = z.compress_batch(xs, level, z.dfZlib)
        for x, c in zip(xs, comp):
            assert zlib.decompress(c) == x
            assert o.uncompress(c) == x
    with pytest.raises(z.ZippyError):
        z.compress(b"x", 10)
    with pytest.raises(z.ZippyError):
        z.compress(b"x", -3)
    with pytest.raises(z.ZippyError):
        z.compress(b"x", 1, z.dfDetect)


def test_compress_gzip_framing(z, corpus):
    # zippy.nim:21-58: FNAME flag, k letters + NUL, CRC32 + ISIZE trailer
    raw = corpus["html"]
    outs = z.compress_batch([raw] * 3, 1, z.dfGzip, fname_lens=[0, 1, 25])
    for k, c in zip((0, 1, 25), outs):
        assert c[:10] == bytes([31, 139, 8, 8, 0, 0, 0, 0, 0, 0])
        assert c[10:11 + k] == bytes(range(97, 97 + k)) + b"\0"
        assert int.from_bytes(c[-8:-4], "little") == zlib.crc32(raw)
        assert int.from_bytes(c[-4:], "little") == len(raw)
    zl = z.compress(raw, 1, z.dfZlib)
    assert zl[:2] == b"\x78\x01" and int.from_bytes(zl[-4:], "big") == zlib.adler32(raw)   # zippy.nim:60-78


def test_compress_c2_blocks_and_ratio(z, o, corpus):
    # BASELINE config 2 at reduced count: seeded 64 KiB text blocks, level 1, gzip
    T = util.text_corpus(corpus)
    blocks = [util.c2_block(T, i) for i in range(256)]
    comp = z.compress_batch(blocks, z.BestSpeed, z.dfGzip)
    for b, c in zip(blocks[:64], comp[:64]):
        assert o.uncompress(c) == b
    back = z.uncompress_batch(comp)
    assert all(r == b for r, b in zip(back, blocks))
    gpu = sum(map(len, comp))
    ref = sum(len(o.compress(b, 1, o.dfGzip)) for b in blocks[:64]) * 4
    assert gpu < 1.10 * ref, (gpu, ref)   # ratio within 10% of the reference's level 1 on text


def test_compress_large_members(z, o, corpus):
    rng = random.Random(7)
    big = [corpus["urls.10K"], corpus["plrabn12.txt"], b"\0" * (1 << 20), corpus["fireworks.jpg"] * 3,
           bytes(rng.randrange(256) for _ in range(300001)), corpus["html_x_4"] + corpus["kppkn.gtb"]]
    for fmt in (z.dfGzip, z.dfZlib, z.dfDeflate):
        comp = z.compress_batch(big, 1, fmt)
        for x, c in zip(big, comp):
            assert o.uncompress(c, o.dfDeflate if fmt == z.dfDeflate else o.dfDetect) == x
        back = z.uncompress_batch(comp, z.dfDeflate if fmt == z.dfDeflate else z.dfDetect)
        assert all(b == x for b, x in zip(back, big))


def test_stress_run_length_blobs(z, o):
    # tests/stress.nim:10-58
    rng = random.Random(4242)
    xs = []
    for _ in range(100):
        x = util.run_length_blob(rng)
        y = bytearray(x)
        rng.shuffle(y)
        xs += [x, bytes(y)]
    comp = z.compress_batch(xs, z.DefaultCompression, z.dfZlib)
    for x, c in zip(xs, comp):
        assert zlib.decompress(c) == x
    back = z.uncompress_batch(comp)
    assert all(b == x for b, x in zip(back, xs))


# ------------------------------------------------------------------ checksums + seam
def test_checksums(z, corpus):
    rng = random.Random(5)
    xs = [b"", b"a", b"abc"] + [bytes(rng.randrange(256) for _ in range(n))
                                 for n in (4, 5, 127, 128, 129, 255, 4095, 32767, 32768, 32769, 65536, 100003)]
    xs += [corpus["alice29.txt"], corpus["urls.10K"], b"\xff" * 300000]
    crcs = z.checksum_batch(xs, "crc32")
    ads = z.checksum_batch(xs, "adler32")
    for x, c, a in zip(xs, crcs, ads):
        assert int(c) == zlib.crc32(x), len(x)
        assert int(a) == zlib.adler32(x), len(x)
    assert z.crc32(corpus["html"]) == zlib.crc32(corpus["html"])
    assert z.adler32(corpus["html"]) == zlib.adler32(corpus["html"])


def test_checksums_of_large_buffers(z):
    """One very large input (more than 2048 pieces of 32 KiB) is folded by a whole CTA, next to small ones folded
    by one warp each; ragged last pieces; both checksums against zlib."""
    rng = np.random.default_rng(11)
    big1 = rng.integers(0, 256, (64 << 20) + 12345, dtype=np.uint8).tobytes()     # 2049 pieces
    big2 = rng.integers(0, 256, (150 << 20) + 7, dtype=np.uint8).tobytes()
    xs = [b"tiny", big1, b"", big2, big1[:70000]]
    crcs = z.checksum_batch(xs, "crc32")
    ads = z.checksum_batch(xs, "adler32")
    for x, c, a in zip(xs, crcs, ads):
        assert int(c) == zlib.crc32(x), len(x)
        assert int(a) == zlib.adler32(x), len(x)
    assert z.crc32(big2) == zlib.crc32(big2) and z.adler32(big2) == zlib.adler32(big2)


def test_seam_deflate_inflate(z, o, corpus):
    # deflate.nim:207 / inflate.nim:268 signatures: raw stream, inflate from byte `pos`
    for name in ("alice29.txt", "html", "empty.gold", "fireworks.jpg"):
        raw = corpus[name]
        for level in (1, -1, 0, -2):
            d = z.deflate(raw, level)
            assert zlib.decompress(d, -15) == raw
            assert o.inflate(d) == raw
            assert z.inflate(d) == raw
            assert z.inflate(b"\xaa\xbb\xcc" + d, pos=3) == raw
    with pytest.raises(z.ZippyError):
        z.inflate(b"\x07")          # BTYPE 3 (inflate.nim:289)
    with pytest.raises(z.ZippyError):
        z.inflate(b"")


def test_device_resident_batch(z, o, corpus):
    torch = pytest.importorskip("torch")
    T = util.text_corpus(corpus)
    n = 64
    blocks = [util.c2_block(T, i) for i in range(n)]
    host = np.frombuffer(b"".join(blocks), dtype=np.uint8)
    d_src = torch.from_numpy(host.copy()).cuda()
    offs = np.arange(n + 1, dtype=np.uint64) * 65536
    cap = n * 70000
    d_dst = torch.zeros(cap, dtype=torch.uint8, device="cuda")
    ctx = z.Context()
    ctx.set_stream(ctx.LEGACY_DEFAULT_STREAM)   # ordered with torch's default-stream work on the same buffers
    oo = ctx.compress_batch_device(d_src.data_ptr(), offs, 1, z.dfGzip, d_dst.data_ptr(), cap)
    comp = d_dst.cpu().numpy()
    for i in range(n):
        assert o.uncompress(comp[int(oo[i]):int(oo[i + 1])].tobytes()) == blocks[i]
    sizes, st = ctx.uncompressed_sizes_device(d_dst.data_ptr(), oo, z.dfDetect)
    assert not st.any() and (sizes == 65536).all()
    d_back = torch.empty(n * 65536, dtype=torch.uint8, device="cuda")
    lens, st = ctx.uncompress_batch_device(d_dst.data_ptr(), oo, z.dfDetect, d_back.data_ptr(), offs)
    assert not st.any() and (lens == 65536).all()
    assert torch.equal(d_back, d_src)
    t = ctx.timing()
    assert t["kernel_launches"] >= 2
    ctx.close()


def test_launch_groups_and_tight_capacity(z, o, corpus, monkeypatch):
    """Batches are processed in launch groups whose output offsets are chained on the device;
    force tiny groups, and give the device variant a destination smaller than the worst-case
    bound (the path that sizes the zero-fill from the real extent)."""
    torch = pytest.importorskip("torch")
    monkeypatch.setenv("ZB200_GROUP_CHUNKS", "3")
    ctx = z.Context()
    monkeypatch.delenv("ZB200_GROUP_CHUNKS")
    rng = random.Random(11)
    items = [corpus["alice29.txt"], b"", corpus["html"], corpus["urls.10K"], b"x" * 70000, corpus["geo.protodata"],
             bytes(rng.randrange(256) for _ in range(200000)), corpus["kppkn.gtb"][:65536], b"abc"]
    base = np.frombuffer(b"".join(items), dtype=np.uint8)
    offs = np.zeros(len(items) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(x) for x in items])
    for fmt in (z.dfGzip, z.dfZlib):
        out, oo = ctx.compress_batch(base, offs, 1, fmt, fname_lens=[i % 26 for i in range(len(items))])
        for i, x in enumerate(items):
            assert o.uncompress(out[int(oo[i]):int(oo[i + 1])].tobytes()) == x, i
    # device variant with a tight destination
    d_src = torch.from_numpy(base.copy()).cuda()
    total = int(oo[-1])
    cap = (total + 64 + 3) & ~3
    d_dst = torch.full((cap,), 0xAB, dtype=torch.uint8, device="cuda")   # dirty buffer: the call writes every output byte itself
    torch.cuda.synchronize()   # ctx runs on its own stream
    o2 = ctx.compress_batch_device(d_src.data_ptr(), offs, 1, z.dfZlib, d_dst.data_ptr(), cap)
    host = d_dst.cpu().numpy()
    for i, x in enumerate(items):
        assert zlib.decompress(host[int(o2[i]):int(o2[i + 1])].tobytes()) == x, i
    with pytest.raises(z.ZippyError):
        ctx.compress_batch_device(d_src.data_ptr(), offs, 1, z.dfZlib, d_dst.data_ptr(), 1000)
    ctx.close()


@pytest.mark.parametrize("gated", ["1", "0"])
def test_host_uncompress_member_groups(z, o, corpus, monkeypatch, gated):
    """The host-buffer uncompress runs member groups through a copy-in / inflate / copy-out
    pipeline; force tiny groups (many groups, groups of one oversized member, empty members).
    gated=1: one inflate launch walks the whole batch behind the copy-in and the copy-out stream waits
    on per-group done counts; gated=0: one launch per group."""
    monkeypatch.setenv("ZB200_UNC_GROUP_BYTES", "150000")
    monkeypatch.setenv("ZB200_UNC_GATED", gated)
    ctx = z.Context()
    monkeypatch.delenv("ZB200_UNC_GROUP_BYTES")
    monkeypatch.delenv("ZB200_UNC_GATED")
    rng = random.Random(5)
    raws = [corpus["alice29.txt"], b"", corpus["html"], corpus["urls.10K"], b"q" * 300000, corpus["geo.protodata"],
            bytes(rng.randrange(256) for _ in range(100000)), b"abc", corpus["lcet10.txt"], b""]
    items = [o.compress(r, [1, -1, 6, 0][i % 4], [o.dfGzip, o.dfZlib][i % 2]) for i, r in enumerate(raws)]
    items[3] = items[3][:-5] + bytes([items[3][-5] ^ 0x40]) + items[3][-4:]   # break one trailer
    base = np.frombuffer(b"".join(items), dtype=np.uint8)
    offs = np.zeros(len(items) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(x) for x in items])
    for registered in (False, True):   # pageable source, then a page-locked one (the destination stays pageable)
        if registered:
            base = base.copy()
            z.host_register(base.ctypes.data, base.nbytes)
        out, do, lens, st = ctx.uncompress_batch(base, offs, z.dfDetect)
        if registered:
            z.host_unregister(base.ctypes.data)
        for i, r in enumerate(raws):
            if i == 3:
                assert st[i] != 0
                continue
            assert st[i] == 0, (i, st[i])
            assert out[int(do[i]):int(do[i]) + int(lens[i])].tobytes() == r, i
    ctx.close()


def test_large_members_decode_as_parallel_segments(z, o, corpus, monkeypatch):
    """a large member whose stream is a chain of independent byte-aligned pieces (this
    library's multi-chunk members, zlib full-flush streams) is decoded as parallel segments; every
    other large member falls back to the serial decode.  Same bytes and the same accept / reject
    decision as the oracle in every case."""
    monkeypatch.setenv("ZB200_BIG_MEMBER_BYTES", "20000")
    ctx = z.Context()
    monkeypatch.delenv("ZB200_BIG_MEMBER_BYTES")
    T = util.text_corpus(corpus)
    raw = T[:700000] + corpus["urls.10K"][:300000] + bytes(1000) + corpus["fireworks.jpg"]

    def one(item, fmt=z.dfDetect):
        base = np.frombuffer(item, dtype=np.uint8)
        offs = np.array([0, len(item)], dtype=np.uint64)
        out, do, lens, st = ctx.uncompress_batch(base, offs, fmt)
        return (out[:int(lens[0])].tobytes() if st[0] == 0 else None), int(st[0]), ctx.timing()["kernel_launches"]

    # launches per call: 3 for the ordinary path (inflate + 2 verify); the segment path adds 2 (marker
    # search + optimistic pass) and, when the 64 KiB guess is wrong, 2 for the count pass and 1 for the
    # placed pass.  (a) our own multi-chunk members, all three wrappers: one optimistic pass
    for fmt in (z.dfGzip, z.dfZlib, z.dfDeflate):
        comp = ctx.compress_batch(np.frombuffer(raw, dtype=np.uint8), np.array([0, len(raw)], dtype=np.uint64), 1, fmt)
        blob = comp[0][:int(comp[1][1])].tobytes()
        got, st, launches = one(blob, fmt if fmt == z.dfDeflate else z.dfDetect)
        assert st == 0 and got == raw
        assert launches == 5, launches
    # (b) zlib full-flush stream (independent pieces) and (c) sync-flush stream (pieces depend on history)
    for flush_mode, parallel in ((zlib.Z_FULL_FLUSH, True), (zlib.Z_SYNC_FLUSH, False)):
        c = zlib.compressobj(6, zlib.DEFLATED, 31)
        blob = b"".join(c.compress(raw[i:i + 90000]) + c.flush(flush_mode) for i in range(0, len(raw), 90000)) + c.flush()
        got, st, launches = one(blob)
        assert st == 0 and got == raw
        # count + placed pass for independent pieces; pieces that depend on history fail the count pass and go on
        # to the speculative segments (block search, count, marker decode, two resolve kernels)
        assert (launches == 8) if parallel else (launches > 8), (flush_mode, launches)
    # (d) a stored block whose data contains the marker bytes: a false boundary, serial fallback
    tricky = b"\x00\x00\xff\xff" * 5000 + os.urandom(60000)
    blob = o.compress(tricky, 0, o.dfGzip)
    got, st, _ = one(blob)
    assert st == 0 and got == tricky
    # (e) corrupted large members: same decision as the oracle
    comp = ctx.compress_batch(np.frombuffer(raw, dtype=np.uint8), np.array([0, len(raw)], dtype=np.uint64), 1, z.dfGzip)
    good = comp[0][:int(comp[1][1])].tobytes()
    rng = random.Random(3)
    for _ in range(12):
        bad = bytearray(good)
        pos = rng.randrange(20, len(bad))
        bad[pos] ^= 1 << rng.randrange(8)
        got, st, _ = one(bytes(bad))
        try:
            want = o.uncompress(bytes(bad))
        except o.ZippyError:
            assert st != 0
            continue
        assert st == 0 and got == want
    got, st, _ = one(good[:len(good) // 2])
    assert st != 0
    # (f) the single-stream seam: size query and inflate of a large raw stream
    rawdef = ctx.deflate(raw, 1)
    assert ctx.inflate(rawdef) == raw
    ctx.close()
    # (g) at the default threshold, through the module-level call
    big = T * 8
    assert z.uncompress(z.compress(big, 1, z.dfGzip)) == big


def test_foreign_members_decode_as_speculative_segments(z, o, corpus, monkeypatch):
    """A large member from another encoder (zlib / gzip output: no sync markers, every block refers to the
    previous 32 KiB) is cut at block starts found by testing every bit offset, decoded in parallel with
    marker symbols for the unknown windows, and resolved.  Same bytes and the same accept / reject decision
    as the oracle; anything irregular falls back to the serial decode (inflate.nim:104-291 is the behaviour
    to match)."""
    monkeypatch.setenv("ZB200_BIG_MEMBER_BYTES", "100000")
    ctx = z.Context()
    monkeypatch.delenv("ZB200_BIG_MEMBER_BYTES")
    T = util.text_corpus(corpus)
    raw = T[:2500000] + corpus["urls.10K"] + bytes(70000) + corpus["fireworks.jpg"] + corpus["html_x_4"] + T[:300000]

    def one(item, fmt=z.dfDetect):
        base = np.frombuffer(item, dtype=np.uint8)
        offs = np.array([0, len(item)], dtype=np.uint64)
        out, do, lens, st = ctx.uncompress_batch(base, offs, fmt)
        return (out[int(do[0]):int(do[0]) + int(lens[0])].tobytes() if st[0] == 0 else None), int(st[0]), ctx.timing()["kernel_launches"]

    streams = {}
    for lvl in (1, 6, 9):
        c = zlib.compressobj(lvl, zlib.DEFLATED, 31)
        streams["gzip%d" % lvl] = c.compress(raw) + c.flush()
    streams["zlib6"] = zlib.compress(raw, 6)
    c = zlib.compressobj(6, zlib.DEFLATED, -15)
    streams["raw6"] = c.compress(raw) + c.flush()
    streams["oracle6"] = o.compress(raw, 6, o.dfGzip)          # the reference's own encoder: one block per 4 MiB
    for name, blob in streams.items():
        fmt = z.dfDeflate if name == "raw6" else z.dfDetect
        got, st, launches = one(blob, fmt)
        assert st == 0 and got == raw, name
        if name in ("gzip6", "zlib6", "raw6"):
            assert launches >= 9, (name, launches)   # block search, count, prefill + marker decode, 2 resolves, verify
    # sizes without decoding twice, single call
    assert ctx.decode_one(streams["zlib6"]) == raw and ctx.inflate(streams["raw6"]) == raw
    # corrupted / truncated members: the oracle's verdict
    rng = random.Random(5)
    good = streams["gzip6"]
    for k in range(16):
        bad = bytearray(good)
        pos = rng.randrange(20, len(bad))
        bad[pos] ^= 1 << rng.randrange(8)
        got, st, _ = one(bytes(bad))
        try:
            want = o.uncompress(bytes(bad))
        except o.ZippyError as e:
            assert st == e.code, (k, pos, st, e.code)
            continue
        assert st == 0 and got == want
    for cut in (len(good) // 3, len(good) - 9, len(good) - 1):
        got, st, _ = one(good[:cut])
        with pytest.raises(o.ZippyError) as e:
            o.uncompress(good[:cut])
        assert st == e.value.code, (cut, st, e.value.code)
    # a batch that mixes small members, a foreign large member and one of this library's own
    own = ctx.compress_batch(np.frombuffer(raw, dtype=np.uint8), np.array([0, len(raw)], dtype=np.uint64), 1, z.dfGzip)
    items = [o.compress(util.c2_block(T, i), 1, o.dfGzip) for i in range(40)] + [streams["gzip9"], own[0][:int(own[1][1])].tobytes(),
                                                                               streams["zlib6"]]
    base, offs = z._pack(items)
    out, do, lens, st = ctx.uncompress_batch(base, offs)
    assert not st.any()
    for i in range(40):
        assert out[int(do[i]):int(do[i]) + int(lens[i])].tobytes() == util.c2_block(T, i)
    for i in (40, 41, 42):
        assert out[int(do[i]):int(do[i]) + int(lens[i])].tobytes() == raw
    ctx.close()


def test_speculative_segments_reject_false_block_starts(z, o, corpus, monkeypatch):
    """Adversarial input for the block search: the DATA is itself a deflate stream, carried in stored blocks (or
    barely compressible), so every dynamic-block header of the inner stream is a plausible -- and false -- block
    start of the outer stream.  The counting pass must notice (a segment does not end on the next boundary, or
    fails to decode) and the member must come out right through the serial decode, with the oracle's verdict for
    corrupted variants too."""
    monkeypatch.setenv("ZB200_BIG_MEMBER_BYTES", "60000")
    ctx = z.Context()
    monkeypatch.delenv("ZB200_BIG_MEMBER_BYTES")
    T = util.text_corpus(corpus)
    inner = zlib.compress(T[:1500000], 6)            # ~600 KB of compressed bytes with ~40 dynamic-block headers
    streams = {
        "stored": o.compress(inner, 0, o.dfGzip),    # level 0: stored blocks only
        "zlib6_of_compressed": zlib.compress(inner, 6),
        "mixed": zlib.compress(T[:400000] + inner + T[400000:900000] + inner[:200000], 6),
    }
    c = zlib.compressobj(1, zlib.DEFLATED, 31, 9, zlib.Z_FIXED)   # fixed-Huffman blocks only: nothing to find
    streams["fixed_only"] = c.compress(T[:700000]) + c.flush()
    c = zlib.compressobj(6, zlib.DEFLATED, 31)
    tiny = b"".join(c.compress(T[i:i + 1500]) + c.flush(zlib.Z_FULL_FLUSH) for i in range(0, 300000, 1500)) + c.flush()
    streams["many_tiny_blocks"] = tiny
    want = {"stored": inner, "zlib6_of_compressed": inner, "mixed": T[:400000] + inner + T[400000:900000] + inner[:200000],
            "fixed_only": T[:700000], "many_tiny_blocks": T[:300000]}
    rng = random.Random(11)
    for name, blob in streams.items():
        base = np.frombuffer(blob, dtype=np.uint8)
        offs = np.array([0, len(blob)], dtype=np.uint64)
        out, do, lens, st = ctx.uncompress_batch(base, offs)
        assert st[0] == 0 and out[int(do[0]):int(do[0]) + int(lens[0])].tobytes() == want[name], name
        for _ in range(4):
            bad = bytearray(blob)
            pos = rng.randrange(12, len(bad) - 8)
            bad[pos] ^= 1 << rng.randrange(8)
            out, do, lens, st = ctx.uncompress_batch(np.frombuffer(bytes(bad), dtype=np.uint8), offs)
            try:
                ref = o.uncompress(bytes(bad))
            except o.ZippyError as e:
                assert st[0] == e.code, (name, pos, int(st[0]), e.code)
                continue
            assert st[0] == 0 and out[int(do[0]):int(do[0]) + int(lens[0])].tobytes() == ref, (name, pos)
    ctx.close()


def test_default_level_ratio_vs_reference(z, o, corpus):
    """BASELINE config 4: at level=Default the total compressed size on the urls.10K corpus must
    stay within 3 % of the reference's (oracle port, hash-chain level 6, deflate.nim:262-272)."""
    raw = corpus["urls.10K"]
    for data in (raw, raw[:65536], corpus["alice29.txt"], corpus["html_x_4"]):
        ref = len(o.deflate(data, o.DefaultCompression))
        got = z.deflate(data, z.DefaultCompression)
        assert zlib.decompress(got, -15) == data and o.inflate(got) == data
        if data is raw:
            assert len(got) <= 1.03 * ref, (len(got), ref)
        assert len(got) <= 1.08 * ref, (len(got), ref)
    tiles = [raw] * 6
    comp = z.compress_batch(tiles, z.DefaultCompression, z.dfGzip)
    assert all(o.uncompress(c) == raw for c in comp)
    assert sum(map(len, comp)) <= 1.03 * 6 * len(o.compress(raw, o.DefaultCompression, o.dfGzip))
    lvl1 = len(z.deflate(raw, 1))
    assert len(z.deflate(raw, 9)) < lvl1 and len(z.deflate(raw, 2)) < lvl1


def test_levels_are_monotone_and_track_the_reference(z, o, corpus):
    """internal.nim:177-189 gives every level its own search effort; here the per-level budgets (verified
    candidates, good, lazy) must make the output shrink (or stay) as the level rises, Default must equal
    level 6, and levels 2 and 9 must stay within 6 % / 8 % of the reference's levels 2 and 9."""
    for name in ("urls.10K", "alice29.txt", "html"):
        data = corpus[name]
        sizes = {}
        for lvl in (2, 3, 4, 5, 6, 7, 8, 9):
            got = z.deflate(data, lvl)
            assert zlib.decompress(got, -15) == data and o.inflate(got) == data, (name, lvl)
            sizes[lvl] = len(got)
        for a, b in zip((2, 3, 4, 5, 6, 7, 8), (3, 4, 5, 6, 7, 8, 9)):
            assert sizes[b] <= sizes[a] * 1.002, (name, a, b, sizes)
        assert len(z.deflate(data, z.DefaultCompression)) == sizes[6]
        assert sizes[9] < sizes[2] and sizes[9] <= 1.08 * len(o.deflate(data, 9)), (name, sizes)
        assert sizes[2] <= 1.06 * len(o.deflate(data, 2)), (name, sizes)


def test_output_is_identical_run_to_run(z, corpus):
    """No kernel shares mutable state across warps; the one unordered operation (lanes of ONE store
    instruction that hash to the same table entry, level 1) is resolved by the hardware the same way every
    time.  Asserted here instead of in prose: two contexts, several runs, byte-identical members."""
    T = util.text_corpus(corpus)
    items = [util.c2_block(T, i) for i in range(96)] + [corpus["urls.10K"], corpus["html_x_4"], corpus["kppkn.gtb"]]
    for lvl in (1, z.DefaultCompression, 9):
        ref = None
        for rep in range(3):
            ctx = z.Context() if rep == 2 else z.default_context()
            base, offs = z._pack(items)
            out, oo = ctx.compress_batch(base, offs, lvl, z.dfZlib)
            blob = out.tobytes() + oo.tobytes()
            if ref is None:
                ref = blob
            assert blob == ref, (lvl, rep)
            if rep == 2:
                ctx.close()


def test_size_claims_get_the_reference_verdict(z, o, corpus):
    """gzip.nim:80-88 inflates whatever the stream holds, checks the CRC, then ISIZE.  A member whose ISIZE
    understates (or overstates) its content must therefore end in "Size verification failed" -- not in
    "destination too small" -- and a wrong CRC in "Checksum verification failed", through the single call and
    through the batch call, with the same codes as the oracle."""
    data = corpus["alice29.txt"]
    good = bytearray(o.compress(data, 1, o.dfGzip))
    cases = {}
    for name, delta in (("isize_small", -1000), ("isize_big", 5000), ("isize_zero", -len(data))):
        m = bytearray(good)
        m[-4:] = ((len(data) + delta) & 0xffffffff).to_bytes(4, "little")
        cases[name] = bytes(m)
    m = bytearray(good)
    m[-8] ^= 0x55
    cases["crc"] = bytes(m)
    m[-4:] = (len(data) - 7).to_bytes(4, "little")
    cases["crc_and_isize"] = bytes(m)
    want = {}
    for name, m in cases.items():
        with pytest.raises(o.ZippyError) as e:
            o.uncompress(m)
        want[name] = e.value.code
        with pytest.raises(z.ZippyError) as e2:
            z.uncompress(m)
        assert e2.value.code == want[name], (name, e2.value.code, want[name])
    assert want["isize_small"] == 18 and want["crc"] == 14 and want["crc_and_isize"] == 14
    res = z.uncompress_batch([bytes(good)] + list(cases.values()) + [bytes(good)])
    assert res[0] == data and res[-1] == data
    for (name, _), r in zip(cases.items(), res[1:-1]):
        assert isinstance(r, z.ZippyError) and r.code == want[name], (name, r)
    # a zlib stream (no size field at all) through the single call: decoded once, sized by the library
    zs = zlib.compress(corpus["html_x_4"], 6)
    assert z.uncompress(zs) == corpus["html_x_4"] and z.inflate(zs, 2) == corpus["html_x_4"]


def test_two_devices_two_threads(z, o, corpus):
    """Function attributes (dynamic shared memory limits) are per device: every ctx sets them for its own
    device in zb200_init.  Two contexts on two devices, driven from two host threads at once."""
    import threading
    from zippy_b200 import _native
    if _native.lib().zb200_device_count() < 2:
        pytest.skip("needs two CUDA devices")
    T = util.text_corpus(corpus)
    items = [util.c2_block(T, i) for i in range(64)] + [corpus["urls.10K"]]
    errs = []

    def work(dev):
        try:
            ctx = z.Context(dev)
            base, offs = z._pack(items)
            for lvl in (1, z.DefaultCompression):
                out, oo = ctx.compress_batch(base, offs, lvl, z.dfGzip)
                back, do, lens, st = ctx.uncompress_batch(out, oo)
                assert not st.any()
                for i, it in enumerate(items):
                    assert back[int(do[i]):int(do[i]) + int(lens[i])].tobytes() == it
            assert ctx.crc32(items[0]) == zlib.crc32(items[0])
            ctx.close()
        except Exception as e:  # noqa: BLE001
            errs.append((dev, repr(e)))

    ts = [threading.Thread(target=work, args=(d,)) for d in (0, 1)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs


def test_multi_gpu_entry_points(z, o, corpus):
    """zb200_mgpu_*: members sharded over devices inside ONE call, the shards' bytes concatenated at the gathered
    offsets.  With a single GPU the same code runs with that device listed twice (two ctxs, two host threads)."""
    from zippy_b200 import _native
    nd = _native.lib().zb200_device_count()
    mg = z.MultiGpu(list(range(nd)) if nd > 1 else [0, 0])
    assert mg.device_count() == max(nd, 2)
    T = util.text_corpus(corpus)
    items = [util.c2_block(T, i, 4000 + 977 * i) for i in range(60)] + [b"", corpus["urls.10K"], corpus["html"], b"x"]
    base, offs = z._pack(items)
    out, oo = mg.compress_batch(base, offs, 1, z.dfGzip)
    assert int(oo[-1]) == len(out) and (np.diff(oo.astype(np.int64)) > 0).all()
    for i, it in enumerate(items):
        assert o.uncompress(out[int(oo[i]):int(oo[i + 1])].tobytes()) == it, i
    back, do, lens, st = mg.uncompress_batch(out, oo, [len(x) for x in items])
    assert not st.any()
    for i, it in enumerate(items):
        assert back[int(do[i]):int(do[i]) + int(lens[i])].tobytes() == it, i
    assert [int(c) for c in mg.checksum_batch(base, offs)] == [zlib.crc32(x) for x in items]
    mg.close()


def test_cpp_host_mirror():
    """include/zippy_b200.hpp (host framing in C++ as in zippy.nim, codec through the seam)."""
    import subprocess
    from tests.test_abi import _build_cpp_api_test
    out = subprocess.run([_build_cpp_api_test()], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and out.stdout.startswith("OK"), (out.stdout, out.stderr)


def test_mixed_entropy_corpus(z, o, corpus):
    """BASELINE config 5 at reduced count (SURVEY 8d "C5"): 64 KiB blocks whose class is drawn per block --
    text, urls, html, uniform random (stored path), run-length blobs -- compressed at level 1 and
    inflated again on the GPU; every 16th member is also checked by the oracle and zlib."""
    rng = random.Random(0xC5)
    T = util.text_corpus(corpus)
    srcs = {"urls": corpus["urls.10K"], "html": corpus["html"] * 2}
    blocks = []
    for i in range(768):
        cls = util._sm64(0xC5 + i) % 8
        if cls <= 3:
            blocks.append(util.c2_block(T, i))
        elif cls == 4:
            s0 = rng.randrange(len(srcs["urls"]) - 65536)
            blocks.append(srcs["urls"][s0:s0 + 65536])
        elif cls == 5:
            s0 = rng.randrange(len(srcs["html"]) - 65536)
            blocks.append(srcs["html"][s0:s0 + 65536])
        elif cls == 6:
            blocks.append(rng.randbytes(65536))
        else:
            b = bytearray()
            while len(b) < 65536:
                b += bytes([rng.randrange(256)]) * rng.randrange(256)
            blocks.append(bytes(b[:65536]))
    comp = z.compress_batch(blocks, z.BestSpeed, z.dfGzip)
    for i in range(0, len(blocks), 16):
        assert o.uncompress(comp[i]) == blocks[i] and zlib.decompress(comp[i], 31) == blocks[i], i
    back = z.uncompress_batch(comp)
    assert all(r == b for r, b in zip(back, blocks))
    # incompressible blocks must fall back to stored blocks: bounded expansion (deflate.nim:274-277)
    for b, c in zip(blocks, comp):
        assert len(c) <= len(b) + 10 + 36


def test_device_pointers_of_any_alignment(z, o, corpus):
    """Source buffers are staged by 16-byte TMA granules from wherever they start: give the device
    variants a source pointer that is off by 1..15 bytes and members of odd sizes."""
    torch = pytest.importorskip("torch")
    ctx = z.Context()
    ctx.set_stream(ctx.LEGACY_DEFAULT_STREAM)   # ordered with torch's default-stream work on the same buffers
    items = [corpus["alice29.txt"][:70001], corpus["html"][:12345], b"", corpus["urls.10K"][:65537], b"q" * 31]
    blob = b"".join(items)
    offs = np.zeros(len(items) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(x) for x in items])
    for shift in (1, 7, 15):
        d_all = torch.zeros(len(blob) + 64, dtype=torch.uint8, device="cuda")
        d_all[shift:shift + len(blob)] = torch.frombuffer(bytearray(blob), dtype=torch.uint8).cuda()
        cap = len(blob) + 4096
        d_dst = torch.empty(cap, dtype=torch.uint8, device="cuda")
        for level in (1, -1):
            oo = ctx.compress_batch_device(d_all.data_ptr() + shift, offs, level, z.dfGzip, d_dst.data_ptr(), cap)
            host = d_dst.cpu().numpy()
            for i, x in enumerate(items):
                assert o.uncompress(host[int(oo[i]):int(oo[i + 1])].tobytes()) == x, (shift, level, i)
        crcs = ctx.checksum_batch_device(d_all.data_ptr() + shift, offs, "crc32")
        assert [int(c) for c in crcs] == [zlib.crc32(x) for x in items]
        # inflate from a misaligned compressed buffer into a misaligned output buffer
        comp = b"".join(o.compress(x, 1, o.dfGzip) for x in items)
        coffs = np.zeros(len(items) + 1, dtype=np.uint64)
        coffs[1:] = np.cumsum([len(o.compress(x, 1, o.dfGzip)) for x in items])
        d_c = torch.zeros(len(comp) + 64, dtype=torch.uint8, device="cuda")
        d_c[shift:shift + len(comp)] = torch.frombuffer(bytearray(comp), dtype=torch.uint8).cuda()
        d_o = torch.zeros(len(blob) + 64, dtype=torch.uint8, device="cuda")
        lens, st = ctx.uncompress_batch_device(d_c.data_ptr() + shift, coffs, z.dfDetect, d_o.data_ptr() + shift, offs)
        assert not st.any() and [int(l) for l in lens] == [len(x) for x in items]
        assert d_o[shift:shift + len(blob)].cpu().numpy().tobytes() == blob
    ctx.close()
