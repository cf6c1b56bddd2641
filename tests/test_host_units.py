"""CPU-only unit tests of the host/device-shared maths in zippy_b200/csrc/zb_huff.h,
zb_crc.h, zb_common.h (compiled here with g++; the GPU kernels include the same code)."""
import ctypes
import hashlib
import os
import random
import subprocess
import zlib

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "native", "host_units.cpp")
SO = os.path.join(HERE, "native", "libhost_units.so")


@pytest.fixture(scope="module")
def hu():
    deps = [SRC] + [os.path.join(HERE, "..", "zippy_b200", "csrc", f) for f in ("zb_huff.h", "zb_crc.h", "zb_common.h")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", SO, SRC])
    L = ctypes.CDLL(SO)
    for f in ("t_gf2_mul", "t_xpow8", "t_xpow8_t", "t_crc_combine", "t_adler_combine", "t_crc32_lane_model", "t_crc32_piece_model", "t_adler32_model",
              "t_dist_base", "t_len_base"):
        getattr(L, f).restype = ctypes.c_uint32
    L.t_xpow8.argtypes = [ctypes.c_uint64]
    L.t_xpow8_t.argtypes = [ctypes.c_uint64]
    L.t_crc_combine.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint64]
    L.t_adler_combine.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint64]
    L.t_crc32_lane_model.argtypes = [ctypes.c_char_p, ctypes.c_uint64]
    L.t_adler32_model.argtypes = [ctypes.c_char_p, ctypes.c_uint64]
    L.t_crc32_piece_model.argtypes = [ctypes.c_char_p]
    return L


def test_rfc_tables(hu):
    base_len = [3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163,
                195, 227, 258]
    ext_len = [0] * 8 + [1] * 4 + [2] * 4 + [3] * 4 + [4] * 4 + [5] * 4 + [0]
    base_d = [1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073,
              4097, 6145, 8193, 12289, 16385, 24577]
    for c in range(29):
        assert hu.t_len_base(c) == base_len[c] and hu.t_len_extra(c) == ext_len[c]
    for c in range(30):
        assert hu.t_dist_base(c) == base_d[c] and hu.t_dist_extra(c) == (0 if c < 4 else c // 2 - 1)
    for l in range(3, 259):
        c = hu.t_len_code(l)
        assert base_len[c] <= l and (c == 28 or l < base_len[c + 1]) and (l != 258 or c == 28)
    for d in range(1, 32769):
        c = hu.t_dist_code(d)
        assert base_d[c] <= d and (c == 29 or d < base_d[c + 1])


def _kraft(lens, limit):
    return sum(1 << (limit - l) for l in lens if l)


def _opt_cost(freq):
    import heapq
    h = [f for f in freq if f]
    if len(h) < 2:
        return sum(h)
    heapq.heapify(h)
    cost = 0
    while len(h) > 1:
        a, b = heapq.heappop(h), heapq.heappop(h)
        cost += a + b
        heapq.heappush(h, a + b)
    return cost


@pytest.mark.parametrize("n,limit", [(286, 15), (30, 15), (19, 7)])
def test_huff_lengths(hu, n, limit):
    rng = random.Random(n * 100 + limit)
    cases = []
    for _ in range(300):
        kind = rng.randrange(6)
        if kind == 0:
            f = [rng.randrange(1000) for _ in range(n)]
        elif kind == 1:
            f = [rng.randrange(3) * rng.randrange(2) for _ in range(n)]
        elif kind == 2:  # geometric / fibonacci-like: forces length limiting
            f = [min(65535, int(1.6 ** i)) for i in range(n)]
            rng.shuffle(f)
        elif kind == 3:
            f = [0] * n
            for _ in range(rng.randrange(1, 4)):
                f[rng.randrange(n)] = rng.randrange(1, 65536)
        elif kind == 4:
            f = [1] * n
        else:
            f = [int(65535 / (i + 1)) for i in range(n)]
        cases.append(f)
    cases += [[0] * n, [0] * (n - 1) + [5], [7] + [0] * (n - 1), [1, 1] + [0] * (n - 2)]
    for f in cases:
        fa = (ctypes.c_uint32 * n)(*f)
        lens = (ctypes.c_uint8 * n)()
        hu.t_huff_lengths(fa, n, limit, lens)
        lens = list(lens)
        used = [i for i in range(n) if f[i]]
        assert all(lens[i] > 0 for i in used)
        assert max(lens) <= limit
        assert _kraft(lens, limit) == 1 << limit, (f, lens)  # complete
        if len(used) >= 2:
            assert sum(1 for l in lens if l) == len(used)
            cost = sum(f[i] * lens[i] for i in range(n))
            opt = _opt_cost(f)
            assert cost >= opt
            if max(lens) < limit:
                assert cost == opt  # unconstrained => optimal


def test_codebook_dynamic_header_decodes_with_zlib(hu):
    """Build a codebook from random token statistics, emit a block with a pure-Python
    packer using that codebook, and let zlib decode it."""
    rng = random.Random(3)
    nbytes = hu.t_codebook_size()
    for trial in range(40):
        n = rng.choice([0, 1, 5, 100, 3000, 65536])
        data = bytes(rng.choice(b"abcdefghijklmnop \n") if rng.random() < 0.9 else rng.randrange(256) for _ in range(n))
        # tokens: greedy toy LZ with python dict
        toks, i, last = [], 0, {}
        while i < n:
            k = data[i:i + 4]
            j = last.get(k, -1) if len(k) == 4 else -1
            if j >= 0 and i - j <= 32768:
                l = 4
                while l < 258 and i + l < n and data[j + l] == data[i + l]:
                    l += 1
                toks.append((l, i - j))
                for q in range(i, i + l):
                    last[data[q:q + 4]] = q
                i += l
            else:
                toks.append((0, data[i]))
                last[k] = i
                i += 1
        hist = np.zeros((8, 316), dtype=np.uint16)
        per = (len(toks) + 7) // 8 if toks else 0
        for t, (l, d) in enumerate(toks):
            w = t // per if per else 0
            if l:
                hist[w, 257 + hu.t_len_code(l)] += 1
                hist[w, 286 + hu.t_dist_code(d)] += 1
            else:
                hist[w, d] += 1
        for is_final in (1, 0):
            cb = (ctypes.c_uint8 * nbytes)()
            hu.t_build_codebook(hist.ctypes.data_as(ctypes.POINTER(ctypes.c_uint16)), n, is_final, -1, cb)
            raw = bytes(cb)
            u32 = np.frombuffer(raw[:4 * (288 + 32 + 2 + 8 + 4)], dtype=np.uint32)
            ll, dd = u32[:288], u32[288:320]
            btype, hdr_bits = int(u32[320]), int(u32[321])
            wstart, eob, total_bytes = u32[322:330], int(u32[330]), int(u32[331])
            if btype == 0:
                assert total_bytes == n + 5 * max(1, (n + 65534) // 65535)
                continue
            hdr = raw[4 * 334:4 * 334 + 336]
            bits = []
            for b in range(hdr_bits):
                bits.append((hdr[b >> 3] >> (b & 7)) & 1)

            def put(v, nb):
                for k in range(nb):
                    bits.append((v >> k) & 1)
            for t, (l, d) in enumerate(toks):
                if per and t % per == 0:
                    assert len(bits) == int(wstart[t // per]), (t, len(bits), wstart)
                if l:
                    c = hu.t_len_code(l)
                    put(int(ll[257 + c]) & 0xffff, int(ll[257 + c]) >> 16)
                    put(l - hu.t_len_base(c), hu.t_len_extra(c))
                    c = hu.t_dist_code(d)
                    put(int(dd[c]) & 0xffff, int(dd[c]) >> 16)
                    put(d - hu.t_dist_base(c), hu.t_dist_extra(c))
                else:
                    put(int(ll[d]) & 0xffff, int(ll[d]) >> 16)
            assert len(bits) == eob
            put(int(ll[256]) & 0xffff, int(ll[256]) >> 16)
            if not is_final:
                put(0, 3)
            while len(bits) % 8:
                bits.append(0)
            out = bytearray(len(bits) // 8)
            for k, b in enumerate(bits):
                out[k >> 3] |= b << (k & 7)
            if not is_final:
                out += b"\x00\x00\xff\xff"
            assert len(out) == total_bytes
            if not is_final:
                out += b"\x01\x00\x00\xff\xff"  # terminate the stream for zlib
            assert zlib.decompress(bytes(out), -15) == data


def test_codebooks_are_pinned(hu):
    """The codebook of a histogram is part of the compressed bytes: a change to zb_huff.h that is meant to be an
    optimisation (the counting sort, the byte-wise header writer) must leave every codebook as it was.  200 seeded
    histograms (dense, sparse, text-like with matches, a single symbol, empty), digest of the ZbCodebook structs."""
    nbytes = hu.t_codebook_size()
    rng = np.random.default_rng(2024)
    h256 = hashlib.sha256()
    for t in range(200):
        kind = t % 5
        h = np.zeros((8, 316), dtype=np.uint16)
        if kind == 0:
            h[:, :256] = rng.integers(0, 40, (8, 256))
        elif kind == 1:
            h[:, rng.integers(0, 286, 20)] = rng.integers(1, 3000, (8, 20))
        elif kind == 2:
            h[:, 32:127] = rng.integers(0, 200, (8, 95))
            h[:, 257:280] = rng.integers(0, 60, (8, 23))
            h[:, 286:316] = rng.integers(0, 50, (8, 30))
        elif kind == 3:
            h[0, rng.integers(0, 256)] = rng.integers(1, 8000)
        ln = min(int(h[:, :256].sum() + 3 * h[:, 257:286].sum()), 65536)
        cb = ctypes.create_string_buffer(nbytes)
        hu.t_build_codebook(h.ctypes.data_as(ctypes.POINTER(ctypes.c_uint16)), ln, t & 1, -1, cb)
        h256.update(cb.raw)
    assert h256.hexdigest() == "a2273583dc12be3211016a13c0ec7138f76bd9e11c3531d6855fe4cc59ea0679"


def test_crc_math(hu):
    rng = random.Random(11)
    for n in [0, 1, 2, 3, 4, 5, 7, 8, 127, 128, 129, 131, 255, 256, 1000, 4096, 8191, 8192, 8193, 65536, 100001]:
        x = bytes(rng.randrange(256) for _ in range(n))
        assert hu.t_crc32_lane_model(x, n) == zlib.crc32(x), n
        assert hu.t_adler32_model(x, n) == zlib.adler32(x), n
    x = b"\xff" * 70000
    assert hu.t_adler32_model(x, len(x)) == zlib.adler32(x)
    for _ in range(3):   # the checksum kernel's full-piece CRC path (chains, joins, the fold of the partial words)
        x = bytes(rng.randrange(256) for _ in range(32768))
        assert hu.t_crc32_piece_model(x) == zlib.crc32(x)
    for _ in range(50):
        a = bytes(rng.randrange(256) for _ in range(rng.randrange(0, 3000)))
        b = bytes(rng.randrange(256) for _ in range(rng.randrange(0, 3000)))
        assert hu.t_crc_combine(zlib.crc32(a), zlib.crc32(b), len(b)) == zlib.crc32(a + b)
        assert hu.t_adler_combine(zlib.adler32(a), zlib.adler32(b), len(b)) == zlib.adler32(a + b)
    assert hu.t_crc_combine(zlib.crc32(b"x" * 5), zlib.crc32(b""), 0) == zlib.crc32(b"x" * 5)
    big = 5 * 2 ** 32 + 12345
    assert hu.t_adler_combine(zlib.adler32(b"abc"), 1, 0) == zlib.adler32(b"abc")
    assert hu.t_xpow8(big) == hu.t_gf2_mul(hu.t_xpow8(5 * 2 ** 32), hu.t_xpow8(12345))
    for n in [0, 1, 2, 3, 255, 32768, 65536, 65537, 100001, 2 ** 32 - 1, big, 2 ** 47 - 5]:
        assert hu.t_xpow8_t(n) == hu.t_xpow8(n), n   # the table-driven form the kernels use
