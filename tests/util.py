"""Shared helpers for the parity tests: golden fixtures and seeded synthetic inputs.

Nothing here reads /root/reference (absent on the GPU box); the fixtures were
copied into tests/golden/ by tests/golden/make_golden.py.
"""
import hashlib
import json
import os
import random
import zlib

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden():
    """-> {fixture name: (compressed bytes, manifest entry)}"""
    manifest = json.load(open(os.path.join(GOLDEN_DIR, "manifest.json")))
    out = {}
    for name, meta in sorted(manifest.items()):
        out[name] = (open(os.path.join(GOLDEN_DIR, name), "rb").read(), meta)
    return out


def sha(b):
    return hashlib.sha256(bytes(b)).hexdigest()


def load_corpus():
    """Uncompressed originals of the fixtures, recovered with *system zlib* (a third
    party to both the oracle and the CUDA path) and checked against the manifest."""
    out = {}
    for name, (comp, meta) in load_golden().items():
        raw = zlib.decompress(comp, 47 if name != "fixed.z" else 15)
        assert len(raw) == meta["len"] and sha(raw) == meta["sha256"], name
        key = meta["gold"] or name
        out[key] = raw
    out["all_uint8"] = bytes(range(256))  # tests/test.nim:73-85
    return out


def text_corpus(corpus):
    """T = alice29 || asyoulik || lcet10 || plrabn12."""
    return corpus["alice29.txt"] + corpus["asyoulik.txt"] + corpus["lcet10.txt"] + corpus["plrabn12.txt"]


def _sm64(x):
    x = (x + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return z ^ (z >> 31)


def c2_block(T, i, size=65536):
    """Block i of BASELINE config 2: T[o:o+size], o = sm64(0xC2+i) mod (|T|-size)."""
    o = _sm64(0xC2 + i) % (len(T) - size)
    return T[o:o + size]


def run_length_blob(rng, max_len=100000):
    """tests/stress.nim:13-24: random byte x random run (<=255) until a random length."""
    n = rng.randrange(max_len)
    out = bytearray()
    while len(out) < n:
        out += bytes([rng.randrange(256)]) * rng.randrange(256)
    return bytes(out[:n])


def edge_inputs(seed=1234):
    """Small adversarial inputs for round trips: empty, tiny, runs, near block-size edges."""
    rng = random.Random(seed)
    xs = [b"", b"a", b"ab", b"abc", b"abcd", b"aaaa", b"a" * 5, b"a" * 258, b"a" * 259, b"a" * 300,
          b"ab" * 200, b"abc" * 1000, bytes(range(256)), bytes(range(256)) * 20,
          b"\x00" * 65535, b"\x00" * 65536, b"\x00" * 65537, b"\xff" * 100000]
    # (4 KiB = a parse piece of the level-1 matcher, 32 KiB = one of its two phases, 8 KiB = a packer sub-chunk)
    for n in (1, 14, 15, 16, 31, 32, 33, 63, 64, 65, 2047, 2048, 2049, 4095, 4096, 4097, 8191, 8192, 8193,
              28671, 28672, 28673, 32767, 32768, 32769, 33791, 33792, 33793, 36863, 36864, 36865, 61439, 61440, 61441,
              65535, 65536, 65537, 131072, 200001):
        xs.append(bytes(rng.randrange(256) for _ in range(n)))                    # incompressible
        xs.append(bytes(rng.choice(b"abcdefgh ") for _ in range(n)))              # low entropy
    for _ in range(8):
        xs.append(run_length_blob(rng, 70000))
    return xs
