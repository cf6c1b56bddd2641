"""ctypes loader for the CPU oracle (oracle/zippy_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  zippy_b200/ must never import it.
Mirrors the reference's public names (src/zippy.nim:11-177, crc.nim:53,
adler32.nim:6) so the parity tests read like the reference's own tests.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")

dfDetect, dfZlib, dfGzip, dfDeflate = 0, 1, 2, 3
NoCompression, BestSpeed, BestCompression, DefaultCompression, HuffmanOnly = 0, 1, 9, -1, -2


class ZippyError(Exception):
    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


class _Buf(ctypes.Structure):
    _fields_ = [("data", ctypes.POINTER(ctypes.c_uint8)), ("len", ctypes.c_size_t), ("cap", ctypes.c_size_t)]


def build(force=False):
    src = os.path.join(_HERE, "zippy_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = ctypes.CDLL(_SO)
        L.zo_crc32.restype = ctypes.c_uint32
        L.zo_crc32.argtypes = [ctypes.c_char_p, ctypes.c_size_t]
        L.zo_adler32.restype = ctypes.c_uint32
        L.zo_adler32.argtypes = [ctypes.c_char_p, ctypes.c_size_t]
        for nm in ("zo_crc32_scalar", "zo_adler32_scalar"):
            getattr(L, nm).restype = ctypes.c_uint32
            getattr(L, nm).argtypes = [ctypes.c_char_p, ctypes.c_size_t]
        L.zo_deflate.argtypes = [ctypes.POINTER(_Buf), ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int]
        L.zo_inflate.argtypes = [ctypes.POINTER(_Buf), ctypes.c_char_p, ctypes.c_size_t, ctypes.c_size_t]
        L.zo_compress.argtypes = [ctypes.POINTER(_Buf), ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int,
                                  ctypes.c_int, ctypes.c_int]
        L.zo_uncompress.argtypes = [ctypes.POINTER(_Buf), ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int]
        L.zo_buf_free.argtypes = [ctypes.POINTER(_Buf)]
        L.zo_strerror.restype = ctypes.c_char_p
        for f in (L.zo_compress_batch, L.zo_uncompress_batch):
            f.restype = ctypes.c_uint64
        L.zo_compress_batch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int,
                                        ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        L.zo_uncompress_batch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int,
                                          ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        _lib = L
    return _lib


def _take(buf, rc):
    L = lib()
    try:
        if rc != 0:
            raise ZippyError(rc, L.zo_strerror(rc).decode())
        return ctypes.string_at(buf.data, buf.len) if buf.len else b""
    finally:
        L.zo_buf_free(ctypes.byref(buf))


def crc32(data):
    return lib().zo_crc32(bytes(data), len(data))


def adler32(data):
    return lib().zo_adler32(bytes(data), len(data))


def deflate(data, level=DefaultCompression):
    b = _Buf()
    return _take(b, lib().zo_deflate(ctypes.byref(b), bytes(data), len(data), level))


def inflate(data, pos=0):
    b = _Buf()
    return _take(b, lib().zo_inflate(ctypes.byref(b), bytes(data), len(data), pos))


def compress(data, level=DefaultCompression, dataFormat=dfGzip, fname_len=0):
    b = _Buf()
    return _take(b, lib().zo_compress(ctypes.byref(b), bytes(data), len(data), level, dataFormat, fname_len))


def uncompress(data, dataFormat=dfDetect):
    b = _Buf()
    return _take(b, lib().zo_uncompress(ctypes.byref(b), bytes(data), len(data), dataFormat))


def compress_batch(base, offsets, level=BestSpeed, dataFormat=dfGzip, threads=1):
    """base: bytes-like (numpy uint8 array ok); offsets: numpy uint64[n+1]. Returns (total, lens, statuses)."""
    import numpy as np
    base = np.ascontiguousarray(base, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    n = len(offsets) - 1
    lens = np.zeros(n, dtype=np.uint64)
    st = np.zeros(n, dtype=np.int32)
    total = lib().zo_compress_batch(base.ctypes.data, offsets.ctypes.data, n, level, dataFormat, threads,
                                    lens.ctypes.data, st.ctypes.data)
    return total, lens, st


def uncompress_batch(base, offsets, dataFormat=dfDetect, threads=1):
    import numpy as np
    base = np.ascontiguousarray(base, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    n = len(offsets) - 1
    lens = np.zeros(n, dtype=np.uint64)
    st = np.zeros(n, dtype=np.int32)
    total = lib().zo_uncompress_batch(base.ctypes.data, offsets.ctypes.data, n, dataFormat, threads,
                                      lens.ctypes.data, st.ctypes.data)
    return total, lens, st
