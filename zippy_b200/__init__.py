"""zippy_b200 -- host-side mirror of guzba/zippy's public API over the B200 C ABI.

Same names, argument meaning and error behaviour as the reference module
(src/zippy.nim:11-177, src/zippy/common.nim:1-12, crc.nim:53-75, adler32.nim:6-66):

    compress(src, level=DefaultCompression, dataFormat=dfGzip) -> bytes
    uncompress(src, dataFormat=dfDetect) -> bytes
    crc32(src) / adler32(src) -> int
    ZippyError, dfDetect/dfZlib/dfGzip/dfDeflate, NoCompression/BestSpeed/...

plus the batch forms that make a GPU worthwhile (no reference counterpart).  All codec
work happens in libzippy_b200.so's CUDA kernels; this module only owns buffers, draws the
reference's random gzip FNAME length (zippy.nim:28-42) and maps status codes to ZippyError.
There is no CPU fallback: without the library or a CUDA device every call raises.
"""
import ctypes
import os
import threading

import numpy as np

from . import _native

dfDetect, dfZlib, dfGzip, dfDeflate = 0, 1, 2, 3                       # common.nim:4-5
NoCompression, BestSpeed, BestCompression = 0, 1, 9                    # common.nim:7-12
DefaultCompression, HuffmanOnly = -1, -2

__all__ = ["compress", "uncompress", "crc32", "adler32", "deflate", "inflate", "compress_batch", "uncompress_batch",
           "uncompressed_sizes", "checksum_batch", "ZippyError", "Context", "MultiGpu", "dfDetect", "dfZlib", "dfGzip",
           "dfDeflate", "NoCompression", "BestSpeed", "BestCompression", "DefaultCompression", "HuffmanOnly"]


class ZippyError(Exception):
    """Raised if an operation fails (common.nim:2).  .code is the ZB200_* status."""

    def __init__(self, code, msg=None):
        self.code = code
        super().__init__(msg or _native.lib().zb200_strerror(code).decode())


def _check(ctx, rc):
    if rc != 0:
        extra = ""
        if rc in (20, 21) and ctx is not None:
            extra = " [" + _native.lib().zb200_last_cuda_error(ctx).decode() + "]"
        raise ZippyError(rc, _native.lib().zb200_strerror(rc).decode() + extra)


def _as_u8(buf):
    if isinstance(buf, np.ndarray):
        return np.ascontiguousarray(buf, dtype=np.uint8).reshape(-1)
    return np.frombuffer(bytes(buf) if not isinstance(buf, (bytes, bytearray, memoryview)) else buf, dtype=np.uint8)


def _pack(items):
    """list of bytes-like -> (base uint8 array, offsets uint64[n+1])"""
    lens = np.fromiter((len(x) for x in items), dtype=np.uint64, count=len(items))
    offs = np.zeros(len(items) + 1, dtype=np.uint64)
    np.cumsum(lens, out=offs[1:])
    base = np.frombuffer(b"".join(bytes(x) for x in items), dtype=np.uint8) if len(items) else np.zeros(0, np.uint8)
    return base, offs


class Context:
    """One zb200_ctx: a CUDA device + stream + cached device scratch."""

    def __init__(self, device=-1):
        self._h = ctypes.c_void_p()
        L = _native.lib()
        rc = L.zb200_init(device, ctypes.byref(self._h))
        if rc != 0:
            raise ZippyError(rc, "zb200_init failed: %s (zippy_b200 needs a CUDA device; there is no CPU fallback)"
                             % L.zb200_strerror(rc).decode())

    LEGACY_DEFAULT_STREAM = 1   # cudaStreamLegacy: the handle that names the default stream explicitly

    def set_stream(self, cuda_stream):
        """Run on a caller-owned cudaStream_t (int handle).  0 / None restores the ctx's own (non-blocking)
        stream, which is NOT ordered against work the caller queued elsewhere: a caller that fills device
        buffers on its default stream and wants ordering without synchronising passes LEGACY_DEFAULT_STREAM
        (torch: `ctx.set_stream(torch.cuda.current_stream().cuda_stream or ctx.LEGACY_DEFAULT_STREAM)`)."""
        _check(self._h, _native.lib().zb200_set_stream(self._h, ctypes.c_void_p(cuda_stream or 0)))

    def close(self):
        if self._h:
            _native.lib().zb200_shutdown(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- batches over host buffers -------------------------------------------------
    def compress_batch(self, base, offsets, level=DefaultCompression, dataFormat=dfGzip, fname_lens=None):
        """-> (out uint8 array, out_offsets uint64[n+1]).  fname_lens: per-input gzip FNAME letters (0..25)."""
        L = _native.lib()
        base = _as_u8(base)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        bound = sum(L.zb200_compress_bound(int(offsets[i + 1] - offsets[i]), dataFormat) for i in range(n)) \
            if n <= 4096 else int(L.zb200_compress_bound(int(offsets[-1] - offsets[0]), dataFormat)) + 64 * n
        out = np.empty(int(bound) + 64, dtype=np.uint8)
        out_offs = np.zeros(n + 1, dtype=np.uint64)
        st = np.zeros(max(n, 1), dtype=np.int32)
        fl = None
        if fname_lens is not None:
            fl = np.ascontiguousarray(fname_lens, dtype=np.uint8)
        rc = L.zb200_compress_batch(self._h, base.ctypes.data, offsets.ctypes.data, n, level, dataFormat,
                                    fl.ctypes.data if fl is not None else None, out.ctypes.data, out.size,
                                    out_offs.ctypes.data, st.ctypes.data)
        _check(self._h, rc)
        return out[:int(out_offs[n])], out_offs

    def uncompressed_sizes(self, base, offsets, dataFormat=dfDetect):
        L = _native.lib()
        base = _as_u8(base)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        sizes = np.zeros(max(n, 1), dtype=np.uint64)
        st = np.zeros(max(n, 1), dtype=np.int32)
        _check(self._h, L.zb200_uncompress_sizes(self._h, base.ctypes.data, offsets.ctypes.data, n, dataFormat,
                                                  sizes.ctypes.data, st.ctypes.data))
        return sizes[:n], st[:n]

    def uncompress_batch(self, base, offsets, dataFormat=dfDetect, max_total=None, sizes=None):
        """-> (out uint8 array, out_offsets uint64[n+1], out_lens uint64[n], statuses int32[n]).
        `sizes`: uncompressed sizes known to the caller (a container's directory); skips the sizing pass."""
        L = _native.lib()
        base = _as_u8(base)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        if sizes is None:
            sizes, st0 = self.uncompressed_sizes(base, offsets, dataFormat)
        else:
            sizes, st0 = np.ascontiguousarray(sizes, dtype=np.uint64), np.zeros(n, dtype=np.int32)
        sizes = np.where(st0 == 0, sizes, 0).astype(np.uint64)
        # a gzip ISIZE is a claim, not a fact: DEFLATE cannot expand more than 1032:1, so a
        # larger claim can only end in a size/checksum failure -- never allocate for it.
        comp_lens = offsets[1:] - offsets[:-1]
        sizes = np.minimum(sizes, comp_lens * np.uint64(1032) + np.uint64(1024))
        if max_total is not None and int(sizes.sum()) > max_total:
            raise ZippyError(19)
        dst_offs = np.zeros(n + 1, dtype=np.uint64)
        np.cumsum(sizes, out=dst_offs[1:])
        out = np.empty(int(dst_offs[n]) + 64, dtype=np.uint8)
        lens = np.zeros(max(n, 1), dtype=np.uint64)
        st = np.zeros(max(n, 1), dtype=np.int32)
        _check(self._h, L.zb200_uncompress_batch(self._h, base.ctypes.data, offsets.ctypes.data, n, dataFormat,
                                                  out.ctypes.data, dst_offs.ctypes.data, lens.ctypes.data,
                                                  st.ctypes.data))
        st = np.where(st0 != 0, st0, st[:n]).astype(np.int32)
        out = out[:int(dst_offs[n])]
        lens = lens[:n]
        small = np.nonzero(st == 19)[0]
        if len(small):
            # a size claim (gzip ISIZE, a container's directory) understated the content: the reference inflates
            # anyway and lets its checksum / size checks decide (gzip.nim:80-88) -- redo those members one by one
            extra = []
            end = int(dst_offs[n])
            dst_offs = dst_offs.copy()
            for i in small:
                try:
                    b = self.decode_one(base[int(offsets[i]):int(offsets[i + 1])], dataFormat)
                    st[i] = 0
                    dst_offs[i] = end          # appended behind the slots; callers use out[off : off + len]
                    lens[i] = len(b)
                    end += len(b)
                    extra.append(np.frombuffer(b, dtype=np.uint8))
                except ZippyError as e:
                    st[i] = e.code
            if extra:
                out = np.concatenate([out] + extra)
        return out, dst_offs, lens, st

    def checksum_batch(self, base, offsets, kind="crc32"):
        L = _native.lib()
        base = _as_u8(base)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        out = np.zeros(max(n, 1), dtype=np.uint32)
        _check(self._h, L.zb200_checksum_batch(self._h, base.ctypes.data, offsets.ctypes.data, n,
                                                0 if kind == "crc32" else 1, out.ctypes.data))
        return out[:n]

    # ---- device-resident batches (raw device pointers; e.g. torch tensor .data_ptr()) ----
    def compress_batch_device(self, d_src, offsets, level, dataFormat, d_dst, dst_cap, fname_lens=None):
        L = _native.lib()
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        out_offs = np.zeros(n + 1, dtype=np.uint64)
        fl = np.ascontiguousarray(fname_lens, dtype=np.uint8) if fname_lens is not None else None
        _check(self._h, L.zb200_compress_batch_device(self._h, d_src, offsets.ctypes.data, n, level, dataFormat,
                                                       fl.ctypes.data if fl is not None else None, d_dst, dst_cap,
                                                       out_offs.ctypes.data, None))
        return out_offs

    def compress_batch_h2d(self, h_src, offsets, level, dataFormat, d_dst, dst_cap, fname_lens=None):
        """Host inputs (raw pointer; page-locked memory lets the copies overlap the kernels) -> members left
        in device memory at d_dst.  Returns the member offsets (uint64[n+1])."""
        L = _native.lib()
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        out_offs = np.zeros(n + 1, dtype=np.uint64)
        fl = np.ascontiguousarray(fname_lens, dtype=np.uint8) if fname_lens is not None else None
        _check(self._h, L.zb200_compress_batch_h2d(self._h, h_src, offsets.ctypes.data, n, level, dataFormat,
                                                    fl.ctypes.data if fl is not None else None, d_dst, dst_cap,
                                                    out_offs.ctypes.data, None))
        return out_offs

    def download(self, d_src, h_dst, nbytes):
        """Device -> host copy on the ctx stream; returns when the bytes have landed."""
        _check(self._h, _native.lib().zb200_download(self._h, d_src, h_dst, nbytes))

    def uncompress_batch_device(self, d_src, offsets, dataFormat, d_dst, dst_offsets):
        L = _native.lib()
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        dst_offsets = np.ascontiguousarray(dst_offsets, dtype=np.uint64)
        n = len(offsets) - 1
        lens = np.zeros(max(n, 1), dtype=np.uint64)
        st = np.zeros(max(n, 1), dtype=np.int32)
        _check(self._h, L.zb200_uncompress_batch_device(self._h, d_src, offsets.ctypes.data, n, dataFormat, d_dst,
                                                         dst_offsets.ctypes.data, lens.ctypes.data, st.ctypes.data))
        return lens[:n], st[:n]

    def uncompressed_sizes_device(self, d_src, offsets, dataFormat):
        L = _native.lib()
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        sizes = np.zeros(max(n, 1), dtype=np.uint64)
        st = np.zeros(max(n, 1), dtype=np.int32)
        _check(self._h, L.zb200_uncompress_sizes_device(self._h, d_src, offsets.ctypes.data, n, dataFormat,
                                                         sizes.ctypes.data, st.ctypes.data))
        return sizes[:n], st[:n]

    def checksum_batch_device(self, d_src, offsets, kind="crc32"):
        L = _native.lib()
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        out = np.zeros(max(n, 1), dtype=np.uint32)
        _check(self._h, L.zb200_checksum_batch_device(self._h, d_src, offsets.ctypes.data, n,
                                                       0 if kind == "crc32" else 1, out.ctypes.data))
        return out[:n]

    def timing(self):
        t = _native.Timing()
        _check(self._h, _native.lib().zb200_last_timing(self._h, ctypes.byref(t)))
        return {f: getattr(t, f) for f, _ in t._fields_}

    # ---- the single-input seam (deflate.nim:207, inflate.nim:268, crc.nim:53, adler32.nim:6) ----
    def deflate(self, src, level=DefaultCompression):
        L = _native.lib()
        src = _as_u8(src)
        cap = L.zb200_deflate_bound(src.size)
        out = np.empty(cap + 8, dtype=np.uint8)
        n = ctypes.c_size_t(0)
        _check(self._h, L.zb200_deflate(self._h, src.ctypes.data, src.size, level, out.ctypes.data, out.size,
                                        ctypes.byref(n)))
        return out[:n.value].tobytes()

    def decode_one(self, src, dataFormat=dfDetect, pos=0):
        """One input of unknown size, decoded once: zb200_decode_begin (inflate + trailer check into device
        memory, size reported) then zb200_decode_finish (copy out)."""
        L = _native.lib()
        src = _as_u8(src)
        n = ctypes.c_size_t(0)
        _check(self._h, L.zb200_decode_begin(self._h, src.ctypes.data, src.size, dataFormat, pos, ctypes.byref(n)))
        out = np.empty(n.value + 8, dtype=np.uint8)
        m = ctypes.c_size_t(0)
        _check(self._h, L.zb200_decode_finish(self._h, out.ctypes.data, n.value, ctypes.byref(m)))
        return out[:m.value].tobytes()

    def inflate(self, src, pos=0):
        return self.decode_one(src, dfDeflate, pos)

    def crc32(self, src):
        src = _as_u8(src)
        v = ctypes.c_uint32(0)
        _check(self._h, _native.lib().zb200_crc32(self._h, src.ctypes.data, src.size, ctypes.byref(v)))
        return v.value

    def adler32(self, src):
        src = _as_u8(src)
        v = ctypes.c_uint32(0)
        _check(self._h, _native.lib().zb200_adler32(self._h, src.ctypes.data, src.size, ctypes.byref(v)))
        return v.value


class MultiGpu:
    """zb200_mgpu: several devices behind one call (one host thread and one ctx per device)."""

    def __init__(self, devices=None):
        self._h = ctypes.c_void_p()
        L = _native.lib()
        arr = (ctypes.c_int * len(devices))(*devices) if devices else None
        rc = L.zb200_mgpu_init(arr, len(devices) if devices else 0, ctypes.byref(self._h))
        if rc != 0:
            raise ZippyError(rc, "zb200_mgpu_init failed: %s" % L.zb200_strerror(rc).decode())

    def close(self):
        if self._h:
            _native.lib().zb200_mgpu_shutdown(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def device_count(self):
        return _native.lib().zb200_mgpu_device_count(self._h)

    def compress_batch(self, base, offsets, level=DefaultCompression, dataFormat=dfGzip):
        """-> (one concatenated uint8 stream, global member offsets uint64[n+1])"""
        L = _native.lib()
        base = _as_u8(base)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        bound = int(L.zb200_compress_bound(int(offsets[-1] - offsets[0]), dataFormat)) + 128 * n + 4096
        out = np.empty(bound, dtype=np.uint8)
        oo = np.zeros(n + 1, dtype=np.uint64)
        _check(None, L.zb200_mgpu_compress_batch(self._h, base.ctypes.data, offsets.ctypes.data, n, level, dataFormat, None,
                                                  out.ctypes.data, out.size, oo.ctypes.data, None))
        return out[:int(oo[n])], oo

    def uncompress_batch(self, base, offsets, sizes, dataFormat=dfDetect):
        """sizes: output slot sizes (uint64[n]).  -> (out, dst_offsets, lens, statuses)"""
        L = _native.lib()
        base = _as_u8(base)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        do = np.zeros(n + 1, dtype=np.uint64)
        np.cumsum(np.asarray(sizes, dtype=np.uint64), out=do[1:])
        out = np.empty(int(do[n]) + 64, dtype=np.uint8)
        lens = np.zeros(max(n, 1), dtype=np.uint64)
        st = np.zeros(max(n, 1), dtype=np.int32)
        _check(None, L.zb200_mgpu_uncompress_batch(self._h, base.ctypes.data, offsets.ctypes.data, n, dataFormat,
                                                    out.ctypes.data, do.ctypes.data, lens.ctypes.data, st.ctypes.data))
        return out[:int(do[n])], do, lens[:n], st[:n]

    def checksum_batch(self, base, offsets, kind="crc32"):
        L = _native.lib()
        base = _as_u8(base)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        out = np.zeros(max(n, 1), dtype=np.uint32)
        _check(None, L.zb200_mgpu_checksum_batch(self._h, base.ctypes.data, offsets.ctypes.data, n,
                                                  0 if kind == "crc32" else 1, out.ctypes.data))
        return out[:n]


def host_register(ptr, nbytes):
    """Page-lock a caller-owned host range (cudaHostRegister) so host-buffer calls overlap their copies."""
    _check(None, _native.lib().zb200_host_register(ptr, nbytes))


def host_unregister(ptr):
    _check(None, _native.lib().zb200_host_unregister(ptr))


_default = None
_default_lock = threading.Lock()


def default_context():
    global _default
    with _default_lock:
        if _default is None:
            _default = Context(-1)
        return _default


# ---- the reference's public procs ------------------------------------------------------
def compress(src, level=DefaultCompression, dataFormat=dfGzip):
    """zippy.compress (zippy.nim:11-98)."""
    if level < -2 or level > 9:
        raise ZippyError(1, "Invalid compression level %d" % level)          # deflate.nim:208-209
    if dataFormat not in (dfGzip, dfZlib, dfDeflate):
        raise ZippyError(2, "Invalid data format dfDetect")                  # zippy.nim:83-84
    fl = None
    if dataFormat == dfGzip:
        fl = [os.urandom(1)[0] % 26]                                         # zippy.nim:28-42
    base, offs = _pack([src])
    out, _ = default_context().compress_batch(base, offs, level, dataFormat, fl)
    return out.tobytes()


def uncompress(src, dataFormat=dfDetect):
    """zippy.uncompress (zippy.nim:100-177)."""
    if dataFormat not in (dfDetect, dfZlib, dfGzip, dfDeflate):
        raise ZippyError(2)
    return default_context().decode_one(src, dataFormat)


def crc32(src):
    return default_context().crc32(src)


def adler32(src):
    return default_context().adler32(src)


def deflate(src, level=DefaultCompression):
    return default_context().deflate(src, level)


def inflate(src, pos=0):
    return default_context().inflate(src, pos)


def compress_batch(items, level=DefaultCompression, dataFormat=dfGzip, fname_lens=None):
    """list of bytes -> list of bytes (one zippy.compress per item, one GPU launch sequence)."""
    base, offs = _pack(items)
    out, oo = default_context().compress_batch(base, offs, level, dataFormat, fname_lens)
    return [out[int(oo[i]):int(oo[i + 1])].tobytes() for i in range(len(items))]


def uncompress_batch(items, dataFormat=dfDetect):
    """list of bytes -> list of (bytes | ZippyError)."""
    base, offs = _pack(items)
    out, do, lens, st = default_context().uncompress_batch(base, offs, dataFormat)
    res = []
    for i in range(len(items)):
        res.append(ZippyError(int(st[i])) if st[i] != 0 else out[int(do[i]):int(do[i]) + int(lens[i])].tobytes())
    return res


def uncompressed_sizes(items, dataFormat=dfDetect):
    base, offs = _pack(items)
    return default_context().uncompressed_sizes(base, offs, dataFormat)


def checksum_batch(items, kind="crc32"):
    base, offs = _pack(items)
    return default_context().checksum_batch(base, offs, kind)
