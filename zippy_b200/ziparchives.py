"""ZIP archives over the batch entry points (a "next" row).

Host-side mirror of the reference's src/zippy/ziparchives.nim: `createZipArchive` (:458-634) is a
loop of crc32 + compress(BestSpeed, dfDeflate) over the entries and `extractFile` / `extractAll`
(:37-93, :398-452) a loop of uncompress(dfDeflate) + crc32 -- here each loop is ONE batched GPU
call (zb200_compress_batch / zb200_uncompress_batch / zb200_checksum_batch).  The container
format itself (local headers, central directory, ZIP64 records, CP437 names) stays on the host,
as in the reference.  Same checks and error messages as the reference; names follow its API:

    open_zip_archive(path | bytes) -> ZipArchiveReader   (openZipArchive, :183)
    reader.walk_files()                                   (walkFiles, :31)
    reader.extract_file(name) -> bytes                    (extractFile, :37)
    reader.extract_files(names=None) -> {name: bytes}     (batched; no reference counterpart)
    extract_all(zip_path, dest)                           (extractAll, :398)
    create_zip_archive({name: bytes}) -> bytes            (createZipArchive, :622-634)
"""
import os
import stat
import struct
import time

import numpy as np

from . import BestSpeed, ZippyError, default_context, dfDeflate

_LOCAL = 0x04034B50
_CENTRAL = 0x02014B50
_EOCD = 0x06054B50
_EOCD64 = 0x06064B50
_LOC64 = 0x07064B50
_ERR_ARCHIVE = 3  # ZB200_ERR_UNCOMPRESS: the generic "invalid buffer" class


def _fail(msg):
    raise ZippyError(_ERR_ARCHIVE, msg)


def _eof():
    _fail("Attempted to read past end of file, corrupted archive?")


def _safe_path(path):
    """internal.nim verifyPathIsSafeToExtract: no absolute paths, no drive letters, no '..' parts."""
    if path.startswith("/") or path.startswith("\\") or (len(path) > 1 and path[1] == ":"):
        _fail("Absolute path not allowed " + path)
    if ".." in path.replace("\\", "/").split("/"):
        _fail("Path ../ not allowed " + path)


def _dos_time(t=None):
    lt = time.localtime(t)
    return ((lt.tm_hour << 11) | (lt.tm_min << 5) | (lt.tm_sec // 2),
            ((max(0, lt.tm_year - 1980)) << 9) | (lt.tm_mon << 5) | lt.tm_mday)


def _from_dos_time(tm, dt):
    sec, mi, ho = (tm & 31) * 2, (tm >> 5) & 63, (tm >> 11) & 31
    day, mon, yr = dt & 31, (dt >> 5) & 15, ((dt >> 9) & 127) + 1980
    if sec <= 59 and mi <= 59 and ho <= 23 and 1 <= mon <= 12 and 1 <= day <= 31:
        try:
            return time.mktime((yr, mon, day, ho, mi, sec, 0, 0, -1))
        except (OverflowError, ValueError):
            return None
    return None


class _Record:
    __slots__ = ("is_dir", "header_offset", "path", "crc", "csize", "usize", "mode")


class ZipArchiveReader:
    """ziparchives.nim:27-29 (records keep the central directory's order)."""

    def __init__(self, data, ctx=None):
        self._d = data if isinstance(data, (bytes, bytearray, memoryview)) else bytes(data)
        self._ctx = ctx
        self.records = {}
        self._parse()

    # ---- central directory (ziparchives.nim:183-395) ----
    def _u16(self, p):
        return struct.unpack_from("<H", self._d, p)[0]

    def _u32(self, p):
        return struct.unpack_from("<I", self._d, p)[0]

    def _u64(self, p):
        return struct.unpack_from("<Q", self._d, p)[0]

    def _parse(self):
        d, size = self._d, len(self._d)
        eocd = bytes(d).rfind(struct.pack("<I", _EOCD), 0, max(0, size - 22) + 4)
        if eocd < 0 or eocd + 22 > size:
            _eof()
        zip64 = eocd - 20 >= 0 and self._u32(eocd - 20) == _LOC64
        if zip64:
            if self._u32(eocd - 16) != 0:
                _fail("Unsupported archive, disk number")
            pos = self._u64(eocd - 12)
            if self._u32(eocd - 4) != 1:
                _fail("Unsupported archive, num disks")
            if pos + 64 > size:
                _eof()
            if self._u32(pos) != _EOCD64:
                _fail("Invalid central directory file header")
            disk, start_disk = self._u32(pos + 16), self._u32(pos + 20)
            n_disk, n_total = self._u64(pos + 24), self._u64(pos + 32)
            cd_size, cd_start = self._u64(pos + 40), self._u64(pos + 48)
        else:
            disk, start_disk, n_disk, n_total = (self._u16(eocd + 4), self._u16(eocd + 6), self._u16(eocd + 8),
                                                 self._u16(eocd + 10))
            cd_size, cd_start = self._u32(eocd + 12), self._u32(eocd + 16)
        if disk != 0:
            _fail("Unsupported archive, disk number")
        if start_disk != 0:
            _fail("Unsupported archive, start disk")
        if n_disk != n_total:
            _fail("Unsupported archive, record number")
        # an archive may be appended to another file (an .exe, a .jpg): locate the directory from
        # the end by counting its records backwards, fall back to the recorded offset
        socd, p = cd_start, eocd
        sig = struct.pack("<I", _CENTRAL)
        raw = bytes(d)
        for k in range(n_total):
            p = raw.rfind(sig, 0, p + 3)   # the last record header that starts before p
            if p < 0:
                break
            if k == n_total - 1:
                socd = p
        shift = socd - cd_start
        pos = socd
        for _ in range(n_total):
            if pos + 46 > size:
                _eof()
            if self._u32(pos) != _CENTRAL:
                _fail("Invalid central directory file header")
            flags, method = self._u16(pos + 8), self._u16(pos + 10)
            crc, csize, usize = self._u32(pos + 16), self._u32(pos + 20), self._u32(pos + 24)
            nlen, xlen, clen = self._u16(pos + 28), self._u16(pos + 30), self._u16(pos + 32)
            fdisk, xattr, hoff = self._u16(pos + 34), self._u32(pos + 38), self._u32(pos + 42)
            if method not in (0, 8):
                _fail("Unsupported archive, compression method")
            if fdisk != 0:
                _fail("Invalid file disk number")
            pos += 46
            if pos + nlen > size:
                _eof()
            name_b = bytes(d[pos:pos + nlen])
            pos += nlen
            q, end = pos, pos + xlen
            while q + 4 <= end:  # ZIP64 extended information (id 1): only the saturated fields are present
                fid, flen = self._u16(q), self._u16(q + 2)
                q += 4
                if fid == 1:
                    z, zend = q, q + flen
                    for field in ("usize", "csize", "hoff"):
                        cur = {"usize": usize, "csize": csize, "hoff": hoff}[field]
                        if cur == 0xFFFFFFFF:
                            if z + 8 > zend or z + 8 > size:
                                _eof()
                            v = self._u64(z)
                            z += 8
                            if field == "usize":
                                usize = v
                            elif field == "csize":
                                csize = v
                            else:
                                hoff = v
                    break
                q += flen
            pos += xlen + clen
            if pos > socd + cd_size:
                _fail("Invalid central directory size")
            if flags & 0x800:  # language encoding flag: UTF-8
                name = name_b.decode("utf-8", "replace")
            else:
                try:
                    name = name_b.decode("utf-8")
                except UnicodeDecodeError:
                    name = name_b.decode("cp437")  # DOS / OEM names
            if name in self.records:
                _fail("Unsupported archive, duplicate entry")
            r = _Record()
            r.is_dir = bool(xattr & 0x10) or bool((xattr >> 16) & stat.S_IFDIR) or name.endswith("/")
            r.header_offset = hoff + shift
            r.path, r.crc, r.csize, r.usize = name, crc, csize, usize
            r.mode = (xattr >> 16) & 0o777
            self.records[name] = r

    # ---- access ----
    def walk_files(self):
        for r in self.records.values():
            if not r.is_dir:
                yield r.path

    def _payload(self, r):
        """(method, start) of an entry's data; checks of extractFile (ziparchives.nim:52-78)."""
        pos, size = r.header_offset, len(self._d)
        if pos + 30 > size:
            _eof()
        if self._u32(pos) != _LOCAL:
            _fail("Invalid file header")
        method = self._u16(pos + 8)
        pos += 30 + self._u16(pos + 26) + self._u16(pos + 28)
        if pos + r.csize > size:
            _eof()
        if method not in (0, 8):
            _fail("Unsupported archive, compression method")
        return method, pos

    def extract_files(self, names=None):
        """All requested entries in one batched inflate + one batched CRC-32 on the GPU."""
        ctx = self._ctx or default_context()
        names = list(self.walk_files()) if names is None else list(names)
        recs = []
        for nme in names:
            r = self.records.get(nme)
            if r is None or r.is_dir:
                _fail("No file record found for " + nme)
            recs.append(r)
        out = {}
        packed, offs, sizes, which = [], [0], [], []
        for r in recs:
            method, pos = self._payload(r)
            if method == 0:
                out[r.path] = bytes(self._d[pos:pos + r.csize])
            else:
                packed.append(bytes(self._d[pos:pos + r.csize]))
                offs.append(offs[-1] + r.csize)
                sizes.append(r.usize)
                which.append(r)
        if which:
            base = np.frombuffer(b"".join(packed), dtype=np.uint8) if offs[-1] else np.zeros(1, dtype=np.uint8)
            data, do, lens, st = ctx.uncompress_batch(base, np.array(offs, dtype=np.uint64), dfDeflate,
                                                      sizes=np.array(sizes, dtype=np.uint64))
            for i, r in enumerate(which):
                if st[i] != 0:
                    raise ZippyError(int(st[i]))
                out[r.path] = data[int(do[i]):int(do[i]) + int(lens[i])].tobytes()
        # crc32 of every extracted file against the directory (ziparchives.nim:92-93), one batch
        blobs = [out[r.path] for r in recs]
        if blobs:
            o2 = np.zeros(len(blobs) + 1, dtype=np.uint64)
            o2[1:] = np.cumsum([len(b) for b in blobs])
            joined = np.frombuffer(b"".join(blobs), dtype=np.uint8) if o2[-1] else np.zeros(1, dtype=np.uint8)
            crcs = ctx.checksum_batch(joined, o2, "crc32")
            for r, c in zip(recs, crcs):
                if int(c) != r.crc:
                    _fail("Verifying crc32 failed")
        return out

    def extract_file(self, path):
        return self.extract_files([path])[path]

    def close(self):
        self._d = b""


def open_zip_archive(src, ctx=None):
    if isinstance(src, (bytes, bytearray, memoryview)):
        return ZipArchiveReader(src, ctx)
    with open(src, "rb") as f:
        return ZipArchiveReader(f.read(), ctx)


def extract_all(zip_path, dest, ctx=None):
    """ziparchives.nim:398-452: dest must not exist, its parent must; nothing is left behind on failure."""
    if dest == "" or os.path.isdir(dest):
        _fail("Destination " + dest + " already exists")
    head = os.path.dirname(dest.rstrip("/\\"))
    if head and not os.path.isdir(head):
        _fail("Path to " + dest + " does not exist")
    reader = open_zip_archive(zip_path, ctx)
    for r in reader.records.values():
        _safe_path(r.path)
    try:
        files = reader.extract_files()
        for r in reader.records.values():
            target = os.path.join(dest, r.path)
            if r.is_dir:
                os.makedirs(target, exist_ok=True)
            else:
                os.makedirs(os.path.dirname(target), exist_ok=True)
                with open(target, "wb") as f:
                    f.write(files[r.path])
                if r.mode:
                    os.chmod(target, r.mode)
        for r in reader.records.values():  # second pass: directories would be touched by their files
            tm, dt = struct.unpack_from("<HH", reader._d, r.header_offset + 10)
            t = _from_dos_time(tm, dt)
            if t is not None:
                os.utime(os.path.join(dest, r.path), (t, t))
    except Exception:
        import shutil
        shutil.rmtree(dest, ignore_errors=True)
        raise
    finally:
        reader.close()


def create_zip_archive(entries, ctx=None):
    """{name: bytes} -> archive bytes.  Layout of ziparchives.nim:458-620: version 45, UTF-8 flag,
    ZIP64 extra fields everywhere, entries written from the LAST key to the first (the reference pops
    keys off the end), empty files stored, everything else deflated at BestSpeed -- in one batch."""
    ctx = ctx or default_context()
    names = list(entries.keys())[::-1]
    for nme in names:
        if nme == "":
            _fail("Invalid empty file name")
        if nme[0] == "/":
            _fail("File paths must be relative")
        if len(nme.encode("utf-8")) > 0xFFFF:
            _fail("File name len > uint16.high")
    blobs = [bytes(entries[nme]) for nme in names]
    n = len(blobs)
    offs = np.zeros(n + 1, dtype=np.uint64)
    if n:
        offs[1:] = np.cumsum([len(b) for b in blobs])
    joined = np.frombuffer(b"".join(blobs), dtype=np.uint8) if n and offs[-1] else np.zeros(1, dtype=np.uint8)
    crcs = ctx.checksum_batch(joined, offs, "crc32") if n else []
    comp, co = (ctx.compress_batch(joined, offs, BestSpeed, dfDeflate) if n else (np.zeros(0, np.uint8), offs))
    tm, dt = _dos_time()
    out = bytearray()
    recs = []
    for i, nme in enumerate(names):
        nb = nme.encode("utf-8")
        ulen = len(blobs[i])
        data = b"" if ulen == 0 else comp[int(co[i]):int(co[i + 1])].tobytes()
        method = 0 if ulen == 0 else 8
        recs.append((nb, len(out), ulen, len(data), method, int(crcs[i])))
        out += struct.pack("<IHHHHHIIIHH", _LOCAL, 45, 1 << 11, method, tm, dt, int(crcs[i]), 0xFFFFFFFF, 0xFFFFFFFF,
                           len(nb), 20)
        out += nb + struct.pack("<HHQQ", 1, 16, ulen, len(data)) + data
    cd_start = len(out)
    for nb, hoff, ulen, clen, method, crc in recs:
        out += struct.pack("<IHHHHHHIIIHHHHHII", _CENTRAL, 45, 45, 1 << 11, method, tm, dt, crc, 0xFFFFFFFF, 0xFFFFFFFF,
                           len(nb), 28, 0, 0, 0, 0, 0xFFFFFFFF)
        out += nb + struct.pack("<HHQQQ", 1, 24, ulen, clen, hoff)
    cd_end = len(out)
    out += struct.pack("<IQHHIIQQQQ", _EOCD64, 44, 45, 45, 0, 0, len(recs), len(recs), cd_end - cd_start, cd_start)
    out += struct.pack("<IIQI", _LOC64, 0, cd_end, 1)
    out += struct.pack("<IHHHHIIH", _EOCD, 0, 0, 0xFFFF, 0xFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0)
    return bytes(out)
