"""Tarballs (a "next" row): host-side mirror of src/zippy/tarballs.nim:25-141.

`extract_all(tar_path, dest)` reads a .tar or .tar.gz: a gzip tarball is ONE gzip member, inflated
by the GPU path (uncompressGzip, tarballs.nim:50; a single member is decoded by one 8-lane group,
so this is a convenience, not a fast path), then the 512-byte header walk stays on the host exactly
as in the reference: ustar prefix, GNU 'L' long names, files / directories / symlinks, pax and
vendor records skipped, anything else is an error; paths are checked before anything is written
and nothing is left behind on failure."""
import os
import shutil

from . import ZippyError, dfGzip, uncompress

_ERR = 3


def _fail(msg):
    raise ZippyError(_ERR, msg)


def _oct(field):
    """tarballs.nim:5-23: the first run of ASCII digits in the field, base 8 (0 if none)."""
    i = 0
    while i < len(field) and not (48 <= field[i] <= 57):
        i += 1
    j = i
    while j < len(field) and 48 <= field[j] <= 57:
        j += 1
    if j == i:
        return 0
    try:
        return int(bytes(field[i:j]), 8)
    except ValueError as e:
        _fail(str(e))


def _cstr(field):
    k = bytes(field).find(b"\0")
    return bytes(field if k < 0 else field[:k]).decode("utf-8", "surrogateescape")


def _safe(path):
    if path.startswith("/") or path.startswith("\\") or (len(path) > 1 and path[1] == ":"):
        _fail("Absolute path not allowed " + path)
    if ".." in path.replace("\\", "/").split("/"):
        _fail("Path ../ not allowed " + path)


def read_tarball(data, gunzip=None):
    """-> list of (kind, path, payload | linkname, mode, mtime); kind in 'file', 'dir', 'symlink'."""
    data = bytes(data)
    if len(data) < 2:
        _fail("Invalid buffer, unable to uncompress")
    if data[0] == 31 and data[1] == 139:
        data = (gunzip or (lambda b: uncompress(b, dfGzip)))(data)
    out, pos, long_name = [], 0, ""
    while pos < len(data):
        if pos + 512 > len(data):
            _fail("Attempted to read past end of file, corrupted tarball?")
        h = data[pos:pos + 512]
        name, mode, size, mtime = _cstr(h[0:100]), _oct(h[100:107]), _oct(h[124:135]), _oct(h[136:147])
        typeflag, linkname = chr(h[156]), _cstr(h[157:257])
        prefix = _cstr(h[345:500]) if _cstr(h[257:263]) == "ustar" else ""
        pos += 512
        if pos + size > len(data):
            _fail("Attempted to read past end of file, corrupted tarball?")
        if name or long_name:
            if long_name:
                path, long_name = long_name, ""
            else:
                path = os.path.join(prefix, name) if prefix else name
            _safe(path)
            if typeflag in ("0", "\0"):
                out.append(("file", path, data[pos:pos + size], mode, mtime))
            elif typeflag == "5":
                out.append(("dir", path, b"", mode, mtime))
            elif typeflag == "2":
                out.append(("symlink", path, linkname, mode, mtime))
            elif typeflag == "L":
                long_name = _cstr(data[pos:pos + size])
            elif typeflag in ("g", "x") or "A" <= typeflag <= "Z":
                pass
            else:
                _fail("Unsupported header type " + typeflag)
        pos += (size + 511) & ~511
    return out


def extract_all(tar_path, dest, gunzip=None):
    if dest == "" or os.path.isdir(dest):
        _fail("Destination " + dest + " already exists")
    head = os.path.dirname(dest.rstrip("/\\"))
    if head and not os.path.isdir(head):
        _fail("Path to " + dest + " does not exist")
    with open(tar_path, "rb") as f:
        entries = read_tarball(f.read(), gunzip)
    try:
        times = []
        for kind, path, payload, mode, mtime in entries:
            target = os.path.join(dest, path)
            if kind == "file":
                os.makedirs(os.path.dirname(target) or dest, exist_ok=True)
                with open(target, "wb") as f:
                    f.write(payload)
                if mode:
                    os.chmod(target, mode & 0o777)
                times.append((target, mtime))
            elif kind == "dir":
                os.makedirs(target, exist_ok=True)
                times.append((target, mtime))
            else:
                os.makedirs(os.path.dirname(target) or dest, exist_ok=True)
                os.symlink(payload, target)
        for target, mtime in times:  # second pass: directories would be touched by their files
            if mtime > 0:
                os.utime(target, (mtime, mtime))
    except Exception:
        shutil.rmtree(dest, ignore_errors=True)
        raise
