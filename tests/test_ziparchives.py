"""ZIP archives over the batch entry points (zippy_b200/ziparchives.py).

not-gpu: the container logic (central directory, ZIP64 records, appended archives, name checks)
with a zlib-backed stand-in for the codec context -- test scaffolding only, it lives here.
gpu: the real context: the reference's fixtures extracted on the GPU byte-exact against the
manifest made with Python's zipfile, archives created on the GPU read back by zipfile."""
import hashlib
import io
import json
import os
import zipfile
import zlib

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ZDIR = os.path.join(HERE, "golden", "ziparchives")
MANIFEST = json.load(open(os.path.join(ZDIR, "zip_manifest.json")))


class ZlibCtx:
    """Stand-in with the three Context methods ziparchives.py calls (CPU tests only)."""

    def checksum_batch(self, base, offsets, kind="crc32"):
        b = bytes(base)
        return np.array([zlib.crc32(b[int(offsets[i]):int(offsets[i + 1])]) for i in range(len(offsets) - 1)],
                        dtype=np.uint32)

    def compress_batch(self, base, offsets, level, fmt, fname_lens=None):
        b = bytes(base)
        outs = []
        for i in range(len(offsets) - 1):
            c = zlib.compressobj(1, zlib.DEFLATED, -15)
            outs.append(c.compress(b[int(offsets[i]):int(offsets[i + 1])]) + c.flush())
        oo = np.zeros(len(outs) + 1, dtype=np.uint64)
        oo[1:] = np.cumsum([len(x) for x in outs])
        return np.frombuffer(b"".join(outs), dtype=np.uint8), oo

    def uncompress_batch(self, base, offsets, fmt, sizes=None):
        b = bytes(base)
        outs, st = [], []
        for i in range(len(offsets) - 1):
            try:
                outs.append(zlib.decompress(b[int(offsets[i]):int(offsets[i + 1])], -15))
                st.append(0)
            except zlib.error:
                outs.append(b"")
                st.append(3)
        do = np.zeros(len(outs) + 1, dtype=np.uint64)
        do[1:] = np.cumsum([len(x) for x in outs])
        return (np.frombuffer(b"".join(outs) or b"\0", dtype=np.uint8), do,
                np.array([len(x) for x in outs], dtype=np.uint64), np.array(st, dtype=np.int32))


def _za():
    pytest.importorskip("numpy")
    import zippy_b200.ziparchives as za
    return za


def _check_fixture(za, name, ctx):
    reader = za.open_zip_archive(os.path.join(ZDIR, name), ctx)
    want = [e for e in MANIFEST[name]["entries"] if not e["is_dir"]]
    assert list(reader.walk_files()) == [e["name"] for e in want]
    files = reader.extract_files()
    for e in want:
        assert len(files[e["name"]]) == e["len"], e["name"]
        assert hashlib.sha256(files[e["name"]]).hexdigest() == e["sha256"], e["name"]
    one = want[len(want) // 2]["name"]
    assert reader.extract_file(one) == files[one]
    reader.close()


def _roundtrip_create(za, ctx):
    entries = {"README.txt": b"Hello, World!", "dir/empty.bin": b"", "dir/sub/data.bin": bytes(range(256)) * 300,
               "café.txt": "naïve".encode("utf-8")}
    blob = za.create_zip_archive(entries, ctx)
    with zipfile.ZipFile(io.BytesIO(blob)) as zf:   # an independent reader accepts it and verifies the CRCs
        assert zf.testzip() is None
        assert [i.filename for i in zf.infolist()] == list(entries)[::-1]   # written last key first
        for k, v in entries.items():
            assert zf.read(k) == v
            assert zf.getinfo(k).compress_type == (zipfile.ZIP_STORED if not v else zipfile.ZIP_DEFLATED)
    back = za.open_zip_archive(blob, ctx).extract_files()
    assert back == entries


def test_zip_fixtures_container_logic_cpu():
    za = _za()
    for name in MANIFEST:
        _check_fixture(za, name, ZlibCtx())


def test_zip_create_and_read_back_cpu():
    _roundtrip_create(_za(), ZlibCtx())


def test_zip_errors_cpu(tmp_path):
    za = _za()
    from zippy_b200 import ZippyError
    ctx = ZlibCtx()
    for bad in ({"": b"x"}, {"/abs": b"x"}, {"n" * 70000: b"x"}):
        with pytest.raises(ZippyError):
            za.create_zip_archive(bad, ctx)
    with pytest.raises(ZippyError):
        za.open_zip_archive(b"not a zip archive at all", ctx)
    blob = bytearray(za.create_zip_archive({"a.txt": b"some text " * 50, "b.txt": b"other"}, ctx))
    r = za.open_zip_archive(bytes(blob), ctx)
    with pytest.raises(ZippyError):
        r.extract_file("missing.txt")
    # flip a payload byte of a.txt: inflate error or CRC mismatch, never silent
    hdr = r.records["a.txt"].header_offset
    blob[hdr + 30 + 5 + 20 + 8] ^= 0x55
    with pytest.raises(ZippyError):
        za.open_zip_archive(bytes(blob), ctx).extract_file("a.txt")
    # extract_all refuses an existing destination and unsafe paths
    z1 = tmp_path / "ok.zip"
    z1.write_bytes(za.create_zip_archive({"x/y.txt": b"1", "z.txt": b""}, ctx))
    dest = tmp_path / "out"
    za.extract_all(str(z1), str(dest), ctx)
    assert (dest / "x" / "y.txt").read_bytes() == b"1" and (dest / "z.txt").read_bytes() == b""
    with pytest.raises(ZippyError):
        za.extract_all(str(z1), str(dest), ctx)
    evil = io.BytesIO()
    with zipfile.ZipFile(evil, "w") as zf:
        zf.writestr("../escape.txt", b"x")
    z2 = tmp_path / "evil.zip"
    z2.write_bytes(evil.getvalue())
    with pytest.raises(ZippyError):
        za.extract_all(str(z2), str(tmp_path / "out2"), ctx)
    assert not (tmp_path / "out2").exists()


def test_zip_reads_zipfile_archives_cpu():
    """Archives written by another implementation: stored + deflated, non-ZIP64, with a comment."""
    za = _za()
    buf = io.BytesIO()
    with zipfile.ZipFile(buf, "w", zipfile.ZIP_DEFLATED) as zf:
        zf.comment = b"archive comment"
        zf.writestr("a/b.txt", b"hello " * 1000)
        zf.writestr(zipfile.ZipInfo("stored.bin"), b"\x00\x01\x02", compress_type=zipfile.ZIP_STORED)
        zf.writestr("a/", b"")
    r = za.open_zip_archive(buf.getvalue(), ZlibCtx())
    assert list(r.walk_files()) == ["a/b.txt", "stored.bin"]
    assert r.extract_file("a/b.txt") == b"hello " * 1000 and r.extract_file("stored.bin") == b"\x00\x01\x02"


def test_zip64_many_entries_cpu():
    """More than 65 535 entries: only the ZIP64 end-of-central-directory record can count them
    (ziparchives.nim:595-618 always writes it; the reader takes the counts from it, :203-236)."""
    za = _za()
    ctx = ZlibCtx()
    entries = {"d%03d/f%05d.txt" % (i % 200, i): (b"entry %d\n" % i) * (i % 3) for i in range(70000)}
    blob = za.create_zip_archive(entries, ctx)
    r = za.open_zip_archive(blob, ctx)
    assert len(r.records) == 70000
    names = list(r.walk_files())
    assert names == list(entries)[::-1]
    pick = names[::997]
    got = r.extract_files(pick)
    for k in pick:
        assert got[k] == entries[k]
    with zipfile.ZipFile(io.BytesIO(blob)) as zf:
        assert len(zf.infolist()) == 70000
        assert zf.read(names[12345]) == entries[names[12345]]


def _zipfile_made_archive(path):
    buf = io.BytesIO()
    with zipfile.ZipFile(buf, "w", zipfile.ZIP_DEFLATED) as zf:
        zf.comment = b"archive comment"
        zf.writestr("a/b.txt", b"hello " * 1000)
        zf.writestr(zipfile.ZipInfo("stored.bin"), b"\x00\x01\x02", compress_type=zipfile.ZIP_STORED)
        zf.writestr("a/", b"")
        zf.writestr("caf\u00e9.txt", "na\u00efve".encode("utf-8"))
    with open(path, "wb") as f:
        f.write(b"JUNK" * 100 + buf.getvalue())     # appended to another file: offsets are off by 400


def _run_cpp_zip(tmp_path, link_args):
    """include/zippy_b200_zip.hpp (the C++ form of ziparchives.nim) through tests/native/cpp_zip_test.cpp."""
    import subprocess
    root = os.path.dirname(HERE)
    exe = str(tmp_path / "cpp_zip_test")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-o", exe, os.path.join(HERE, "native", "cpp_zip_test.cpp")] + link_args)
    src, dst = str(tmp_path / "in.zip"), str(tmp_path / "out.zip")
    _zipfile_made_archive(src)
    out = subprocess.run([exe, src, dst], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and out.stdout.startswith("OK"), (out.stdout, out.stderr)
    with zipfile.ZipFile(dst) as zf:                  # an independent reader accepts what the C++ side wrote
        assert zf.testzip() is None
        assert [i.filename for i in zf.infolist()] == ["caf\u00e9.txt", "dir/sub/data.bin", "dir/empty.bin", "README.txt"]
        assert zf.read("README.txt") == b"Hello, World!" and len(zf.read("dir/sub/data.bin")) == 300 * 256
    return root


def test_cpp_zip_layer_container_logic_cpu(tmp_path):
    # zlib-backed stand-in for the C ABI (tests/native/mock_abi_zlib.cpp): container logic without a GPU
    _run_cpp_zip(tmp_path, [os.path.join(HERE, "native", "mock_abi_zlib.cpp"), "-lz"])


@pytest.mark.gpu
def test_cpp_zip_layer_gpu(tmp_path):
    libdir = os.path.join(os.path.dirname(HERE), "zippy_b200")
    _run_cpp_zip(tmp_path, ["-L" + libdir, "-l:libzippy_b200.so", "-Wl,-rpath," + libdir])


@pytest.mark.gpu
def test_zip_fixtures_gpu():
    import zippy_b200 as z
    za = _za()
    ctx = z.default_context()
    for name in MANIFEST:
        _check_fixture(za, name, ctx)


@pytest.mark.gpu
def test_zip_create_gpu_read_back_by_zipfile(tmp_path):
    import zippy_b200 as z
    from tests import util
    za = _za()
    ctx = z.default_context()
    _roundtrip_create(za, ctx)
    corpus = util.load_corpus()
    entries = {("corpus/%s" % k): v for k, v in corpus.items()}
    entries["corpus/empty"] = b""
    blob = za.create_zip_archive(entries, ctx)
    with zipfile.ZipFile(io.BytesIO(blob)) as zf:
        assert zf.testzip() is None
        for k, v in entries.items():
            assert zf.read(k) == v
    # and the reverse: a zipfile-made archive (level 9) extracted on the GPU, through extract_all
    buf = io.BytesIO()
    with zipfile.ZipFile(buf, "w", zipfile.ZIP_DEFLATED, compresslevel=9) as zf:
        for k, v in entries.items():
            zf.writestr(k, v)
    p = tmp_path / "in.zip"
    p.write_bytes(buf.getvalue())
    za.extract_all(str(p), str(tmp_path / "out"), ctx)
    for k, v in entries.items():
        assert (tmp_path / "out" / k).read_bytes() == v
