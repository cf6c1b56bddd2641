"""tarballs (zippy_b200/tarballs.py).  Gold = Python's tarfile, the role `tar -xf`
plays in the reference's tests/test_tarballs_read.nim.  not-gpu: the header walk with zlib as the
gunzip; gpu: the .tar.gz inflated by the GPU path."""
import io
import os
import tarfile
import zlib

import pytest


def _make(tmp_path, fmt, gz):
    src = tmp_path / "src"
    (src / "pkg" / "sub").mkdir(parents=True)
    (src / "pkg" / "a.txt").write_bytes(b"alpha\n" * 1000)
    (src / "pkg" / "sub" / "empty").write_bytes(b"")
    # names over 100 bytes: GNU 'L' records are followed (tarballs.nim:117-118); pax 'x' records are
    # skipped by the reference (:119-120), so pax / ustar archives here keep to short names
    long_name = ("long_" * 30 + "name.bin") if fmt == tarfile.GNU_FORMAT else "short_name.bin"
    (src / "pkg" / "sub" / long_name).write_bytes(bytes(range(256)) * 40)
    os.symlink("a.txt", src / "pkg" / "link")
    os.chmod(src / "pkg" / "a.txt", 0o640)
    buf = io.BytesIO()
    with tarfile.open(fileobj=buf, mode="w:gz" if gz else "w", format=fmt) as tf:
        tf.add(src / "pkg", arcname="pkg")
    p = tmp_path / ("t.tar.gz" if gz else "t.tar")
    p.write_bytes(buf.getvalue())
    return p


def _compare(a, b):
    for root, dirs, files in os.walk(a):
        rel = os.path.relpath(root, a)
        for d in dirs:
            assert os.path.isdir(os.path.join(b, rel, d)), d
        for f in files:
            pa, pb = os.path.join(root, f), os.path.join(b, rel, f)
            if os.path.islink(pa):
                assert os.readlink(pa) == os.readlink(pb)
            else:
                assert open(pa, "rb").read() == open(pb, "rb").read(), f
                assert (os.stat(pa).st_mode & 0o777) == (os.stat(pb).st_mode & 0o777), f
                # pax records (sub-second times) are skipped, as in the reference: whole seconds of the ustar field
                assert abs(os.stat(pa).st_mtime - os.stat(pb).st_mtime) <= 1.0, f


def _run(tmp_path, fmt, gz, gunzip):
    import zippy_b200.tarballs as tb
    p = _make(tmp_path, fmt, gz)
    gold = tmp_path / "gold"
    with tarfile.open(p) as tf:
        tf.extractall(gold, filter="fully_trusted")
    mine = tmp_path / "mine"
    tb.extract_all(str(p), str(mine), gunzip)
    _compare(str(gold), str(mine))
    _compare(str(mine), str(gold))
    return tb, p, mine


@pytest.mark.parametrize("fmt", [tarfile.GNU_FORMAT, tarfile.USTAR_FORMAT, tarfile.PAX_FORMAT])
@pytest.mark.parametrize("gz", [False, True])
def test_tar_extract_matches_tarfile_cpu(tmp_path, fmt, gz):
    gunzip = lambda b: zlib.decompress(b, 31)
    tb, p, mine = _run(tmp_path, fmt, gz, gunzip)
    from zippy_b200 import ZippyError
    with pytest.raises(ZippyError):            # destination exists
        tb.extract_all(str(p), str(mine), gunzip)
    data = p.read_bytes() if not gz else zlib.decompress(p.read_bytes(), 31)
    with pytest.raises(ZippyError):            # truncated archive
        tb.read_tarball(data[:512 * 3 + 100])   # ends inside a 512-byte block
    bad = bytearray(data[:512])
    bad[124:136] = b"00000000000\0"            # size 0
    bad[156] = ord("7")                        # an unsupported header type
    with pytest.raises(ZippyError):
        tb.read_tarball(bytes(bad))


def test_tar_rejects_unsafe_paths_cpu(tmp_path):
    import zippy_b200.tarballs as tb
    from zippy_b200 import ZippyError
    buf = io.BytesIO()
    with tarfile.open(fileobj=buf, mode="w") as tf:
        ti = tarfile.TarInfo("../evil.txt")
        ti.size = 1
        tf.addfile(ti, io.BytesIO(b"x"))
    p = tmp_path / "evil.tar"
    p.write_bytes(buf.getvalue())
    with pytest.raises(ZippyError):
        tb.extract_all(str(p), str(tmp_path / "out"))
    assert not (tmp_path / "out").exists()


def _run_cpp_tar(tmp_path, link_args):
    """include/zippy_b200_tar.hpp (the C++ form of tarballs.nim) through tests/native/cpp_tar_test.cpp:
    its view of a GNU-format .tar.gz must equal tarfile's, entry by entry."""
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    exe = str(tmp_path / "cpp_tar_test")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-o", exe, os.path.join(here, "native", "cpp_tar_test.cpp")] + link_args)
    p = _make(tmp_path, tarfile.GNU_FORMAT, True)
    out = subprocess.run([exe, str(p)], capture_output=True, text=True, timeout=600)
    lines = out.stdout.strip().split("\n")
    assert out.returncode == 0 and lines[-1] == "OK", (out.stdout, out.stderr)
    want = []
    with tarfile.open(p) as tf:
        for m in tf.getmembers():
            kind = "f" if m.isfile() else "d" if m.isdir() else "l"
            data = tf.extractfile(m).read() if m.isfile() else (m.linkname.encode() if m.issym() else b"")
            name = m.name + ("/" if m.isdir() else "")
            want.append("%s|%s|%d|%d|%o|%d" % (kind, name, len(data), zlib.crc32(data) if m.isfile() else 0, m.mode & 0o777,
                                               int(m.mtime)))
    assert lines[:-1] == want


def test_cpp_tar_layer_header_walk_cpu(tmp_path):
    here = os.path.dirname(os.path.abspath(__file__))
    _run_cpp_tar(tmp_path, [os.path.join(here, "native", "mock_abi_zlib.cpp"), "-lz"])


@pytest.mark.gpu
def test_cpp_tar_layer_gpu(tmp_path):
    libdir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "zippy_b200")
    _run_cpp_tar(tmp_path, ["-L" + libdir, "-l:libzippy_b200.so", "-Wl,-rpath," + libdir])


@pytest.mark.gpu
def test_tar_gz_extract_gpu(tmp_path):
    _run(tmp_path, tarfile.GNU_FORMAT, True, None)
