#!/usr/bin/env python3
"""Regenerates tests/golden/ziparchives/ from the reference checkout (build container only).

Copies the reference's ZIP fixtures (tests/data/ziparchives/Bagnon-10.2.31.zip and cat.jpg -- an
archive appended to a JPEG; tests/test_ziparchives_read.nim:13,40-48) and records every entry's
name / length / sha256 / crc32 as extracted by Python's `zipfile` (an independent implementation,
the role `unzip` plays in the reference's test; never our code)."""
import hashlib
import json
import os
import shutil
import sys
import zipfile
import zlib

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/tests/data/ziparchives"
HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ziparchives")
os.makedirs(HERE, exist_ok=True)
manifest = {}
for name in ("Bagnon-10.2.31.zip", "cat.jpg"):
    shutil.copyfile(os.path.join(REF, name), os.path.join(HERE, name))
    entries = []
    with zipfile.ZipFile(os.path.join(HERE, name)) as zf:
        for info in zf.infolist():
            data = b"" if info.is_dir() else zf.read(info)
            entries.append({"name": info.filename, "is_dir": info.is_dir(), "len": len(data),
                            "sha256": hashlib.sha256(data).hexdigest(), "crc32": zlib.crc32(data)})
    manifest[name] = {"entries": entries, "ref_test": "tests/test_ziparchives_read.nim"}
json.dump(manifest, open(os.path.join(HERE, "zip_manifest.json"), "w"), indent=1, sort_keys=True)
print({k: len(v["entries"]) for k, v in manifest.items()})
