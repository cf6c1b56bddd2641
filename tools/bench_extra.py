#!/usr/bin/env python3
"""Secondary measurements for BASELINE.json configs 3 and 4 (not the driver's bench line):
  c3: batch uncompress of 65536 gzip members = the 23 reference fixtures tiled,
      byte-exact against the fixtures' manifest; GiB/s of compressed input and of output.
  c4: compress level=Default of urls.10K tiled to ~4 GiB, total size vs the oracle's level -1.
Device-resident buffers, CUDA events on the ctx stream, 1 GPU."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GIB = float(1 << 30)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="c3,c4")
    ap.add_argument("--members", type=int, default=65536)
    ap.add_argument("--tiles", type=int, default=6118)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--c5-blocks", type=int, default=131072, help="blocks of this GPU's shard of config 5 (1048576 / 8)")
    ap.add_argument("--c5-first", type=int, default=0, help="global index of the shard's first block")
    args = ap.parse_args()
    import torch
    import zippy_b200 as z
    from oracle import oracle as o
    from tests import util
    dev = torch.device("cuda", 0)
    ctx = z.Context(0)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream or ctx.LEGACY_DEFAULT_STREAM)   # same stream as torch's work: ordered
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    if "c3" in args.what:
        golden = util.load_golden()
        names = sorted(n for n in golden if n.endswith(".gz"))
        assert len(names) == 23
        cyc = [golden[n][0] for n in names]
        cyc_bytes = np.frombuffer(b"".join(cyc), dtype=np.uint8)
        n = args.members
        reps = (n + 22) // 23
        d_src = torch.from_numpy(cyc_bytes.copy()).to(dev).repeat(reps)
        lens = np.array([len(c) for c in cyc] * reps, dtype=np.uint64)[:n]
        offs = np.zeros(n + 1, dtype=np.uint64)
        np.cumsum(lens, out=offs[1:])
        sizes, st = ctx.uncompressed_sizes_device(d_src.data_ptr(), offs, z.dfDetect)
        assert not st.any()
        doffs = np.zeros(n + 1, dtype=np.uint64)
        np.cumsum(sizes, out=doffs[1:])
        d_dst = torch.empty(int(doffs[n]) + 64, dtype=torch.uint8, device=dev)
        for _ in range(2):
            out_lens, st = ctx.uncompress_batch_device(d_src.data_ptr(), offs, z.dfDetect, d_dst.data_ptr(), doffs)
        assert not st.any() and (out_lens == sizes).all()
        host = d_dst[:int(doffs[23])].cpu().numpy()
        for i, nme in enumerate(names):  # byte-exact against the reference's fixtures
            assert util.sha(host[int(doffs[i]):int(doffs[i + 1])].tobytes()) == golden[nme][1]["sha256"], nme
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(args.steps):
            ctx.uncompress_batch_device(d_src.data_ptr(), offs, z.dfDetect, d_dst.data_ptr(), doffs)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        t = ctx.timing()
        print(json.dumps({"workload": "C3: %d gzip members (23 reference fixtures tiled), batch uncompress" % n,
                          "in_bytes": int(offs[n]), "out_bytes": int(doffs[n]), "ms": ms,
                          "in_gibs": int(offs[n]) / GIB / (ms / 1e3), "out_gibs": int(doffs[n]) / GIB / (ms / 1e3),
                          "inflate_ms": t["inflate_ms"], "verify_ms": t["verify_ms"],
                          "parity": "first cycle sha256 == manifest; all members CRC+ISIZE verified on device"}))
        del d_src, d_dst

    if "crc" in args.what:
        # standalone blocked checksums: 65536 x 64 KiB buffers, then the same 4 GiB as ONE buffer
        import zlib
        n = args.members
        g = torch.Generator(device=dev)
        g.manual_seed(1)
        d_src = torch.randint(0, 256, (n * 65536,), dtype=torch.uint8, device=dev, generator=g)
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) \
            if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
        for label, offs in (("%d x 64 KiB" % n, np.arange(n + 1, dtype=np.uint64) * 65536),
                            ("1 x %d MiB" % (n // 16), np.array([0, n * 65536], dtype=np.uint64))):
            for kind in ("crc32", "adler32"):
                times = []
                for _ in range(5):   # the first calls also pay for the piece table upload and cold TLBs
                    out = ctx.checksum_batch_device(d_src.data_ptr(), offs, kind)
                    times.append(ctx.timing()["checksum_ms"])
                ms = float(np.median(times[2:]))
                h = d_src[:65536].cpu().numpy().tobytes() if len(offs) > 2 else None
                if h is not None:
                    assert int(out[0]) == (zlib.crc32(h) if kind == "crc32" else zlib.adler32(h))
                gbs = n * 65536 / (ms / 1e3) / 1e9
                print(json.dumps({"workload": "checksum %s, %s" % (kind, label), "ms": ms, "GB_s": gbs,
                                  "hbm_peak_GB_s": peak, "frac_of_hbm_peak": gbs / peak, "value0": int(out[0])}))
        del d_src

    if "c5" in args.what:
        # BASELINE config 5, one GPU's shard: 131072 x 64 KiB blocks of mixed entropy (SURVEY 8d:
        # class = sm64(0xC5 + i) mod 8: 0-3 text, 4 urls.10K window, 5 html window, 6 random bytes,
        # 7 run-length blob), level 1 gzip, then the GPU inflates everything back.
        corpus = util.load_corpus()
        T = util.text_corpus(corpus)
        nb = args.c5_blocks
        first = args.c5_first
        cls = np.array([util._sm64(0xC5 + first + i) % 8 for i in range(nb)], dtype=np.int64)
        d_src = torch.empty(nb * 65536, dtype=torch.uint8, device=dev)
        view = d_src.view(nb, 65536)
        for name, sel in (("text", cls < 4), ("urls", cls == 4), ("html", cls == 5)):
            raw = T if name == "text" else corpus["urls.10K" if name == "urls" else "html"]
            idx = np.nonzero(sel)[0]
            if not len(idx):
                continue
            offs_b = np.array([(util._sm64(0xC5C5 + first + int(i)) >> 3) % (len(raw) - 65536) for i in idx], dtype=np.int64)
            win = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(dev).unfold(0, 65536, 1)
            for s0 in range(0, len(idx), 4096):
                rows = torch.index_select(win, 0, torch.from_numpy(offs_b[s0:s0 + 4096]).to(dev))
                view[torch.from_numpy(idx[s0:s0 + 4096]).to(dev)] = rows
        g = torch.Generator(device=dev)
        g.manual_seed(0xC5 + first)
        idx = torch.from_numpy(np.nonzero(cls == 6)[0]).to(dev)
        view[idx] = torch.randint(0, 256, (len(idx), 65536), dtype=torch.uint8, device=dev, generator=g)
        idx = torch.from_numpy(np.nonzero(cls == 7)[0]).to(dev)
        need = len(idx) * 65536
        runs = torch.randint(1, 256, (need // 100 + 1024,), device=dev, generator=g)
        vals = torch.randint(0, 256, (len(runs),), dtype=torch.uint8, device=dev, generator=g)
        blob = torch.repeat_interleave(vals, runs)[:need]
        assert blob.numel() == need
        view[idx] = blob.view(len(idx), 65536)
        offs = np.arange(nb + 1, dtype=np.uint64) * 65536
        cap = nb * (65536 + 96) + 4096
        d_dst = torch.empty(cap, dtype=torch.uint8, device=dev)
        for _ in range(2):
            oo = ctx.compress_batch_device(d_src.data_ptr(), offs, 1, z.dfGzip, d_dst.data_ptr(), cap)
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(args.steps):
            oo = ctx.compress_batch_device(d_src.data_ptr(), offs, 1, z.dfGzip, d_dst.data_ptr(), cap)
        e1.record(stream)
        torch.cuda.synchronize()
        c_ms = e0.elapsed_time(e1) / args.steps
        sizes = np.diff(oo.astype(np.int64))
        per_class = {int(k): float(sizes[cls == k].sum()) / (float((cls == k).sum()) * 65536.0) for k in range(8) if (cls == k).any()}
        d_back = torch.empty(nb * 65536, dtype=torch.uint8, device=dev)
        lens, st = ctx.uncompress_batch_device(d_dst.data_ptr(), oo, z.dfDetect, d_back.data_ptr(), offs)
        assert not st.any() and (lens == 65536).all() and torch.equal(d_back, d_src)
        torch.cuda.synchronize()
        e0.record(stream)
        ctx.uncompress_batch_device(d_dst.data_ptr(), oo, z.dfDetect, d_back.data_ptr(), offs)
        e1.record(stream)
        torch.cuda.synchronize()
        u_ms = e0.elapsed_time(e1)
        for i in (int(np.nonzero(cls == k)[0][0]) for k in range(8) if (cls == k).any()):   # one block per class through the oracle
            m = d_dst[int(oo[i]):int(oo[i + 1])].cpu().numpy().tobytes()
            assert o.uncompress(m) == d_src[i * 65536:(i + 1) * 65536].cpu().numpy().tobytes()
        print(json.dumps({"workload": "C5 shard: %d x 64 KiB mixed-entropy blocks (first block %d), level 1 gzip" % (nb, first),
                          "in_bytes": nb * 65536, "out_bytes": int(oo[nb]), "compress_ms": c_ms,
                          "compress_in_gibs": nb * 65536 / GIB / (c_ms / 1e3), "ratio": int(oo[nb]) / float(nb * 65536),
                          "ratio_by_class": per_class, "uncompress_ms": u_ms,
                          "uncompress_out_gibs": nb * 65536 / GIB / (u_ms / 1e3),
                          "parity": "GPU round trip bit-exact over the shard; one block per class through the oracle"}))
        del d_src, d_dst, d_back

    if "big" in args.what:
        # ONE large input -> one gzip member of 64 KiB chunks -> back, device-resident;
        # the inflate runs as parallel segments (and, for comparison, serially on a 32 MiB slice)
        T = util.text_corpus(util.load_corpus())
        reps = (1 << 30) // len(T) + 1
        d_src = torch.frombuffer(bytearray(T), dtype=torch.uint8).to(dev).repeat(reps)[:1 << 30].contiguous()
        n_bytes = d_src.numel()
        offs = np.array([0, n_bytes], dtype=np.uint64)
        cap = n_bytes + n_bytes // 8 + (1 << 20)
        d_dst = torch.empty(cap, dtype=torch.uint8, device=dev)
        d_back = torch.empty(n_bytes, dtype=torch.uint8, device=dev)
        for _ in range(2):
            oo = ctx.compress_batch_device(d_src.data_ptr(), offs, 1, z.dfGzip, d_dst.data_ptr(), cap)
        torch.cuda.synchronize()
        e0.record(stream)
        oo = ctx.compress_batch_device(d_src.data_ptr(), offs, 1, z.dfGzip, d_dst.data_ptr(), cap)
        e1.record(stream)
        torch.cuda.synchronize()
        c_ms = e0.elapsed_time(e1)
        for _ in range(2):
            lens, st = ctx.uncompress_batch_device(d_dst.data_ptr(), oo, z.dfDetect, d_back.data_ptr(), offs)
        assert not st.any() and int(lens[0]) == n_bytes and torch.equal(d_back, d_src)
        torch.cuda.synchronize()
        e0.record(stream)
        ctx.uncompress_batch_device(d_dst.data_ptr(), oo, z.dfDetect, d_back.data_ptr(), offs)
        e1.record(stream)
        torch.cuda.synchronize()
        u_ms = e0.elapsed_time(e1)
        launches = ctx.timing()["kernel_launches"]
        os.environ["ZB200_BIG_MEMBER_BYTES"] = str(1 << 62)
        ctx2 = z.Context(0)
        del os.environ["ZB200_BIG_MEMBER_BYTES"]
        small = 32 << 20
        o2 = ctx2.compress_batch_device(d_src.data_ptr(), np.array([0, small], dtype=np.uint64), 1, z.dfGzip, d_dst.data_ptr(), cap)
        import time
        t0 = time.time()
        lens, st = ctx2.uncompress_batch_device(d_dst.data_ptr(), o2, z.dfDetect, d_back.data_ptr(), np.array([0, small], dtype=np.uint64))
        torch.cuda.synchronize()
        ser = time.time() - t0
        assert not st.any() and torch.equal(d_back[:small], d_src[:small])
        print(json.dumps({"workload": "one 1 GiB input as a single gzip member (level 1), device-resident",
                          "compress_ms": c_ms, "compress_gibs": n_bytes / GIB / (c_ms / 1e3), "member_bytes": int(oo[1]),
                          "uncompress_ms": u_ms, "uncompress_out_gibs": n_bytes / GIB / (u_ms / 1e3),
                          "uncompress_kernel_launches": int(launches),
                          "serial_decode_32MiB_s": ser, "serial_out_gibs": small / GIB / ser}))
        ctx2.close()
        del d_src, d_dst, d_back

    if "c4" in args.what:
        raw = util.load_corpus()["urls.10K"]
        n = args.tiles
        d_src = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(dev).repeat(n)
        offs = np.arange(n + 1, dtype=np.uint64) * len(raw)
        cap = n * (len(raw) + 256)
        d_dst = torch.empty(cap, dtype=torch.uint8, device=dev)
        for _ in range(2):
            oo = ctx.compress_batch_device(d_src.data_ptr(), offs, z.DefaultCompression, z.dfGzip, d_dst.data_ptr(), cap)
        first = d_dst[:int(oo[1])].cpu().numpy().tobytes()
        assert o.uncompress(first) == raw
        ref = len(o.compress(raw, o.DefaultCompression, o.dfGzip))
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(args.steps):
            ctx.compress_batch_device(d_src.data_ptr(), offs, z.DefaultCompression, z.dfGzip, d_dst.data_ptr(), cap)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        print(json.dumps({"workload": "C4: urls.10K x %d tiles, compress level=Default dfGzip" % n,
                          "in_bytes": int(offs[n]), "out_bytes": int(oo[n]), "ms": ms,
                          "in_gibs": int(offs[n]) / GIB / (ms / 1e3), "ratio": int(oo[n]) / float(offs[n]),
                          "oracle_level6_ratio": ref / float(len(raw)),
                          "size_vs_reference": int(oo[n]) / float(ref * n), "timing": ctx.timing()}))


if __name__ == "__main__":
    main()
