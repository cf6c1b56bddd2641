#!/usr/bin/env python3
"""bench.py -- BASELINE.json's metric on BASELINE.json's config.

Metric : GiB/s of INPUT for batched compress (level 1 = BestSpeed, gzip members), next to the
         CPU baseline and the HBM roofline of the dominant kernel.
Step   : one pass of the compress hot path over one batch of synthetic input.
Workload (N=1, SURVEY.md 8(d) "C2"): 65536 x 64 KiB blocks cut from the text corpus
         T = alice29 || asyoulik || lcet10 || plrabn12 at seeded offsets
         o_i = splitmix64(0xC2 + i) mod (|T| - 65536); 4 GiB per GPU, one gzip member per block.
         N>1: every rank gets its own 65536 blocks (indices rank*65536 ..), i.e. weak scaling;
         the only collective is one NCCL all_gather of the per-member compressed sizes, from
         which every rank derives the global concatenation offsets.

value  : whole-job GiB/s with inputs already resident in HBM (device variant of the C ABI).
e2e    : same metric through the host-buffer C-ABI call: pinned host input -> H2D -> kernels ->
         D2H of the compressed members, all inside the timed region.
--impl reference : the CPU path (oracle port of the reference; Nim is not available, so the
         reference itself cannot be built -- see DESIGN.md) on all host cores, bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GIB = float(1 << 30)
BLOCK = 65536
BLOCKS_PER_GPU = 65536


def sm64(x):
    x = (x + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return z ^ (z >> 31)


def text_corpus():
    from tests import util
    return util.text_corpus(util.load_corpus())


def bench_config(n, level, world):
    return {"workload": "C2: batch %d x 64 KiB synthetic text-entropy blocks per GPU, compress level=%d "
                        "dfGzip, one gzip member per block" % (n, level),
            "blocks_per_gpu": n, "block_bytes": BLOCK, "level": level, "data_format": "dfGzip",
            "l2": "inputs (4 GiB/GPU) larger than L2; no flush needed", "launch_group_chunks": n,
            "parallelism": "independent members sharded over %d GPU(s); NCCL all_gather of sizes" % world}


def block_offsets(T_len, first, count):
    return np.array([sm64(0xC2 + i) % (T_len - BLOCK) for i in range(first, first + count)], dtype=np.int64)


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md clocks line).

    Primary source is NVML in-process (a few microseconds per query, so even a 0.3 s region gets
    dozens of samples); `nvidia-smi -lms` is the fallback when pynvml cannot open the device."""

    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, gpu_index):
        self.rows = []  # (sm_mhz, sm_max_mhz, set(reasons))
        self.stop = threading.Event()
        self.proc = None
        self.gpu = gpu_index
        self.nvml = None
        self.source = None

    def _open_nvml(self):
        import pynvml
        import torch
        pynvml.nvmlInit()
        h = None
        try:
            uuid = str(torch.cuda.get_device_properties(self.gpu).uuid)
            h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        except Exception:
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.gpu
            if vis:
                try:
                    idx = int(vis.split(",")[self.gpu])
                except Exception:
                    idx = self.gpu
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)  # probe
        self.nvml, self.h = pynvml, h
        self.masks = [(pynvml.nvmlClocksThrottleReasonHwSlowdown, "hw_slowdown"),
                      (pynvml.nvmlClocksThrottleReasonHwThermalSlowdown, "hw_thermal_slowdown"),
                      (pynvml.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_thermal_slowdown"),
                      (pynvml.nvmlClocksThrottleReasonSwPowerCap, "sw_power_cap")]

    def _sample_nvml(self):
        n, h = self.nvml, self.h
        mx = float(n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM))
        while not self.stop.is_set():
            try:
                sm = float(n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM))
                bits = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                self.rows.append((sm, mx, {nm for m, nm in self.masks if bits & m}))
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        try:
            self._open_nvml()
            self.source = "nvml"
            self.t = threading.Thread(target=self._sample_nvml, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.source = "nvidia-smi"
        self.t = threading.Thread(target=self._read_smi, daemon=True)
        self.t.start()

    def _read_smi(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.strip().split(",")]
            if len(f) >= 7:
                try:
                    self.rows.append((float(f[0]), float(f[1]),
                                      {nm for nm, v in zip(self.NAMES, f[3:7]) if v.lower().startswith("active")}))
                except ValueError:
                    pass
            if self.stop.is_set():
                break

    def mark(self):
        """Samples taken before this point (start-up, idle clocks) are dropped."""
        self.first = len(self.rows)

    def finish(self):
        if self.source is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["no clock source (nvml, nvidia-smi)"]}
        self.stop.set()
        if self.proc:
            try:
                self.proc.terminate()
            except Exception:
                pass
        rows = self.rows[getattr(self, "first", 0):]
        sm = [r[0] for r in rows]
        reasons = set()
        for r in rows:
            reasons |= r[2]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(r[1] for r in rows) if rows else None,
                "samples": len(sm), "source": self.source, "reasons": sorted(reasons)}


def host_threads():
    """Threads the CPU arm may really use: the affinity mask, capped by a cgroup CPU quota if there is one."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    try:
        q, p = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if q != "max":
            n = min(n, max(1, int(int(q) / float(p) + 0.5)))
    except Exception:
        pass
    return max(1, n)


def cpu_baseline(T, n_blocks_hint, seconds=12.0, threads=None, level=1):
    """Oracle (port of the reference's level-1 path) over independent blocks on all host cores."""
    from oracle import oracle as o
    cores = threads or host_threads()
    offs = block_offsets(len(T), 0, 4096)
    Tn = np.frombuffer(T, dtype=np.uint8)

    def make(nb):
        buf = np.empty(nb * BLOCK, dtype=np.uint8)
        for i in range(nb):
            s = int(offs[i % len(offs)])
            buf[i * BLOCK:(i + 1) * BLOCK] = Tn[s:s + BLOCK]
        return buf, np.arange(nb + 1, dtype=np.uint64) * BLOCK

    pilot_n = max(cores * 4, 16)
    buf, bo = make(pilot_n)
    t0 = time.perf_counter()
    o.compress_batch(buf, bo, level, o.dfGzip, threads=cores)
    dt = max(time.perf_counter() - t0, 1e-3)
    nb = int(min(max(pilot_n, pilot_n * seconds / dt), 65536))
    buf, bo = make(nb)
    t0 = time.perf_counter()
    total, lens, st = o.compress_batch(buf, bo, level, o.dfGzip, threads=cores)
    dt = time.perf_counter() - t0
    assert not st.any()
    return {"value": nb * BLOCK / GIB / dt, "unit": "GiB/s", "cores": cores, "kind": "port",
            "sample": "%d x 64 KiB C2 text blocks, oracle level %d gzip, %d threads, %.1f s" % (nb, level, cores, dt),
            "ratio": float(total) / (nb * BLOCK)}, nb, dt


def cpu_uncompress_baseline(T, seconds=4.0, threads=None):
    """Oracle inflate (port of inflate.nim) of level-1 gzip members of C2 blocks, all host cores."""
    from oracle import oracle as o
    cores = threads or host_threads()
    offs = block_offsets(len(T), 0, 512)
    members = [o.compress(T[int(s):int(s) + BLOCK], 1, o.dfGzip) for s in offs]
    reps = max(1, (cores * 8 + len(members) - 1) // len(members))
    lens = np.array([len(m) for m in members] * reps, dtype=np.uint64)
    base = np.frombuffer(b"".join(members) * reps, dtype=np.uint8)
    mo = np.zeros(len(lens) + 1, dtype=np.uint64)
    np.cumsum(lens, out=mo[1:])
    t0 = time.perf_counter()
    total, _, st = o.uncompress_batch(base, mo, o.dfGzip, threads=cores)
    dt = max(time.perf_counter() - t0, 1e-4)
    rounds = int(max(1, min(64, seconds / dt)))
    t0 = time.perf_counter()
    for _ in range(rounds):
        total, _, st = o.uncompress_batch(base, mo, o.dfGzip, threads=cores)
    dt = time.perf_counter() - t0
    assert not st.any() and int(total) == len(lens) * BLOCK
    return {"out_gibs": rounds * len(lens) * BLOCK / GIB / dt, "in_gibs": rounds * int(mo[-1]) / GIB / dt, "cores": cores,
            "kind": "port", "sample": "%d gzip members of C2 blocks x %d rounds, oracle inflate, %d threads, %.1f s"
            % (len(lens), rounds, cores, dt)}


def run_reference(args, rank, world):
    """--impl reference: the CPU arm.  Rank 0 only."""
    if rank != 0:
        return
    T = text_corpus()
    cores = host_threads()
    per_step = max(6.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    times, nbs = [], []
    for s in range(args.warmup + args.steps):
        cb, nb, dt = cpu_baseline(T, 0, seconds=per_step, threads=cores, level=args.level)
        if s >= args.warmup:
            times.append(dt)
            nbs.append(nb)
    val = sum(nbs) * BLOCK / GIB / sum(times)
    out = {"impl": "reference", "metric": "compress_level1_gzip_input_throughput", "value": val, "unit": "GiB/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "u8", "data": "synthetic",
           "config": bench_config(BLOCKS_PER_GPU, args.level, args.gpus),
           "cpu_baseline": {"value": val, "unit": "GiB/s", "cores": cores, "kind": "port",
                            "sample": "%d x 64 KiB C2 blocks per step (bounded sample of the config), oracle port of "
                                      "zippy level 1 (Nim unavailable: the reference cannot be compiled here), "
                                      "all %d host threads" % (int(np.mean(nbs)), cores)},
           "e2e": {"value": val, "unit": "GiB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    out["uncompress"] = {"cpu_baseline": cpu_uncompress_baseline(T, seconds=3.0, threads=cores)}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--blocks", type=int, default=BLOCKS_PER_GPU, help="64 KiB blocks per GPU (default = config)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--level", type=int, default=1)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import zippy_b200 as z

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    n = args.blocks
    T = text_corpus()
    # ---- synthetic batch, generated on the device from the seeded offsets ----
    T_d = torch.frombuffer(bytearray(T), dtype=torch.uint8).to(dev)
    offs = torch.from_numpy(block_offsets(len(T), rank * n, n)).to(dev)
    windows = T_d.unfold(0, BLOCK, 1)
    d_src = torch.empty(n * BLOCK, dtype=torch.uint8, device=dev)
    step_rows = 4096
    for s in range(0, n, step_rows):
        e = min(n, s + step_rows)
        d_src[s * BLOCK:e * BLOCK] = torch.index_select(windows, 0, offs[s:e]).reshape(-1)
    src_offsets = np.arange(n + 1, dtype=np.uint64) * BLOCK
    cap = n * (BLOCK + 64) + 4096
    d_dst = torch.empty(cap, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()

    # one launch group per step: each kernel runs once per step, so the per-launch durations
    # and algorithmic bytes below are per-step figures (default groups are 32768 chunks)
    os.environ.setdefault("ZB200_DEV_GROUP_CHUNKS", str(max(n, 1)))
    ctx = z.Context(local_rank)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    clocks = ClockSampler(local_rank)
    clocks.start()   # samples before clocks.mark() (start-up, warm-up) are dropped

    sizes_all = None

    def step_device():
        nonlocal sizes_all
        oo = ctx.compress_batch_device(d_src.data_ptr(), src_offsets, args.level, z.dfGzip, d_dst.data_ptr(), cap)
        if world > 1:
            # the path's one exchange: all ranks learn every member's compressed size
            mine = torch.from_numpy((oo[1:] - oo[:-1]).astype(np.int64)).to(dev)
            allsz = torch.empty(world * n, dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(allsz, mine)
            sizes_all = torch.cumsum(allsz, 0)   # global concatenation offsets
        return oo

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- warm-up + parity checks outside the timed region ----
    oo = None
    for _ in range(max(args.warmup, 1)):
        oo = step_device()
    sync_all()
    comp_bytes = int(oo[n])
    # (1) oracle + zlib inflate a sample of members; (2) GPU inflates the whole batch back
    if rank == 0:
        import zlib
        from oracle import oracle as o
        host = d_dst[:int(oo[64])].cpu().numpy()
        for i in range(64):
            m = host[int(oo[i]):int(oo[i + 1])].tobytes()
            want = d_src[i * BLOCK:(i + 1) * BLOCK].cpu().numpy().tobytes()
            assert o.uncompress(m) == want and zlib.decompress(m, 31) == want, "parity failure on member %d" % i
    d_back = torch.empty(n * BLOCK, dtype=torch.uint8, device=dev)
    lens, st = ctx.uncompress_batch_device(d_dst.data_ptr(), oo, z.dfDetect, d_back.data_ptr(), src_offsets)
    t_inf = ctx.timing()
    assert not st.any() and bool((lens == BLOCK).all()), "GPU inflate reported errors"
    assert torch.equal(d_back, d_src), "round trip mismatch at full size"
    del d_back
    inflate_ms = t_inf["inflate_ms"] + t_inf["verify_ms"]

    # ---- timed region: device-resident ----
    kern = {"lz_ms": 0.0, "huff_ms": 0.0, "scan_ms": 0.0, "pack_ms": 0.0}
    sync_all()
    clocks.mark()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    launches = 0
    for _ in range(args.steps):
        step_device()
        t = ctx.timing()
        for k in kern:
            kern[k] += t[k]
        launches += t["kernel_launches"]
    e1.record(stream)
    sync_all()
    ms = e0.elapsed_time(e1)
    clk = clocks.finish()
    tmax = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_total = float(tmax.item())
    ms_per_step = ms_total / args.steps
    value = world * n * BLOCK / GIB / (ms_per_step / 1e3)

    # ---- e2e: host (pinned) buffers through the host-buffer C-ABI call ----
    e2e = None
    unc_e2e = None
    if not args.no_e2e:
        hcap = comp_bytes + (64 << 20)
        h_src = torch.empty(n * BLOCK, dtype=torch.uint8).pin_memory()
        h_src.copy_(d_src)
        h_dst = torch.empty(hcap, dtype=torch.uint8).pin_memory()
        from zippy_b200 import _native
        L = _native.lib()
        out_offs = np.zeros(n + 1, dtype=np.uint64)
        stat = np.zeros(n, dtype=np.int32)

        def step_host():
            rc = L.zb200_compress_batch(ctx._h, h_src.data_ptr(), src_offsets.ctypes.data, n, args.level, z.dfGzip,
                                        None, h_dst.data_ptr(), hcap, out_offs.ctypes.data, stat.ctypes.data)
            assert rc == 0, rc
            if world > 1:
                mine = torch.from_numpy((out_offs[1:] - out_offs[:-1]).astype(np.int64)).to(dev)
                allsz = torch.empty(world * n, dtype=torch.int64, device=dev)
                dist.all_gather_into_tensor(allsz, mine)

        for _ in range(2):
            step_host()
        sync_all()
        e0.record(stream)
        for _ in range(args.steps):
            step_host()
        e1.record(stream)
        sync_all()
        t2 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        ms2 = float(t2.item()) / args.steps
        th = ctx.timing()
        e2e = {"value": world * n * BLOCK / GIB / (ms2 / 1e3), "unit": "GiB/s", "ms_per_step": ms2,
               "h2d_bytes_per_step": int(n * BLOCK), "d2h_bytes_per_step": int(out_offs[n]),
               "h2d_ms": th["h2d_ms"], "d2h_ms": th["d2h_ms"]}
        assert int(out_offs[n]) == comp_bytes
        # the reverse direction through the host-buffer call (rank 0, 1 GPU runs only): compressed
        # members in pinned host memory -> original bytes in pinned host memory
        unc_e2e = None
        if world == 1:
            lens_b = np.zeros(n, dtype=np.uint64)
            h_src.zero_()

            def step_back():
                rc = L.zb200_uncompress_batch(ctx._h, h_dst.data_ptr(), out_offs.ctypes.data, n, z.dfDetect,
                                              h_src.data_ptr(), src_offsets.ctypes.data, lens_b.ctypes.data,
                                              stat.ctypes.data)
                assert rc == 0, rc

            step_back()
            sync_all()
            assert not stat.any() and bool((lens_b == BLOCK).all())
            assert torch.equal(h_src[:64 << 20], d_src[:64 << 20].cpu()), "host round trip mismatch"
            e0.record(stream)
            for _ in range(args.steps):
                step_back()
            e1.record(stream)
            sync_all()
            ms3 = e0.elapsed_time(e1) / args.steps
            tb = ctx.timing()
            unc_e2e = {"out_gibs": n * BLOCK / GIB / (ms3 / 1e3), "in_gibs": comp_bytes / GIB / (ms3 / 1e3), "ms": ms3,
                       "h2d_bytes": comp_bytes, "d2h_bytes": int(n * BLOCK), "h2d_ms": tb["h2d_ms"], "d2h_ms": tb["d2h_ms"],
                       "inflate_ms": tb["inflate_ms"] + tb["verify_ms"]}
        del h_src, h_dst

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    per_step = {k: v / args.steps for k, v in kern.items()}
    dom = max(per_step, key=per_step.get)
    algo_bytes = n * BLOCK + comp_bytes                      # N_in * (1 + r) per launch
    achieved = algo_bytes / (per_step[dom] / 1e3) / 1e9 if per_step[dom] > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": "k_" + dom[:-3], "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "peak_source": "measured" if "hbm_gbs" in peaks else "fallback",
                "algorithmic_bytes_per_launch": algo_bytes, "kernel_ms": per_step[dom], "traffic": None,
                "kernel_ms_all": per_step}
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "latest_traffic.json")))
        roofline["traffic"] = prof.get(roofline["kernel"])
    except Exception:
        pass

    cpu = None
    cpu_unc = None
    if not args.no_cpu:
        cpu, _, _ = cpu_baseline(T, n, level=args.level)
        cpu_unc = cpu_uncompress_baseline(T)

    out = {"metric": "compress_level1_gzip_input_throughput", "value": value, "unit": "GiB/s", "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
           "config": bench_config(n, args.level, world),
           "ratio": comp_bytes / float(n * BLOCK), "clocks": clk, "gpu_launches": int(launches),
           "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
           "uncompress": {"out_gibs": n * BLOCK / GIB / (inflate_ms / 1e3), "in_gibs": comp_bytes / GIB / (inflate_ms / 1e3),
                          "ms": inflate_ms, "note": "GPU inflate + CRC verify of this batch's own members (device-resident)",
                          "e2e": unc_e2e, "cpu_baseline": cpu_unc}}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
