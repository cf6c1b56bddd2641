"""The driver's bench.py contract (task statement, section 4): the committed line of the last GPU run
has every key the driver and the judge read, and the CPU legs run here (no GPU needed for them)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _latest_line():
    """The newest committed bench line (profiles/bench_r<N>_final.json)."""
    import glob
    import re
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "bench_r*_final.json")),
                   key=lambda f: int(re.search(r"bench_r(\d+)_final", f).group(1)))
    return json.load(open(files[-1]))


def test_committed_bench_line_has_the_contract_keys():
    d = _latest_line()
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["unit"] == "GiB/s" and d["higher_is_better"] is True and d["scaling"] == "weak" and d["dtype"] == "u8"
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and "workload" in d["config"]
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in d["e2e"], k
    assert d["e2e"]["h2d_bytes_per_step"] == 65536 * 65536 and 0 < d["e2e"]["value"] < d["value"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in d["roofline"], k
    assert abs(d["roofline"]["frac"] - d["roofline"]["achieved"] / d["roofline"]["peak"]) < 1e-9
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in d["cpu_baseline"], k
    for k in ("sm_mhz", "sm_max_mhz", "reasons"):
        assert k in d["clocks"], k
    assert d["clocks"]["samples"] > 0 and not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown",
                                                                             "sw_thermal_slowdown"}


def test_cpu_legs_run_and_report():
    import bench
    T = bench.text_corpus()
    assert bench.host_threads() >= 1
    cb, nb, dt = bench.cpu_baseline(T, 0, seconds=0.5, threads=2, level=1)
    assert cb["kind"] == "port" and cb["cores"] == 2 and cb["value"] > 0 and 0.3 < cb["ratio"] < 0.6 and nb > 0
    cu = bench.cpu_uncompress_baseline(T, seconds=0.3, threads=2)
    assert cu["out_gibs"] > cu["in_gibs"] > 0 and cu["cores"] == 2
    cfg = bench.bench_config(65536, 1, 1)
    assert "65536 x 64 KiB" in cfg["workload"] and cfg["level"] == 1
