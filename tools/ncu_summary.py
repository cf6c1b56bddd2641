#!/usr/bin/env python3
"""Summarise `ncu --set full` reports (gpurun_out/*.ncu-rep) into profiles/: one markdown table per kernel
and profiles/latest_traffic.json (DRAM bytes per launch, scaled to bench.py's launch size where the capture
used a smaller batch -- the scale factor is recorded).

usage: ncu_summary.py <tag> <out.md> <rep>[:<scale>[:<label>]] ...
  scale: factor from the captured launch to the bench's launch (e.g. 4 for a 16384-chunk capture of a
  65536-chunk step); label: note for the table"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = [("gpu__time_duration.sum", "duration"),
           ("smsp__inst_executed.sum", "warp instructions"),
           ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
           ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
           ("launch__registers_per_thread", "registers / thread"),
           ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps / scheduler / cycle"),
           ("lts__t_sector_hit_rate.pct", "L2 hit %"),
           ("dram__bytes_read.sum", "DRAM read"),
           ("dram__bytes_write.sum", "DRAM write"),
           ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
           ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard (L2/DRAM) per issue"),
           ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard (smem) per issue"),
           ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: fixed-latency wait per issue"),
           ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall: barrier per issue")]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}


def load(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {"name": r[hdr.index("Kernel Name")]}
        for m, _ in METRICS:
            if m in hdr:
                d[m] = (r[hdr.index(m)], units[hdr.index(m)])
        out.append(d)
    return out


def main():
    tag, out_md = sys.argv[1], sys.argv[2]
    traffic = {}
    md = ["# %s -- `ncu --set full --clock-control none` captures (B200, one GPU)\n" % tag,
          "Times under ncu are cold-cache and serialised: read the SHARES and the per-launch counters, not the absolutes.\n"]
    for spec in sys.argv[3:]:
        parts = spec.split(":")
        rep, scale = parts[0], float(parts[1]) if len(parts) > 1 else 1.0
        label = parts[2] if len(parts) > 2 else ""
        for k in load(rep):
            short = k["name"].split("(")[0]
            md.append("\n## `%s` -- %s (%s)\n" % (short, os.path.basename(rep), label or "captured launch = bench launch"))
            md.append("| metric | value |\n|---|---|")
            for m, nice in METRICS:
                if m in k:
                    md.append("| %s | %s %s |" % (nice, k[m][0], k[m][1]))
            try:
                rd = float(k["dram__bytes_read.sum"][0]) * UNIT[k["dram__bytes_read.sum"][1]]
                wr = float(k["dram__bytes_write.sum"][0]) * UNIT[k["dram__bytes_write.sum"][1]]
                key = short.replace("void ", "").split("<")[0]
                traffic[key] = {"dram_bytes_per_launch": int((rd + wr) * scale), "captured_bytes": int(rd + wr), "scale": scale,
                                "source": "profiles/%s (%s)" % (os.path.basename(out_md), os.path.basename(rep))}
                md.append("| DRAM read + write, scaled x%g to the bench launch | %.3f GB |" % (scale, (rd + wr) * scale / 1e9))
            except Exception:
                pass
    open(out_md, "w").write("\n".join(md) + "\n")
    p = os.path.join(ROOT, "profiles", "latest_traffic.json")
    json.dump(traffic, open(p, "w"), indent=1)
    print("wrote", out_md, p)


if __name__ == "__main__":
    main()
