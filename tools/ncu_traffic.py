#!/usr/bin/env python
"""Fold one `ncu --set full` capture of k_trace into profiles/traffic.json (what bench.py's `roofline` quotes).

    python tools/ncu_traffic.py <report.ncu-rep> <key> <tris> <rays_per_launch> <build_tag> "<source note>"

<key> is the bench workload (c2, c3, c4) or a free name (sweep_1e7).  Takes the k_trace launch with the largest grid.
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PICK = {
    "inst_executed": "smsp__inst_executed.sum",
    "simt_lanes": "smsp__thread_inst_executed_per_inst_executed.ratio",
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex_pct": "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "lsu_writeback_pct": "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed",
    "l1_hit_pct": "l1tex__t_sector_hit_rate.pct",
    "l2_hit_pct": "lts__t_sector_hit_rate.pct",
    "lts_pct": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram_bytes_read": "dram__bytes_read.sum",
    "dram_bytes_write": "dram__bytes_write.sum",
    "lts_t_bytes": "lts__t_bytes.sum",
    "l1tex_t_bytes": "l1tex__t_bytes.sum",
    "kernel_ms_ncu": "gpu__time_duration.sum",
    "registers": "launch__registers_per_thread",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
}
SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "byte": 1.0, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0,
         "ns": 1e-6, "nsecond": 1e-6, "s": 1e3, "second": 1e3}


def main():
    rep, key, tris, rays, tag, note = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), sys.argv[5], sys.argv[6]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    best = None
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        if "k_trace" not in d["Kernel Name"]:
            continue
        grid = int(d["Grid Size"].strip("()").split(",")[0])
        if best is None or grid > best[0]:
            best = (grid, d)
    if best is None:
        raise SystemExit("no k_trace launch in the report")
    d = best[1]
    u = dict(zip(hdr, units))
    e = {"tris": tris, "rays_per_launch": rays, "build_tag": tag, "kernel": re.sub(r"\((int|bool)\)", "", d["Kernel Name"]).split("(")[0].replace("void <unnamed>::", "").replace("void ", "")[:60],
         "grid": best[0]}
    for name, metric in PICK.items():
        if metric in d and d[metric] != "":
            val = float(d[metric].replace(",", ""))
            unit = u.get(metric, "")
            if name.endswith("bytes") or name.startswith("dram_bytes") or name == "kernel_ms_ncu":
                val *= SCALE.get(unit, 1.0)
            e[name] = int(val) if (name.endswith("bytes") or name.startswith("dram_bytes") or name in ("inst_executed", "registers")) else round(val, 4)
    e["dram_bytes_per_launch"] = e.get("dram_bytes_read", 0) + e.get("dram_bytes_write", 0)
    e["source"] = note
    path = os.path.join(ROOT, "profiles", "traffic.json")
    tj = json.load(open(path)) if os.path.exists(path) else {}
    tj[key] = e
    json.dump(tj, open(path, "w"), indent=1)
    print(json.dumps({key: e}, indent=1))


if __name__ == "__main__":
    main()
