#!/usr/bin/env python
"""profiles/<tag>_traffic.json from an ncu capture of the launches of ONE step of tpobs_grid_kernel (one launch in the
plain form, two — near pass and rest pass — in the near-first form): DRAM bytes and warp instructions summed over the
step's launches, issue/occupancy figures of its longest launch.   python tools/make_traffic.py <rep> <out.json>"""
import csv
import io
import json
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units, data = rows[0], rows[1], rows[2:]


def col(r, name):
    v = float(r[hdr.index(name)].replace(",", ""))
    u = units[hdr.index(name)]
    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "msecond": 1e3, "usecond": 1.0, "nsecond": 1e-3, "second": 1e6}
    return v * scale.get(u, 1.0)


launches = []
for r in data:
    launches.append({
        "kernel": r[hdr.index("Kernel Name")][:120],
        "duration_us_under_ncu": col(r, "gpu__time_duration.sum"),
        "dram_bytes_read": col(r, "dram__bytes_read.sum"),
        "dram_bytes_write": col(r, "dram__bytes_write.sum"),
        "warp_instructions": col(r, "smsp__inst_executed.sum"),
        "issue_active_pct": col(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "warps_active_pct": col(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        "registers_per_thread": col(r, "launch__registers_per_thread"),
        "grid_size": col(r, "launch__grid_size"),
    })
longest = max(launches, key=lambda l: l["duration_us_under_ncu"])
json.dump({
    "source": "%s (ncu --set full --clock-control none on `bench.py --no-planner --no-cpu-baseline --no-structured --steps 3 --warmup 3`, the launches of one timed 4096-node step)" % rep,
    "launches": launches,
    "tpobs_grid_kernel": {
        "dram_bytes_read": sum(l["dram_bytes_read"] for l in launches),
        "dram_bytes_write": sum(l["dram_bytes_write"] for l in launches),
        "duration_us_under_ncu": sum(l["duration_us_under_ncu"] for l in launches),
        "warp_instructions": sum(l["warp_instructions"] for l in launches),
        "issue_active_pct": longest["issue_active_pct"],
        "warps_active_pct": longest["warps_active_pct"],
    },
}, open(out, "w"), indent=1)
print(open(out).read())
