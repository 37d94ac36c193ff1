#!/usr/bin/env python
"""Aggregate an ncu source-page CSV into segments of equal execution count (hot-spot listing)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[1], rows[2:]
ia, ie, isamp, ith = (hdr.index(n) for n in ("Source", "Instructions Executed", "# Samples", "Avg. Threads Executed"))
tot = sum(int(r[ie]) for r in data)
ts = sum(int(r[isamp]) for r in data)
print("total warp-instructions", tot, "samples", ts)
prev, start, segs = None, 0, []
for idx, r in enumerate(data):
    e = int(r[ie])
    if prev is None or abs(e - prev) > 0.15 * max(prev, 1):
        if prev is not None:
            segs.append((start, idx - 1, prev))
        start = idx
    prev = e
segs.append((start, len(data) - 1, prev))
for a, b, e in segs:
    n = b - a + 1
    smp = sum(int(data[i][isamp]) for i in range(a, b + 1))
    if e * n > tot * 0.01 or smp > ts * 0.02:
        ops = collections.Counter(
            (data[i][ia].split()[1] if data[i][ia].strip().startswith("@") else data[i][ia].split()[0])
            for i in range(a, b + 1))
        print("sass[%d:%d] n=%d exec=%d instr=%.1f%% samples=%.1f%% thr=%s %s" % (
            a, b, n, e, 100 * e * n / tot, 100 * smp / ts, data[a][ith], dict(ops.most_common(6))))
for idx, r in enumerate(data):
    s = int(r[isamp])
    if s > ts * 0.02:
        print("  hot %d %-60s %.1f%% of samples" % (idx, r[ia].strip()[:60], 100 * s / ts))
