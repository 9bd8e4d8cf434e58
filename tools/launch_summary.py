#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X`).

    python tools/launch_summary.py gpurun_out/fit_launches.csv [skip_first_n_launches]
"""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
for r in rd:
    v = float(r[iv].replace(",", ""))
    u = r[iu]
    us = v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3)
    rows.append((re.sub(r"\(.*", "", r[ik]), us))
rows = rows[skip:]
tot = defaultdict(lambda: [0, 0.0])
for k, us in rows:
    tot[k][0] += 1
    tot[k][1] += us
total = sum(v[1] for v in tot.values())
print(f"# {len(rows)} launches, {total / 1e3:.3f} ms of kernel time (cold-cache, serialised by ncu)")
for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{us / 1e3:9.3f} ms  {100 * us / total:5.1f} %  {n:5d} x {us / n:8.1f} us  {k}")
