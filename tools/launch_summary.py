#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and shares."""
import collections
import csv
import re
import sys

path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
rows = [r for r in csv.DictReader(lines) if r.get("Metric Name") == "gpu__time_duration.sum"]
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    k = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
    k = re.sub(r"^at::.*", "torch (synthetic data generation / plumbing copies)", k)
    v = float(r["Metric Value"].replace(",", ""))
    tot[k][0] += 1
    tot[k][1] += v / 1e3 if r["Metric Unit"] in ("ns", "nsecond") else v
s = sum(v[1] for v in tot.values())
print(f"# {len(rows)} launches, {s / 1e3:.2f} ms of device time (per-launch times are cold-cache and serialised: compare shares)")
print(f"{'ms':>10s} {'share':>6s} {'launches':>8s}  kernel")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1] / 1e3:10.3f} {100 * v[1] / s:5.1f}% {v[0]:8d}  {k[:110]}")
