#!/usr/bin/env python3
"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list (B200_PROFILING.md recipe)."""
import collections
import csv
import re
import sys

lines = [ln for ln in open(sys.argv[1]) if not ln.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    v = v / 1000 if row["Metric Unit"] == "ns" else (v * 1000 if row["Metric Unit"] == "ms" else v)
    agg.setdefault(re.sub(r"\(.*", "", row["Kernel Name"])[:72], []).append(v)
tot = sum(sum(v) for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:72s} n={len(v):3d} avg={sum(v) / len(v):9.1f} us  share={100 * sum(v) / tot:5.1f} %")
