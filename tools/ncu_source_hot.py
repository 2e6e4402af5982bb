#!/usr/bin/env python3
"""Hot SASS ranges of one kernel from `ncu --page source --csv`:  python tools/ncu_source_hot.py file.csv [top]
Prints total executed warp instructions, then the SASS listing annotated with executed count / avg threads / stall samples."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr) and r[ix["Instructions Executed"]].replace(",", "").isdigit()]
tot = sum(int(r[ix["Instructions Executed"]] or 0) for r in data)
samp = sum(int(r[ix["# Samples"]] or 0) for r in data)
print("total warp instructions", tot, "samples", samp, "sass lines", len(data))
mode = sys.argv[2] if len(sys.argv) > 2 else "list"
if mode == "list":
    for r in data:
        n = int(r[ix["Instructions Executed"]] or 0)
        if n * 400 < tot:
            continue
        print(f"{r[ix['Address']][-5:]} {100.0 * n / tot:5.2f}% thr {float(r[ix['Avg. Threads Executed']] or 0):5.1f} smp {100.0 * int(r[ix['# Samples']] or 0) / max(1, samp):5.2f}%  {r[ix['Source']][:90]}")
