#!/usr/bin/env python
"""Per-kernel totals and shares from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = OrderedDict()
for r in rows[1:]:
    if r[mi] != "gpu__time_duration.sum":
        continue
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    ms = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v
    name = r[ki].split("(")[0]
    n, t = tot.get(name, (0, 0.0))
    tot[name] = (n + 1, t + ms)
total = sum(t for _, t in tot.values())
print("# " + (sys.argv[2] if len(sys.argv) > 2 else ""))
print("# (times are cold-cache / serialised under ncu: compare SHARES, not absolutes)")
print("# kernel, launches, total_ms, share")
for name, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{name}, {n}, {t:.3f}, {t / total:.3f}")
