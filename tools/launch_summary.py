"""Per-kernel summary of an ncu launch list (ncu --metrics gpu__time_duration.sum --csv --log-file X.csv <cmd>):
    python tools/launch_summary.py gpurun_out/launches.csv > profiles/launches.md"""
import csv
import sys
from collections import OrderedDict

rows = []
with open(sys.argv[1], newline="") as fh:
    lines = [l for l in fh if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
    rows.append((r["Kernel Name"], us))
agg = OrderedDict()
for name, us in rows:
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += us
total = sum(a[1] for a in agg.values())
print(f"{len(rows)} launches, {total:.1f} us of kernel time (serialised, cold-cache: compare SHARES)\n")
print("| kernel | launches | avg us | share |\n|---|---|---|---|")
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{name[:90]}` | {n} | {t / n:.1f} | {100 * t / total:.1f} % |")
