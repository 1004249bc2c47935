"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total ms and share per kernel."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
h, data = rows[hdr], rows[hdr + 1:]
ik, iv, iu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
skip = sys.argv[2].split(",") if len(sys.argv) > 2 else []
agg = collections.OrderedDict()
for r in data:
    n = r[ik].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
    if any(s and s in n for s in skip):
        continue
    v = float(r[iv].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu], 1.0)
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print("| kernel | launches | total ms (ncu, cold, serialised) | share |")
print("|---|---:|---:|---:|")
for n, a in agg.items():
    print(f"| `{n}` | {a[0]} | {a[1]:.3f} | {100 * a[1] / tot:.1f}% |")
print(f"| **total** | {sum(a[0] for a in agg.values())} | {tot:.3f} | 100% |")
