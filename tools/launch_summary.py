"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
h = rows[hdr]
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 2:]:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    k = r[ki].split("(")[0]
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(t for _, t in agg.values())
print(f"{'kernel':60s} {'launches':>8s} {'ms':>10s} {'share':>7s}")
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(f"{k[:60]:60s} {n:8d} {t / 1e6:10.2f} {t / tot:7.1%}")
print(f"{'total':60s} {sum(n for n, _ in agg.values()):8d} {tot / 1e6:10.2f}")
