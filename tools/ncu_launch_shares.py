"""Per-kernel launch count, average device time and share of an ncu launch list (--metrics gpu__time_duration.sum --csv)."""
import csv, sys
from collections import defaultdict
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5 and r[0].isdigit()]
d = defaultdict(list)
for r in rows:
    d[r[4][:70]].append(float(r[-1].replace(",", "")))
tot = sum(sum(v) for v in d.values()) or 1.0
for k, v in d.items():
    print(f"{k:72s} n={len(v):4d} avg {sum(v) / len(v) / 1e3:8.2f} us  share {sum(v) / tot:.3f}")
