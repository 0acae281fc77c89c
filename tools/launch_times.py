"""Per-kernel times from an `ncu --metrics gpu__time_duration.sum --csv` launch list (ms, last launches)."""
import csv, sys
from collections import defaultdict
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
d = defaultdict(list)
for r in rows[1:]:
    d[r[ki].split("(")[0].replace("<unnamed>::", "")].append(float(r[vi].replace(",", "")) / 1e6)
tot = 0.0
for k, v in d.items():
    print(f"{k:28s} x{len(v):3d}  last {v[-1]:8.3f} ms   min {min(v):8.3f}")
    tot += v[-1]
print(f"{'sum of last launches':28s}       {tot:8.3f} ms")
