"""Per-kernel share of the device time in an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: python tools/launch_shares.py gpurun_out/launches.csv"""
import csv
import io
import sys
from collections import defaultdict

lines = open(sys.argv[1]).read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
acc = defaultdict(lambda: [0, 0.0])
for r in rows:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    k = r["Kernel Name"][:90]
    acc[k][0] += 1
    acc[k][1] += float(r["Metric Value"].replace(",", ""))
tot = sum(v[1] for v in acc.values()) or 1.0
print(f"{len(rows)} launches, {tot / 1e6:.3f} ms of device time (cold-cache, serialised: compare shares)")
for k, v in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[0]:5d} launches {v[1] / 1e3:12.1f} us total {v[1] / v[0] / 1e3:10.1f} us avg {100 * v[1] / tot:6.2f}%  {k}")
