#!/usr/bin/env python
"""DRAM traffic of the dominant kernel of each bench config from `ncu --set full` reports, as
profiles/ncu_traffic.json (read by bench.py for `roofline.traffic`: per launch, at the bench size).
usage: python tools/ncu_traffic.py C2:gpurun_out/prof_tpp_C2_r2.ncu-rep:100000:100000 C4:...:problems:bench_problems"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
out = {}
try:  # entries of configs that are not re-captured stay as they are
    with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
        out = json.load(f)
except Exception:
    out = {}
for spec in sys.argv[1:]:
    name, rep, problems, bench_problems = spec.split(":")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        continue
    hdr, units, r = rows[0], rows[1], rows[2]

    def val(k):
        i = hdr.index(k)
        return float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0)
    out[name] = {"kernel": r[hdr.index("Kernel Name")].split("<")[0].split("(")[0].replace("void ", "").split("::")[-1],
                 "dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
                 "problems": int(problems), "bench_problems": int(bench_problems),
                 "source": f"ncu --set full --clock-control none, {os.path.basename(rep)} (profiles/*_ncu_full_summary.txt)"}
with open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out, indent=1))
