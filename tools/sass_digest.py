#!/usr/bin/env python
"""Per-kernel digest of the built library's SASS: the opcodes that prove tcgen05 / TMEM / TMA / FFMA2 /
DFMA are there, plus `cuobjdump -res-usage` (registers, shared memory, spills).  CPU only (cuobjdump).
usage: python tools/sass_digest.py [tinyopt_b200/libtinyopt_b200.so] > profiles/r2_sass_digest.txt"""
import collections
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else "tinyopt_b200/libtinyopt_b200.so"
OPS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMAPF", "UTMASTG", "UBLKCP", "SYNCS", "FFMA2", "FFMA", "DFMA",
       "HMMA", "LDGSTS", "LDS", "STS", "SHFL", "BAR", "ATOM", "MUFU"]
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
counts, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m:
        op = m.group(1)
        counts[cur]["_total"] += 1
        for o in OPS:
            if op == o or op.startswith(o + "."):
                counts[cur][o] += 1
                break
res = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True).stdout
usage, fn = {}, None
for line in res.splitlines():
    m = re.match(r"\s*Function (\S+):", line)
    if m:
        fn = m.group(1)
        continue
    if fn and "REG:" in line:
        usage[fn] = " ".join(line.split())
        fn = None
dem = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
print(f"# SASS digest of {so}: {len(counts)} kernels, {tot['_total']} instructions")
print("# whole library: " + ", ".join(f"{o} {tot[o]}" for o in OPS if tot[o]))
print("# per kernel (only opcodes that occur); res-usage from cuobjdump -res-usage")
for (mangled, c), name in zip(counts.items(), dem):
    short = re.sub(r"\(.*", "", name)
    ops = ", ".join(f"{o} {c[o]}" for o in OPS if c[o])
    print(f"{short}\n    instr {c['_total']}: {ops}\n    {usage.get(mangled, '')}")
