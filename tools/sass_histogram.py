"""Instruction histogram per kernel of the built library (cuobjdump -sass, no GPU needed) -> profiles/<name>.json.

    python tools/sass_histogram.py profiles/r02_sass_histogram.json

Shows which tensor path each kernel uses: DMMA (FP64 mma.sync), UTCIMMA (tcgen05.mma kind::i8), UTMALDG (TMA tensor loads),
LDTM (tcgen05.ld), LDGSTS (cp.async), SYNCS (mbarrier).
"""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "pygpr_b200", "libpygpr_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
res, cur = {}, None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        cur = res.setdefault(name, collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
    if m and cur is not None:
        cur[m.group(1)] += 1
KEY = ["DMMA", "UTCIMMA", "UTCHMMA", "UTMALDG", "LDTM", "LDGSTS", "SYNCS", "UBLKCP", "HMMA", "IMMA", "DFMA", "MUFU", "I2F", "LDS", "STS", "LDG", "STG", "BAR"]
summary = {}
for k, c in sorted(res.items()):
    summary[k] = {"instructions": sum(c.values()), **{n: c[n] for n in KEY if c[n]}}
json.dump({"library": "pygpr_b200/libpygpr_b200.so", "source": "cuobjdump -sass", "kernels": summary}, open(sys.argv[1], "w"), indent=1)
print("kernels:", len(summary))
