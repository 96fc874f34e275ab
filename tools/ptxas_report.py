"""Registers / spills / shared memory per kernel from `nvcc -Xptxas -v` (rebuilds the library; no GPU needed).

    python tools/ptxas_report.py profiles/r02_ptxas_resources.json
"""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pygpr_b200.csrc import build as b  # noqa: E402

cmd = [os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")] + b.NVCC_FLAGS + ["-Xptxas", "-v", "-o", b.OUT] + b.SOURCES
err = subprocess.run(cmd, capture_output=True, text=True).stderr
out, cur = {}, None
for line in err.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        out[cur] = {}
        continue
    if cur is None:
        continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m:
        out[cur].update(stack_bytes=int(m.group(1)), spill_store_bytes=int(m.group(2)), spill_load_bytes=int(m.group(3)))
    m = re.search(r"Used (\d+) registers", line)
    if m:
        out[cur]["registers"] = int(m.group(1))
        s = re.search(r"(\d+) bytes smem", line)
        out[cur]["static_smem_bytes"] = int(s.group(1)) if s else 0
json.dump({"source": "nvcc -Xptxas -v, sm_100a, " + " ".join(b.NVCC_FLAGS), "kernels": dict(sorted(out.items()))}, open(sys.argv[1], "w"), indent=1)
for k, v in sorted(out.items()):
    print(k[:50].ljust(50), v)
