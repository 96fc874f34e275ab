"""ncu case for the dependency chain of a single-matrix factorisation: engine.fit of ONE matrix of order n (default 4096), twice.

    python tools/profile_fit1.py [n]
    ncu --set full --import-source on --clock-control none -k regex:potrf_diag_kernel -s 45 -c 1 -o gpurun_out/diag python tools/profile_fit1.py
"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pygpr_b200 import engine as eng  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
d = 8
g = torch.Generator().manual_seed(0)
x = torch.rand(1, n, d, generator=g, dtype=torch.float64).cuda()
y = torch.sin(x.sum(-1)) + 0.1 * torch.randn(1, n, generator=g, dtype=torch.float64).cuda()
hp = torch.empty(1, d + 2, dtype=torch.float64)
hp[:, 0] = 1.0; hp[:, 1:d + 1] = 2.0 / math.sqrt(d); hp[:, d + 1] = 0.3
hp = hp.cuda()
for it in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); st = eng.fit(x, y, hp, True); e1.record(); torch.cuda.synchronize()
    print("fit n=%d B=1: %.3f ms, %.2f TFLOP/s, nlml %.9g" % (n, e0.elapsed_time(e1), n ** 3 / 3 / e0.elapsed_time(e1) * 1e-9, float(st.nlml()[0])))
