"""Device time of the factorisation: FP64 DMMA Cholesky (engine.fit) vs the blocked Cholesky with emulated-FP64 panel solves and
trailing updates (engine.fit_emulated) at config-4 expert sizes, shared communication block.

    python tools/emulated_fit_bench.py [experts=32] [slices=7]
"""
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pygpr_b200 import engine as eng  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 32
S = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 7
m0, d = 4096, 16
g = torch.Generator().manual_seed(0)
xg = torch.rand(1, m0, d, generator=g, dtype=torch.float64)
yg = torch.sin(xg.sum(-1)) + 0.1 * torch.randn(1, m0, generator=g, dtype=torch.float64)
xl = torch.rand(B, m0, d, generator=g, dtype=torch.float64)
yl = torch.sin(xl.sum(-1)) + 0.1 * torch.randn(B, m0, generator=g, dtype=torch.float64)
x = torch.cat((xg.expand(B, -1, -1), xl), 1).cuda()
y = torch.cat((yg.expand(B, -1), yl), 1).cuda()
hp = torch.empty(B, d + 2, dtype=torch.float64)
hp[:, 0] = 1.0; hp[:, 1:d + 1] = 2.0 / math.sqrt(d); hp[:, d + 1] = 0.3
hp = hp.cuda()
st_g = eng.fit(xg.cuda(), yg.cuda(), hp[:1], True)
eng.tril_inverse(st_g)
Lbuf = torch.empty(B, 2 * m0, 2 * m0, dtype=torch.float64, device="cuda")


def timed(fn, reps=2):
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, r


flops = B * ((2.0 * m0) ** 3 - float(m0) ** 3) / 3.0
t0, st0 = timed(lambda: eng.fit(x, y, hp, True, L_out=Lbuf, shared=st_g))
ref = {"L": torch.tril(st0.L).clone(), "nlml": st0.nlml().clone(), "alpha": st0.alpha.clone()}
print(json.dumps({"impl": "fp64_dmma", "experts": B, "ms": t0, "tflops": flops / t0 * 1e-9}), flush=True)
CONFIGS = ((1024, 32),) if "one" in sys.argv else ((1024, 32), (2048, 32), (1024, 16), (512, 32))
for panel, chunk in CONFIGS:
    t1, st1 = timed(lambda: eng.fit_emulated(x, y, hp, True, slices=S, L_out=Lbuf, shared=st_g, panel=panel, chunk=chunk))
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
    print(json.dumps({"impl": "emulated", "slices": S, "panel": panel, "chunk": chunk, "ms": t1, "tflops_equiv": flops / t1 * 1e-9,
                      "speedup": t0 / t1, "L_vs_fp64": rel(torch.tril(st1.L), ref["L"]), "nlml_vs_fp64": rel(st1.nlml(), ref["nlml"]),
                      "alpha_vs_fp64": rel(st1.alpha, ref["alpha"])}), flush=True)
