"""Stage-by-stage device time of the emulated-FP64 prediction (engine.predict_shared_emulated) at config-4 expert sizes.

    python tools/emulated_breakdown.py [experts=16] [slices=6] [chunk=8]
"""
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pygpr_b200 as P  # noqa: E402
from pygpr_b200 import _lib, engine as eng  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
S = int(sys.argv[2]) if len(sys.argv) > 2 else 6
E = int(sys.argv[3]) if len(sys.argv) > 3 else 8
m0, d, m = 4096, 16, 4096
g = torch.Generator().manual_seed(0)
x = torch.rand(B + 1, m0, d, generator=g, dtype=torch.float64)
y = torch.sin(x.sum(-1)) + 0.1 * torch.randn(B + 1, m0, generator=g, dtype=torch.float64)
xs = torch.rand(m, d, generator=g, dtype=torch.float64).cuda()
hp = torch.empty(B, d + 2, dtype=torch.float64)
hp[:, 0] = 1.0; hp[:, 1:d + 1] = 2.0 / math.sqrt(d); hp[:, d + 1] = 0.3
cov = P.Compose([P.Squared_exponential(), P.White_noise()])
model = P.GRBCM(x[1:], y[1:], x[0], y[0], cov)
model.gpl.set_params(hp); model.gpg.set_params(hp[0])
model.update()
st_g, st_l = model.gpg._state, model.gpl._state
eng.ensure_inverse_factors(st_g, st_l)
st_l.plane_cache = None
ng, n, npad = st_g.n, st_l.n, st_l.npad
nl = n - ng


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


as_, nrms = eng.scale_inputs(xs, st_l.hp[:1], 1, m)
as_e, nrms_e = as_.expand(E, -1, -1).contiguous(), nrms.expand(E, -1).contiguous()
Ks = torch.empty(E, m, n, dtype=torch.float64, device="cuda")
X = torch.empty(E, nl, npad, dtype=torch.float64, device="cuda")
V = torch.empty(E, m, nl, dtype=torch.float64, device="cuda")
ml = torch.empty(E, m, dtype=torch.float64, device="cuda"); vl = torch.empty_like(ml)
mg = torch.zeros(1, m, dtype=torch.float64, device="cuda"); vg = torch.ones(1, m, dtype=torch.float64, device="cuda")
_p, _s = eng._p, eng._stream
out = {"experts_per_chunk": E, "slices": S}
out["se_cross_ms"] = timed(lambda: _lib.call("gpr_se_cross", _p(as_e), _p(nrms_e), m, _p(st_l.a[:E]), _p(st_l.nrm[:E]), npad, _p(st_l.hp[:E]),
                                              st_l.hp.shape[-1], E, m, n, st_l.dpad, _p(Ks), _s()))
out["split_Ks_ms"] = timed(lambda: eng.ozaki_split(Ks.reshape(E * m, n), S))
pf = torch.empty(S, E * m, npad, dtype=torch.int8, device="cuda"); sf = torch.empty(E * m, dtype=torch.float64, device="cuda")
out["fused_cross_planes_ms (replaces the two above)"] = timed(lambda: _lib.call(
    "gpr_se_cross_planes", _p(as_e), _p(nrms_e), m, _p(st_l.a[:E]), _p(st_l.nrm[:E]), npad, _p(st_l.hp[:E]), st_l.hp.shape[-1], E, n,
    st_l.dpad, S, 1.0, _p(pf), _p(sf), _s()))
out["inverse_rows_ms"] = timed(lambda: _lib.call("gpr_tril_inverse_rows", _p(st_l.L[:E]), _p(st_l.invdiag[:E]), E, npad, ng, _p(X), _s()))
out["split_X_ms"] = timed(lambda: eng.ozaki_split(X.reshape(E * nl, npad), S))
pa, pb = eng.ozaki_split(Ks.reshape(E * m, n), S), eng.ozaki_split(X.reshape(E * nl, npad), S)
out["gemm_ms"] = timed(lambda: eng.ozaki_gemm_nt(pa, pb, out=V, b_lower=ng, batch=E))
out["reduce_ms"] = timed(lambda: _lib.call("gpr_vt_reduce", _p(V), nl, _p(st_l.z[:E, ng:]), st_l.z.stride(0), _p(mg), _p(vg), E, m, nl, _p(ml), _p(vl), _s()))
out["sum_ms"] = sum(v for k, v in out.items() if k.endswith("_ms") and k not in ("se_cross_ms", "split_Ks_ms"))
out["per_expert_ms"] = out["sum_ms"] / E
flops = E * (float(n) ** 2 - float(ng) ** 2) * m
out["gemm_tflops_equiv"] = flops / out["gemm_ms"] * 1e-9
st_l.plane_cache = None
out["predict_shared_emulated_ms_B%d" % B] = timed(lambda: eng.predict_shared_emulated(st_g, st_l, xs, S, E))
out["predict_shared_fp64_ms_B%d" % B] = timed(lambda: eng.predict_shared(st_g, st_l, xs))
print(json.dumps(out))
