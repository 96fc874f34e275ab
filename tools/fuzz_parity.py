"""Randomised parity sweep on the GPU: GRBCM / Exact_GP / MLE through the public API against the CPU oracle over ragged
shapes (n, d, m not multiples of the tile sizes; 1..9 experts; aligned and unaligned communication sets; shared and
per-expert hyper-parameters).  Prints the worst relative error per quantity; exits 1 on a violation of 1e-8."""
import math, os, sys
import numpy as np
import torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import pygpr_b200 as P
from oracle import pygpr_oracle as orc

TOL = 1e-8
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
ncases = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cov = P.Compose([P.Squared_exponential(), P.White_noise()])
worst = {}


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def note(k, v, ctx):
    if v > worst.get(k, (0, None))[0]:
        worst[k] = (v, ctx)


bad = 0
for case in range(ncases):
    g = torch.Generator().manual_seed(int(rng.integers(1 << 30)))
    d = int(rng.integers(1, 20))
    nc = int(rng.integers(2, 10))
    ng = int(rng.choice([128, 256, 384, 37, 100, 200, 130]))
    nls = int(rng.choice([1, 5, 64, 127, 128, 129, 200, 300]))
    m = int(rng.choice([2, 7, 128, 129, 250]))
    per = bool(rng.integers(0, 3) == 0)
    xl = torch.rand(nc, nls, d, generator=g, dtype=torch.float64)
    yl = torch.sin(xl.sum(-1)) + 0.1 * torch.randn(nc, nls, generator=g, dtype=torch.float64)
    xg = torch.rand(ng, d, generator=g, dtype=torch.float64)
    yg = torch.sin(xg.sum(-1)) + 0.1 * torch.randn(ng, generator=g, dtype=torch.float64)
    xs = torch.rand(m, d, generator=g, dtype=torch.float64)
    hp = torch.empty(nc, d + 2, dtype=torch.float64)
    hp[:, 0] = 1.0 + 0.2 * float(rng.random()); hp[:, 1:d + 1] = (1.0 + float(rng.random())) / math.sqrt(d); hp[:, d + 1] = 0.3
    if per:
        hp = hp * (0.8 + 0.4 * torch.rand(nc, d + 2, generator=g, dtype=torch.float64))
    hp_g = hp[0].clone()
    ctx = dict(case=case, d=d, nc=nc, ng=ng, nls=nls, m=m, per=per)
    model = P.GRBCM(xl, yl, xg, yg, cov)
    model.gpl.set_params(hp); model.gpg.set_params(hp_g)
    mu, var = model.predict(xs, var="diag")
    mu_o, var_o, _ = orc.grbcm_predict(["se", "wn"], hp_g, hp, xl, yl, xg, yg, xs)
    ctx["shared"] = model.gpl._state.shared_from is not None
    note("grbcm_mu", rel(mu, mu_o), ctx); note("grbcm_var", rel(var, var_o), ctx)
    x_all, y_all = orc.grbcm_inputs(xl, yl, xg, yg)
    l, j = P.MLE(model.gpl).loss_and_grad(hp.numpy())
    l_o, j_o = orc.mle_loss_and_grad(["se", "wn"], hp, x_all, y_all)
    note("nlml", rel(l, l_o.numpy()), ctx); note("grad", rel(j, j_o.numpy()), ctx)
    if not per:
        f, gr = P.GRBCM_MLE(model, chunk=int(rng.integers(0, 4))).loss_and_grad(hp_g.numpy())
        lg_o, jg_o = orc.mle_loss_and_grad(["se", "wn"], hp_g[None], xg[None], yg[None])
        note("grbcm_mle", abs(f - float(l_o.sum() + lg_o.sum())) / abs(float(l_o.sum() + lg_o.sum())), ctx)
        note("grbcm_mle_grad", rel(gr, (j_o.sum(0) + jg_o[0]).numpy()), ctx)
    ys, cs = model.gpl.predict(xs, var="diag")
    ys_o, cs_o = orc.gp_predict(["se", "wn"], hp, x_all, y_all, xs, "diag")
    note("gp_mean", rel(ys.reshape(nc, m), ys_o.reshape(nc, m)), ctx); note("gp_var", rel(cs.reshape(nc, m), cs_o.reshape(nc, m)), ctx)
for k, (v, ctx) in sorted(worst.items()):
    flag = "" if v <= TOL else "   <-- VIOLATION"
    bad += v > TOL
    print("%-16s worst rel err %.3e  at %s%s" % (k, v, ctx, flag))
print("fuzz_parity:", "FAILED" if bad else "ok", "(%d cases)" % ncases)
sys.exit(1 if bad else 0)
