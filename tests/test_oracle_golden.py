"""Pin the CPU oracle to outputs of the unmodified reference (fixtures made by oracle/gen_golden.py)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import pygpr_oracle as orc
from helpers import assert_close

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
T = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float64))

COVAR = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "covar_*.npz")))
GP = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "gp_*.npz")))
GRBCM = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "grbcm_*.npz")))
SPECS = {"se": ["se"], "se_wn": ["se", "wn"], "se_se_wn": ["se", "se", "wn"]}


def test_fixtures_present():
    assert len(COVAR) >= 3 and len(GP) >= 5 and len(GRBCM) >= 3


@pytest.mark.parametrize("name", COVAR)
@pytest.mark.parametrize("tag", list(SPECS))
def test_covar(golden, name, tag):
    g = golden(name)
    x, xs, hp = T(g["x"]), T(g["xs"]), T(g[tag + "_hp"])
    B = x.shape[0]
    K = orc.compose_kernel(SPECS[tag], hp, x)
    Ks = orc.compose_kernel(SPECS[tag], hp, x, xs)
    assert_close(K.numpy().reshape(g[tag + "_K"].shape), g[tag + "_K"], 1e-13, "K")
    assert_close(Ks.numpy().reshape(g[tag + "_Ks"].shape), g[tag + "_Ks"], 1e-13, "K*")
    if tag + "_dK" in g:
        _, dK = orc.compose_kernel_and_grad(SPECS[tag], hp, x)
        assert_close(dK.numpy().reshape(g[tag + "_dK"].shape), g[tag + "_dK"], 1e-13, "dK")


@pytest.mark.parametrize("name", GP)
def test_gp(golden, name):
    g = golden(name)
    x, y, xs, hp = T(g["x"]), T(g["y"]), T(g["xs"]), T(g["hp"])
    spec = ["se", "wn"]
    # ill-conditioned default hyper-parameters: tolerance = the reference's own noise floor
    tol = 1e-4 if name.endswith("default") else 1e-9
    L, alpha = orc.gp_fit(spec, hp, x, y)
    mean, var = orc.gp_predict(spec, hp, x, y, xs, "diag", fit=(L, alpha))
    assert_close(alpha.numpy().reshape(g["alpha"].shape), g["alpha"], tol, "alpha")
    assert_close(mean.numpy().reshape(g["mean"].shape), g["mean"], tol, "mean")
    assert_close(var.numpy().reshape(g["var"].shape), g["var"], tol, "var")
    if g["L"].size:
        assert_close(L.numpy().reshape(g["L"].shape), g["L"], 1e-11, "L")
    if "cov_full" in g:
        _, cf = orc.gp_predict(spec, hp, x, y, xs, "full", fit=(L, alpha))
        assert_close(cf.numpy().reshape(g["cov_full"].shape), g["cov_full"], tol, "cov_full")
    nlml = orc.mle_loss(spec, hp, x, y)
    assert_close(nlml.numpy().reshape(g["nlml"].shape), g["nlml"], 1e-11, "nlml")
    nl2, grad = orc.mle_loss_and_grad(spec, hp, x, y)
    assert_close(nl2.numpy().reshape(g["nlml2"].shape), g["nlml2"], 1e-11, "nlml2")
    assert_close(grad.numpy().reshape(g["grad"].shape), g["grad"], tol, "grad")
    assert_close(g["grad_only"], g["grad"], tol, "reference grad vs loss_and_grad")
    if not name.endswith("default"):
        gamma = orc.learn_rate(spec, hp, x, y, 1e-4)
        assert_close(gamma.numpy().reshape(g["gamma"].shape), g["gamma"], 1e-5, "gamma")


@pytest.mark.parametrize("name", GRBCM)
def test_grbcm(golden, name):
    g = golden(name)
    a = {k: T(v) for k, v in g.items()}
    spec = ["se", "wn"]
    mu, var, parts = orc.grbcm_predict(spec, a["hp_g"], a["hp_l"], a["xl"], a["yl"], a["xg"], a["yg"], a["xs"])
    assert_close(mu.numpy(), g["mu"], 1e-9, "mu_A")
    assert_close(var.numpy(), g["var"], 1e-9, "var_A")
    assert_close(parts["beta"].numpy(), g["beta"], 1e-9, "beta")
    assert_close(parts["prec"].numpy(), g["prec"], 1e-9, "prec")
    # chunked expert loop is arithmetic-identical per expert
    mu2, var2, _ = orc.grbcm_predict(spec, a["hp_g"], a["hp_l"], a["xl"], a["yl"], a["xg"], a["yg"], a["xs"],
                                     expert_chunk=2)
    assert_close(mu2.numpy(), g["mu"], 1e-9, "mu_A chunked")
    # full-covariance path (next-row N1)
    x, y = orc.grbcm_inputs(a["xl"], a["yl"], a["xg"], a["yg"])
    mg, cg = orc.gp_predict(spec, a["hp_g"], a["xg"][None], a["yg"][None], a["xs"], "full")
    ml, cl = orc.gp_predict(spec, a["hp_l"], x, y, a["xs"], "full")
    muf, covf = orc.grbcm_aggregate_full(mg[0], cg[0], ml, cl)
    assert_close(covf.numpy(), g["cov_full"], 1e-7, "cov_full_A")
    assert_close(muf.numpy(), g["mu_full"], 1e-7, "mu_full_A")
