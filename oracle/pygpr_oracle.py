"""CPU oracle for the PyGPR data-parallel hot path.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

This is a plain restatement, in float64 torch-CPU tensor ops, of the algorithm that the
reference (sarath-srinivas/PyGPR, mounted read-only at /root/reference in the build container)
runs for: covariance kernels, exact-GP fit/predict, the negative log marginal likelihood and its
hyper-parameter gradient, the quadratic learn-rate rule and GRBCM aggregation.  Every function
cites the reference file:line it follows.

Why torch-CPU and not C/numpy: the reference *is* pure Python over torch CPU ops; its arithmetic
lives in ATen -> MKL (dgemm / dpotrf / dpotrs), a third-party dependency that is not vendored in
/root/reference (setup.py:11 pins no versions).  Restating the formulas over the same BLAS/LAPACK
entry points keeps the oracle's rounding behaviour (and its wall-clock as the reported CPU
baseline) as close to the reference as a restatement can be.

Parity pinning: the reference ships NO golden vectors (its tests are property tests on unseeded
inputs), so this oracle is pinned against outputs of the unmodified reference itself, generated
in the build container by ``oracle/gen_golden.py`` and committed under ``tests/golden/``
(``tests/test_oracle_golden.py`` checks every fixture).  Parity status: PINNED (against the
reference run here on seeded inputs).

Differences from the reference that are deliberate:
  * explicit dtype=float64 everywhere instead of flipping the process default dtype
    (covar.py:5, gpr.py:10, ...);
  * no in-place mutation of caller tensors (e.g. gr_bcm.py:128 ``var_g.reciprocal_()``);
  * batch dimension is always explicit: x:[B,n,d], hp:[B,P]; the squeeze-away-batch-of-1
    convention (covar.py:162-166) lives in the API mirror, not here.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import torch

F64 = torch.float64
JITTER = 1e-7  # gpr.py:68, loss.py:38,63,96


def _b3(x: torch.Tensor) -> torch.Tensor:
    return x.reshape(-1, x.shape[-2], x.shape[-1]).to(F64)


def _b2(hp: torch.Tensor) -> torch.Tensor:
    return hp.reshape(-1, hp.shape[-1]).to(F64)


# --------------------------------------------------------------------------------------------
# covariance kernels
# --------------------------------------------------------------------------------------------
def se_distance(x: torch.Tensor, xp: Optional[torch.Tensor] = None) -> torch.Tensor:
    """GEMM-form squared distance, covar.py:102-127.  x:[B,n,d], xp:[B|1,m,d] -> [B,m,n]."""
    x = _b3(x)
    x2 = x.square().sum(2)
    if xp is None:
        # covar.py:110-112
        return -2.0 * torch.matmul(x, x.transpose(1, 2)) + (x2.unsqueeze(2) + x2.unsqueeze(1))
    xp = _b3(xp)
    xp2 = xp.square().sum(2)
    # covar.py:119-121: rows = test points, cols = training points
    return -2.0 * torch.matmul(xp, x.transpose(1, 2)) + (xp2.unsqueeze(2) + x2.unsqueeze(1))


def se_kernel(hp: torch.Tensor, x: torch.Tensor, xp: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Squared_exponential.kernel, covar.py:129-167.  hp:[B,d+1]=[sig, ls_1..ls_d]."""
    hp = _b2(hp)
    x = _b3(x)
    sig = hp[:, 0]
    ls = hp[:, 1:].unsqueeze(1)  # covar.py:137-139
    xl = x * ls                  # covar.py:140  (length-scales MULTIPLY, no 1/2 factor)
    if xp is None:
        sqd = se_distance(xl)
    else:
        xpl = _b3(xp) * ls       # covar.py:153
        sqd = se_distance(xl, xpl)
    # covar.py:147-149 / 158-160
    return torch.exp(-sqd) * sig[:, None, None].square()


def se_kernel_and_grad(hp: torch.Tensor, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Squared_exponential.kernel_and_grad, covar.py:169-206 -> K:[B,n,n], dK:[B,d+1,n,n]."""
    hp = _b2(hp)
    x = _b3(x)
    krn = se_kernel(hp, x)
    sig = hp[:, 0]
    ls = hp[:, 1:]
    B, n, d = x.shape
    dkrn = torch.empty(B, d + 1, n, n, dtype=F64)
    dkrn[:, 0] = krn * (2.0 / sig)[:, None, None]                   # covar.py:189
    xt = x.transpose(-2, -1)
    diff = (xt[:, :, :, None] - xt[:, :, None, :]).square()          # covar.py:191-194
    diff = diff * ls[:, :, None, None] * krn[:, None, :, :] * (-2.0)  # covar.py:195-197
    dkrn[:, 1:] = diff                                               # covar.py:199
    return krn, dkrn


def wn_kernel(hp: torch.Tensor, x: torch.Tensor, xp: Optional[torch.Tensor] = None):
    """White_noise.kernel, covar.py:227-245.  Cross-covariance is the scalar 0 (covar.py:243)."""
    if xp is not None:
        return 0.0
    hp = _b2(hp)
    n = x.shape[-2]
    sig_n = hp[:, 0]
    return torch.eye(n, dtype=F64)[None] * sig_n[:, None, None].square()  # covar.py:237-238


def wn_kernel_and_grad(hp: torch.Tensor, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """White_noise.kernel_and_grad, covar.py:247-269: dK = K * 2/sig_n = 2 sig_n I."""
    hp = _b2(hp)
    krn = wn_kernel(hp, x)
    sig_n = hp[:, 0]
    dkrn = (krn * (2.0 / sig_n)[:, None, None]).unsqueeze(1)  # covar.py:262-264
    return krn, dkrn


_KINDS = {"se": (se_kernel, se_kernel_and_grad), "wn": (wn_kernel, wn_kernel_and_grad)}


def n_params(kind: str, d: int) -> int:
    """get_params_shape(...)[-1]: covar.py:89-94 (SE: d+1), covar.py:214-219 (WN: 1)."""
    return d + 1 if kind == "se" else 1


def init_params(spec: Sequence[str], B: int, d: int) -> torch.Tensor:
    """init_params: SE all ones (covar.py:96-100), WN 1e-4 (covar.py:221-225), cat (covar.py:45-48)."""
    parts = []
    for kind in spec:
        v = 1.0 if kind == "se" else 1e-4
        parts.append(torch.full((B, n_params(kind, d)), v, dtype=F64))
    return torch.cat(parts, dim=-1)


def compose_kernel(spec: Sequence[str], hp: torch.Tensor, x: torch.Tensor,
                   xp: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Compose.kernel, covar.py:50-62: split hp by child widths, sum the children."""
    hp = _b2(hp)
    d = x.shape[-1]
    chunks = [n_params(k, d) for k in spec]
    assert hp.shape[-1] == sum(chunks)                      # covar.py:52
    parts = hp.split(chunks, dim=-1)                        # covar.py:54-55
    krn = _KINDS[spec[0]][0](parts[0], x, xp)               # covar.py:57
    for kind, p in zip(spec[1:], parts[1:]):
        krn = krn + _KINDS[kind][0](p, x, xp)               # covar.py:59-60
    return krn


def compose_kernel_and_grad(spec: Sequence[str], hp: torch.Tensor, x: torch.Tensor):
    """Compose.kernel_and_grad, covar.py:64-81: K = sum, dK = cat along the hp axis."""
    hp = _b2(hp)
    d = x.shape[-1]
    chunks = [n_params(k, d) for k in spec]
    assert hp.shape[-1] == sum(chunks)                      # covar.py:66
    parts = hp.split(chunks, dim=-1)
    krn = None
    dk = []
    for kind, p in zip(spec, parts):
        k_i, dk_i = _KINDS[kind][1](p, x)
        krn = k_i if krn is None else krn + k_i
        dk.append(dk_i)
    return krn, torch.cat(dk, dim=-3)                       # covar.py:79


# --------------------------------------------------------------------------------------------
# exact GP
# --------------------------------------------------------------------------------------------
def gp_fit(spec, hp, x, y) -> Tuple[torch.Tensor, torch.Tensor]:
    """Exact_GP.update, gpr.py:65-74 -> (L:[B,n,n] lower Cholesky of K+1e-7 I, alpha:[B,n])."""
    x = _b3(x)
    y = y.reshape(-1, y.shape[-1]).to(F64)
    krn = compose_kernel(spec, hp, x).clone()                                  # gpr.py:67
    krn.diagonal(dim1=-2, dim2=-1).add_(JITTER)                                # gpr.py:68
    L = torch.linalg.cholesky(krn)                                             # gpr.py:69
    alpha = torch.cholesky_solve(y[..., None], L).squeeze(-1)                  # gpr.py:70-72
    return L, alpha


def gp_predict(spec, hp, x, y, xp, var: str = "diag", fit=None):
    """Exact_GP.predict / predict_var / predict_covar, gpr.py:76-120.

    Returns (mean:[B,m], cov) with cov = [B,m] for "diag", [B,m,m] for "full", None otherwise.
    """
    x = _b3(x)
    L, alpha = gp_fit(spec, hp, x, y) if fit is None else fit
    krns = compose_kernel(spec, hp, x, xp)                                     # gpr.py:79  [B,m,n]
    mean = torch.bmm(krns, alpha[..., None]).squeeze(-1)                       # gpr.py:80-85
    if var not in ("diag", "full"):
        return mean, None                                                      # gpr.py:91-92
    B = x.shape[0]
    xpb = _b3(xp).expand(B, -1, -1) if _b3(xp).shape[0] == 1 else _b3(xp)
    krnss = compose_kernel(spec, hp, xpb)                                      # gpr.py:98 (noise INCLUDED)
    lks = torch.cholesky_solve(krns.transpose(-2, -1), L)                      # gpr.py:100
    if var == "diag":
        v = krnss.diagonal(dim1=-2, dim2=-1) - (krns * lks.transpose(-2, -1)).sum(-1)  # gpr.py:102-104
        return mean, v
    return mean, krnss - torch.bmm(krns, lks)                                  # gpr.py:114-118


# --------------------------------------------------------------------------------------------
# negative log marginal likelihood and gradient
# --------------------------------------------------------------------------------------------
def mle_loss(spec, hp, x, y) -> torch.Tensor:
    """MLE.loss, loss.py:35-57 -> NLML:[B] = 1/2 y^T alpha + sum log L_ii + n/2 log 2 pi."""
    x = _b3(x)
    y = y.reshape(-1, y.shape[-1]).to(F64)
    L, alpha = gp_fit(spec, hp, x, y)                                          # loss.py:37-44
    n = y.shape[-1]
    return (0.5 * (alpha * y).sum(-1) + torch.log(L.diagonal(dim1=-2, dim2=-1)).sum(-1)
            + 0.5 * n * math.log(2.0 * math.pi))                               # loss.py:49-51


def mle_loss_and_grad(spec, hp, x, y) -> Tuple[torch.Tensor, torch.Tensor]:
    """MLE.loss_and_grad, loss.py:92-128 (grad alone: loss.py:59-90) -> ([B], [B,P]).

    Follows the reference formulation literally: dK is materialised ([B,P,n,n]) and
    tr(K^-1 dK_k) is taken from cholesky_solve(dK, L) (loss.py:116-119), so memory is
    O(P n^2) per expert.  Use ``mle_loss_and_grad_chunked`` for larger batches.
    """
    x = _b3(x)
    y = y.reshape(-1, y.shape[-1]).to(F64)
    krn, dkrn = compose_kernel_and_grad(spec, hp, x)                           # loss.py:94
    krn = krn.clone()
    krn.diagonal(dim1=-2, dim2=-1).add_(JITTER)                                # loss.py:96
    L = torch.linalg.cholesky(krn)                                             # loss.py:97
    alpha = torch.cholesky_solve(y[..., None], L).squeeze(-1)                  # loss.py:102
    wtb = alpha[:, :, None] * alpha[:, None, :]                                # loss.py:105
    n = y.shape[-1]
    nlml = (0.5 * (alpha * y).sum(-1) + torch.log(L.diagonal(dim1=-2, dim2=-1)).sum(-1)
            + 0.5 * n * math.log(2.0 * math.pi))                               # loss.py:107-109
    kk = torch.cholesky_solve(dkrn, L[:, None, :, :])                          # loss.py:116
    tr1 = (dkrn * wtb[:, None, :, :]).sum((-1, -2))                            # loss.py:118
    tr2 = kk.diagonal(dim1=-1, dim2=-2).sum(-1)                                # loss.py:119
    return nlml, -0.5 * (tr1 - tr2)                                            # loss.py:121


def mle_loss_and_grad_chunked(spec, hp, x, y, chunk: int = 1):
    """Same as mle_loss_and_grad, looping over expert chunks to bound memory."""
    x = _b3(x); hp = _b2(hp); y = y.reshape(-1, y.shape[-1])
    out_l, out_g = [], []
    for s in range(0, x.shape[0], chunk):
        l, g = mle_loss_and_grad(spec, hp[s:s + chunk], x[s:s + chunk], y[s:s + chunk])
        out_l.append(l); out_g.append(g)
    return torch.cat(out_l), torch.cat(out_g)


def learn_rate(spec, hp, x, y, eps: float) -> torch.Tensor:
    """get_learn_rate, hp_update.py:6-28 -> gamma:[B] from one loss_and_grad + two losses."""
    hp = _b2(hp)
    f0, J = mle_loss_and_grad(spec, hp, x, y)                                  # hp_update.py:14
    fp = mle_loss(spec, hp - eps * J, x, y)                                    # hp_update.py:16
    fm = mle_loss(spec, hp + eps * J, x, y)                                    # hp_update.py:17
    C1 = (fp - fm) / (2.0 * eps)                                               # hp_update.py:19
    C2 = (fp + fm - 2 * f0) / (2.0 * eps ** 2)                                 # hp_update.py:20
    return -0.5 * (C1 / C2)                                                    # hp_update.py:22


# --------------------------------------------------------------------------------------------
# GRBCM
# --------------------------------------------------------------------------------------------
def grbcm_inputs(xl, yl, xg, yg):
    """GRBCM.__init__, gr_bcm.py:19-26: communication set FIRST, then the local block."""
    nc = xl.shape[0]
    x = torch.cat((xg.to(F64)[None].expand(nc, -1, -1), xl.to(F64)), dim=1)
    y = torch.cat((yg.to(F64)[None].expand(nc, -1), yl.to(F64)), dim=1)
    return x, y


def grbcm_aggregate(mean_g, var_g, mean_l, var_l):
    """GRBCM.aggregate (var="diag"), gr_bcm.py:116-149.

    mean_g,var_g:[m]; mean_l,var_l:[nc,m] -> (mu_A:[m], var_A:[m], beta:[nc+1,m], prec:[nc+1,m]).
    """
    nc, m = mean_l.shape
    prec = torch.empty(nc + 1, m, dtype=F64)
    beta = torch.empty(nc + 1, m, dtype=F64)
    prec[0] = 1.0 / var_g                                                      # gr_bcm.py:128
    prec[1:] = 1.0 / var_l                                                     # gr_bcm.py:129
    beta[1:] = 0.5 * (torch.log(prec[1:]) - torch.log(prec[0]))                # gr_bcm.py:131
    beta[1] = 1.0                                                              # gr_bcm.py:132
    beta[0] = -(beta[1:].sum(0) - 1.0)                                         # gr_bcm.py:133
    ys = torch.cat((mean_g[None], mean_l))                                     # gr_bcm.py:138
    precs = prec * beta                                                        # gr_bcm.py:140
    var_a = 1.0 / precs.sum(0)                                                 # gr_bcm.py:143
    mu_a = (ys * precs).sum(0) * var_a                                         # gr_bcm.py:144
    return mu_a, var_a, beta, prec


def grbcm_aggregate_full(mean_g, cov_g, mean_l, cov_l):
    """GRBCM.aggregate with var="full" + aggregate_full_covar, gr_bcm.py:99-114,116-149."""
    var_g = torch.diag(cov_g)                                                  # gr_bcm.py:122
    var_l = torch.diagonal(cov_l, dim1=-2, dim2=-1)                            # gr_bcm.py:123
    _, _, beta, prec = grbcm_aggregate(mean_g, var_g, mean_l, var_l)
    covar_gl = torch.cat((cov_g[None], cov_l))                                 # gr_bcm.py:100
    chd = torch.linalg.cholesky(covar_gl)                                      # gr_bcm.py:101
    idt = torch.eye(covar_gl.shape[-1], dtype=F64).expand_as(covar_gl)
    prec_gl = torch.cholesky_solve(idt, chd)                                   # gr_bcm.py:103
    beta_covar = 0.5 * (beta[:, :, None] + beta[:, None, :])                   # gr_bcm.py:107-108
    p = (prec_gl * beta_covar).sum(0)                                          # gr_bcm.py:111
    covars = torch.cholesky_inverse(torch.linalg.cholesky(p))                  # gr_bcm.py:112
    ys = torch.cat((mean_g[None], mean_l))
    mu = (ys * (prec * beta)).sum(0) * torch.diag(covars)                      # gr_bcm.py:147
    return mu, covars


def grbcm_predict(spec, hp_g, hp_l, xl, yl, xg, yg, xs, expert_chunk: int = 0, fits=None):
    """GRBCM.predict (var="diag"), gr_bcm.py:151-155 -> (mu_A:[m], var_A:[m], parts dict).

    expert_chunk>0 loops the local experts in chunks (same arithmetic per expert; bounds memory).
    ``fits`` optionally carries cached (fit_g, [fit_l chunks]) so repeated predicts reuse the
    factorisations exactly like the reference's need_upd flag (gpr.py:65-74).
    """
    x, y = grbcm_inputs(xl, yl, xg, yg)
    nc = x.shape[0]
    hp_l = _b2(hp_l)
    fit_g = fits[0] if fits is not None else None
    mean_g, var_g = gp_predict(spec, hp_g, xg[None], yg[None], xs, "diag", fit=fit_g)  # gr_bcm.py:152
    step = expert_chunk if expert_chunk > 0 else nc
    means, vars_ = [], []
    for ci, s in enumerate(range(0, nc, step)):
        f = fits[1][ci] if fits is not None else None
        mu, v = gp_predict(spec, hp_l[s:s + step], x[s:s + step], y[s:s + step], xs, "diag", fit=f)  # gr_bcm.py:153
        means.append(mu); vars_.append(v)
    mean_l = torch.cat(means); var_l = torch.cat(vars_)
    mu_a, var_a, beta, prec = grbcm_aggregate(mean_g[0], var_g[0], mean_l, var_l)  # gr_bcm.py:155
    return mu_a, var_a, {"mean_g": mean_g[0], "var_g": var_g[0], "mean_l": mean_l, "var_l": var_l,
                         "beta": beta, "prec": prec}


def grbcm_fit(spec, hp_g, hp_l, xl, yl, xg, yg, expert_chunk: int = 0):
    """The two Exact_GP.update() calls GRBCM.predict triggers on first use (gr_bcm.py:152-153)."""
    x, y = grbcm_inputs(xl, yl, xg, yg)
    nc = x.shape[0]
    hp_l = _b2(hp_l)
    fit_g = gp_fit(spec, hp_g, xg[None], yg[None])
    step = expert_chunk if expert_chunk > 0 else nc
    fit_l = [gp_fit(spec, hp_l[s:s + step], x[s:s + step], y[s:s + step]) for s in range(0, nc, step)]
    return fit_g, fit_l
