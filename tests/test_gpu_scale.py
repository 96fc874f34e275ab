"""BASELINE-size checks of the CUDA path (SURVEY.md §8d sizes): the CPU oracle would need minutes per case here, so
these use size-independent properties and, as an independent second opinion, the vendor FP64 routines on the same GPU
(torch.linalg.cholesky / cholesky_solve -> cuSOLVER/cuBLAS; test-side checker only, never on the product path).
Tolerance: relative 1e-8 (north_star) on mean, variance, NLML."""
import math

import pytest
import torch

from helpers import assert_close

pytestmark = pytest.mark.gpu
TOL = 1e-8


def _data(seed, B, n, d, m):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, n, d, generator=g, dtype=torch.float64)
    y = torch.sin(x.sum(-1)) + 0.1 * torch.randn(B, n, generator=g, dtype=torch.float64)
    xs = torch.rand(m, d, generator=g, dtype=torch.float64)
    hp = torch.empty(B, d + 2, dtype=torch.float64)
    hp[:, 0] = 1.0
    hp[:, 1:d + 1] = 2.0 / math.sqrt(d)
    hp[:, d + 1] = 0.3
    return x.cuda(), y.cuda(), xs.cuda(), hp.cuda()


def _dense_reference(x, y, xs, hp):
    """K = sig^2 exp(-|a_i-a_j|^2) + (sig_n^2 + 1e-7) I by direct differences (not the GEMM form), then library
    Cholesky / solves: covar.py:137-149,237-238, gpr.py:65-106, loss.py:49-51."""
    B, n, d = x.shape
    sig2 = hp[:, 0] ** 2
    sn2 = hp[:, d + 1] ** 2
    a = x * hp[:, None, 1:d + 1]
    as_ = xs[None] * hp[:, None, 1:d + 1]
    K = torch.cdist(a, a).square_().neg_().exp_().mul_(sig2[:, None, None])
    K.diagonal(dim1=-2, dim2=-1).add_((sn2 + 1e-7)[:, None])
    Ks = torch.cdist(as_, a).square_().neg_().exp_().mul_(sig2[:, None, None])      # [B, m, n]
    L = torch.linalg.cholesky(K)
    del K
    # (torch's batched cholesky_solve rejects [6, 8192] on this stack: two triangular solves instead)
    alpha = torch.linalg.solve_triangular(L.transpose(1, 2), torch.linalg.solve_triangular(L, y[..., None], upper=False), upper=True)
    mean = torch.bmm(Ks, alpha)[..., 0]
    V = torch.linalg.solve_triangular(L, Ks.transpose(1, 2), upper=False)               # [B, n, m]
    var = (sig2 + sn2)[:, None] - V.square().sum(1)
    nlml = 0.5 * (y * alpha[..., 0]).sum(-1) + L.diagonal(dim1=-2, dim2=-1).log().sum(-1) + 0.5 * n * math.log(2 * math.pi)
    return mean, var, nlml, alpha[..., 0]


@pytest.mark.parametrize("form", ["subst", "gemm"])   # blocked substitution (bench kernel) / triangular GEMM with L^-1
@pytest.mark.parametrize("B,n,d,m", [(6, 8192, 16, 384), (2, 4096, 16, 256), (1, 16384, 8, 128)])
def test_fit_predict_at_baseline_sizes(monkeypatch, B, n, d, m, form):
    """Expert orders of configs 4 (8192, and the 4096 communication expert) and 2 (16384): both schedules of the
    factorisation are exercised (B=6 -> dataflow, B<6 -> per-column with split-K)."""
    from pygpr_b200 import engine as eng
    from pygpr_b200 import config
    monkeypatch.setattr(config.options, "predict_form", form)
    x, y, xs, hp = _data(11, B, n, d, m)
    st = eng.fit(x, y, hp, True)
    mean, var = eng.predict(st, xs, True)
    torch.cuda.synchronize()
    assert int(st.info.abs().sum()) == 0
    mean_r, var_r, nlml_r, alpha_r = _dense_reference(x, y, xs, hp)
    assert_close(mean.cpu().numpy(), mean_r.cpu().numpy(), TOL, "mean")
    assert_close(var.cpu().numpy(), var_r.cpu().numpy(), TOL, "var")
    assert_close(st.nlml().cpu().numpy(), nlml_r.cpu().numpy(), TOL, "nlml")
    assert_close(st.alpha[:, :n].cpu().numpy(), alpha_r.cpu().numpy(), 1e-7, "alpha")
    # size-independent property: K alpha = y, with K rebuilt by the library's own stand-alone kernel builder
    K = eng.se_kernel_matrix(hp[:, :d + 1].contiguous(), x)
    r = torch.bmm(K, st.alpha[:, :n, None])[..., 0] + (hp[:, d + 1] ** 2 + 1e-7)[:, None] * st.alpha[:, :n] - y
    assert float(r.abs().max()) <= 1e-10 * float(y.abs().max()) * n


def test_lml_gradient_at_config2_size():
    """n=16384, d=8 (config 2): the fused gradient against central differences of the NLML from independent fits, and
    against the trace identity  d NLML / d sig_n = sig_n (tr K^-1 - alpha.alpha)  evaluated with library routines at n=4096."""
    from pygpr_b200 import engine as eng
    B, n, d = 1, 16384, 8
    x, y, _, hp = _data(12, B, n, d, 8)
    st = eng.fit(x, y, hp, True)
    grad = eng.lml_grad(st)[0].cpu()
    eps = 1e-5
    for p in (0, 3, d + 1):
        hp_p, hp_m = hp.clone(), hp.clone()
        hp_p[0, p] += eps
        hp_m[0, p] -= eps
        fp = float(eng.fit(x, y, hp_p, True, want_alpha=False).nlml()[0])
        fm = float(eng.fit(x, y, hp_m, True, want_alpha=False).nlml()[0])
        fd = (fp - fm) / (2 * eps)
        assert abs(fd - float(grad[p])) <= 1e-5 * max(1.0, abs(fd)), (p, fd, float(grad[p]))
    # trace identity with library routines at n=4096 (K^-1 explicitly)
    n2 = 4096
    x2, y2, _, hp2 = _data(13, 2, n2, d, 8)
    st2 = eng.fit(x2, y2, hp2, True)
    g2 = eng.lml_grad(st2).cpu()
    a = x2 * hp2[:, None, 1:d + 1]
    K = torch.cdist(a, a).square_().neg_().exp_().mul_((hp2[:, 0] ** 2)[:, None, None])
    K.diagonal(dim1=-2, dim2=-1).add_((hp2[:, d + 1] ** 2 + 1e-7)[:, None])
    Kinv = torch.cholesky_inverse(torch.linalg.cholesky(K))
    alpha = torch.bmm(Kinv, y2[..., None])[..., 0]
    g_noise = hp2[:, d + 1] * (Kinv.diagonal(dim1=-2, dim2=-1).sum(-1) - (alpha * alpha).sum(-1))
    assert_close(g2[:, d + 1].numpy(), g_noise.cpu().numpy(), TOL, "d NLML / d sig_n")
    W = Kinv - alpha[:, :, None] * alpha[:, None, :]
    Kse = K.clone()
    Kse.diagonal(dim1=-2, dim2=-1).sub_((hp2[:, d + 1] ** 2 + 1e-7)[:, None])
    g_sig = (W * Kse).sum((-1, -2)) / hp2[:, 0]
    assert_close(g2[:, 0].numpy(), g_sig.cpu().numpy(), TOL, "d NLML / d sig")
    k = 2
    dx2 = (x2[:, :, None, k] - x2[:, None, :, k]).square_()
    g_ls = -(W * Kse * dx2).sum((-1, -2)) * hp2[:, 1 + k]
    assert_close(g2[:, 1 + k].numpy(), g_ls.cpu().numpy(), TOL, "d NLML / d ls_k")


def test_grbcm_shard_sum_equals_full_sum_at_scale():
    """Config-3-sized GRBCM (expert order 4096): predicting with the experts split into two shards whose [3, m]
    partial sums are added (the all-reduce of SURVEY §8e) equals the unsharded aggregation."""
    import pygpr_b200 as P
    from pygpr_b200 import engine as eng
    g = torch.Generator().manual_seed(21)
    nc, m0, d, m = 6, 2048, 8, 256
    xall = torch.rand(nc + 1, m0, d, generator=g, dtype=torch.float64)
    yall = torch.sin(xall.sum(-1)) + 0.1 * torch.randn(nc + 1, m0, generator=g, dtype=torch.float64)
    xs = torch.rand(m, d, generator=g, dtype=torch.float64)
    hp = torch.empty(nc, d + 2, dtype=torch.float64)
    hp[:, 0] = 1.0; hp[:, 1:d + 1] = 2.0 / math.sqrt(d); hp[:, d + 1] = 0.3
    cov = P.Compose([P.Squared_exponential(), P.White_noise()])

    def model(lo, hi):
        mdl = P.GRBCM(xall[1 + lo:1 + hi].contiguous(), yall[1 + lo:1 + hi].contiguous(), xall[0], yall[0], cov,
                      expert_offset=lo, nc_total=nc)
        mdl.gpl.set_params(hp[lo:hi]); mdl.gpg.set_params(hp[0])
        return mdl

    full = model(0, nc)
    mu, var = full.predict(xs, var="diag")
    parts = []
    for lo, hi in ((0, 3), (3, 6)):
        mdl = model(lo, hi)
        mdl.gpg.update(); mdl.gpl.update()
        mg, vg = eng.predict(mdl.gpg._state, xs.cuda(), True)
        ml, vl = eng.predict(mdl.gpl._state, xs.cuda(), True)
        parts.append(eng.grbcm_partial(ml, vl, vg[0], lo)[0])
    mu2, var2, _, _ = eng.grbcm_finalize(parts[0] + parts[1], mg[0], vg[0])
    assert_close(mu2.cpu().numpy(), torch.as_tensor(mu).cpu().numpy(), 1e-12, "mu_A sharded")
    assert_close(var2.cpu().numpy(), torch.as_tensor(var).cpu().numpy(), 1e-12, "var_A sharded")
    assert bool((var2 > 0).all())


def test_benchmarked_grbcm_path_against_oracle(monkeypatch):
    """The code path bench.py times, at its per-expert sizes: communication set 4096 (32 shared block columns), local blocks 4096
    (expert order 8192), d = 16, one shared hyper-parameter vector -> gpr_potrf_grbcm (one dataflow launch for both
    factorisations) + gpr_predict_gemm / gpr_predict_shared.  Checked against the CPU ORACLE (gr_bcm.py:19-29,116-155,
    gpr.py:65-106, loss.py:49-51) for the communication expert and two local experts, incl. the aggregate over those two, and
    against the general (unshared) path for all six."""
    import pygpr_b200 as P
    from pygpr_b200 import config
    from oracle import pygpr_oracle as orc
    g = torch.Generator().manual_seed(31)
    nc, m0, d, m = 6, 4096, 16, 384
    xall = torch.rand(nc + 1, m0, d, generator=g, dtype=torch.float64)
    yall = torch.sin(xall.sum(-1)) + 0.1 * torch.randn(nc + 1, m0, generator=g, dtype=torch.float64)
    xs = torch.rand(m, d, generator=g, dtype=torch.float64)
    hp = torch.empty(nc, d + 2, dtype=torch.float64)
    hp[:, 0] = 1.0; hp[:, 1:d + 1] = 2.0 / math.sqrt(d); hp[:, d + 1] = 0.3
    cov = P.Compose([P.Squared_exponential(), P.White_noise()])

    def run(share):
        monkeypatch.setattr(config.options, "share", share)
        mdl = P.GRBCM(xall[1:].contiguous(), yall[1:].contiguous(), xall[0], yall[0], cov)
        mdl.gpl.set_params(hp); mdl.gpg.set_params(hp[0])
        mu, var = mdl.predict(xs, var="diag")
        assert (mdl.gpl._state.shared_from is not None) == share
        return mdl, mu, var

    mdl, mu, var = run(True)
    st, stg = mdl.gpl._state, mdl.gpg._state
    from pygpr_b200 import engine as eng
    (mg, vg), (ml, vl) = eng.predict_pair(stg, st, xs.cuda())
    # oracle: communication expert + local experts 0, 1
    ns = 2
    mu_o, var_o, parts = orc.grbcm_predict(["se", "wn"], hp[0], hp[:ns], xall[1:1 + ns], yall[1:1 + ns], xall[0], yall[0], xs, expert_chunk=1)
    assert_close(mg[0].cpu().numpy(), parts["mean_g"].numpy(), TOL, "communication expert mean")
    assert_close(vg[0].cpu().numpy(), parts["var_g"].numpy(), TOL, "communication expert variance")
    assert_close(ml[:ns].cpu().numpy(), parts["mean_l"].numpy(), TOL, "local expert means")
    assert_close(vl[:ns].cpu().numpy(), parts["var_l"].numpy(), TOL, "local expert variances")
    S, _, _ = eng.grbcm_partial(ml[:ns].contiguous(), vl[:ns].contiguous(), vg[0], 0)
    mu2, var2 = eng.grbcm_finalize(S, mg[0], vg[0])[:2]
    assert_close(mu2.cpu().numpy(), mu_o.numpy(), TOL, "aggregate over the two experts (mean)")
    assert_close(var2.cpu().numpy(), var_o.numpy(), TOL, "aggregate over the two experts (variance)")
    xa, ya = orc.grbcm_inputs(xall[1:1 + ns], yall[1:1 + ns], xall[0], yall[0])
    nl_o = torch.cat([orc.mle_loss(["se", "wn"], hp[i:i + 1], xa[i:i + 1], ya[i:i + 1]) for i in range(ns)])
    assert_close(st.nlml()[:ns].cpu().numpy(), nl_o.numpy(), TOL, "NLML of the local experts")
    assert_close(stg.nlml().cpu().numpy(), orc.mle_loss(["se", "wn"], hp[:1], xall[:1], yall[:1]).numpy(), TOL, "NLML of the communication expert")
    # the opt-in emulated-FP64 prediction at the same sizes (6 digit planes on the INT8 tcgen05 tensor cores) against the oracle
    monkeypatch.setattr(config.options, "emulate_slices", 6)
    monkeypatch.setattr(config.options, "emulate_min_order", 0)
    (mg_e, vg_e), (ml_e, vl_e) = eng.predict_pair(stg, st, xs.cuda())
    monkeypatch.setattr(config.options, "emulate_slices", 0)
    assert_close(mg_e[0].cpu().numpy(), parts["mean_g"].numpy(), TOL, "communication expert mean (emulated)")
    assert_close(vg_e[0].cpu().numpy(), parts["var_g"].numpy(), TOL, "communication expert variance (emulated)")
    assert_close(ml_e[:ns].cpu().numpy(), parts["mean_l"].numpy(), TOL, "local expert means (emulated)")
    assert_close(vl_e[:ns].cpu().numpy(), parts["var_l"].numpy(), TOL, "local expert variances (emulated)")
    assert_close(ml_e.cpu().numpy(), ml.cpu().numpy(), 1e-9, "emulated == FP64 path, all experts (mean)")
    assert_close(vl_e.cpu().numpy(), vl.cpu().numpy(), 1e-9, "emulated == FP64 path, all experts (variance)")
    # all six experts: shared-block path == general path
    mdl0, mu0, var0 = run(False)
    assert_close(torch.as_tensor(mu).numpy(), torch.as_tensor(mu0).numpy(), 1e-10, "shared == general (mean)")
    assert_close(torch.as_tensor(var).numpy(), torch.as_tensor(var0).numpy(), 1e-10, "shared == general (variance)")
    assert_close(st.nlml().cpu().numpy(), mdl0.gpl._state.nlml().cpu().numpy(), 1e-11, "shared == general (NLML)")
