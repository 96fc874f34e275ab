"""GPU parity of the C-ABI kernels (through pygpr_b200.engine) against the CPU oracle on seeded inputs."""
import math

import numpy as np
import pytest
import torch

from helpers import assert_close

pytestmark = pytest.mark.gpu

SPEC = ["se", "wn"]
TOL = 1e-8  # north_star: relative 1e-8 on mean / variance / LML in FP64 (well-conditioned inputs)


def _data(seed, B, n, d, m, sig_n=0.3, per_expert=False):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, n, d, generator=g, dtype=torch.float64)
    y = torch.sin(x.sum(-1)) + 0.1 * torch.randn(B, n, generator=g, dtype=torch.float64)
    xs = torch.rand(m, d, generator=g, dtype=torch.float64)
    hp = torch.empty(B, d + 2, dtype=torch.float64)
    hp[:, 0] = 1.0
    hp[:, 1:d + 1] = 2.0 / math.sqrt(d)
    hp[:, d + 1] = sig_n
    if per_expert:
        hp = hp * (0.5 + torch.rand(B, d + 2, generator=g, dtype=torch.float64))
    return x, y, xs, hp


@pytest.fixture(scope="module")
def eng():
    from pygpr_b200 import engine
    return engine


@pytest.fixture(scope="module")
def orc():
    from oracle import pygpr_oracle
    return pygpr_oracle


@pytest.mark.parametrize("B,n,d,m", [(1, 37, 3, 11), (4, 130, 5, 29), (2, 256, 16, 300), (3, 300, 7, 129)])
def test_se_kernels(eng, orc, B, n, d, m):
    x, y, xs, hp = _data(1, B, n, d, m, per_expert=True)
    hp_se = hp[:, :d + 1].contiguous()
    K = eng.se_kernel_matrix(hp_se.cuda(), x.cuda()).cpu()
    assert_close(K.numpy(), orc.se_kernel(hp_se, x).numpy(), 1e-13, "K_se")
    Ks = eng.se_cross_matrix(hp_se.cuda(), x.cuda(), xs.cuda()).cpu()
    assert_close(Ks.numpy(), orc.se_kernel(hp_se, x, xs).numpy(), 1e-13, "K*_se")
    dK = eng.se_kernel_grad(hp_se.cuda(), x.cuda(), K.cuda()).cpu()
    assert_close(dK.numpy(), orc.se_kernel_and_grad(hp_se, x)[1].numpy(), 1e-13, "dK_se")


@pytest.mark.parametrize("B,n,d,m,per", [(1, 37, 3, 11, False), (1, 128, 2, 128, False), (4, 130, 5, 29, True),
                                         (2, 256, 16, 300, False), (3, 300, 7, 129, True), (2, 1000, 1, 200, False),
                                         (1, 1500, 8, 64, False)])
@pytest.mark.parametrize("form", ["subst", "gemm"])
def test_fit_predict(eng, orc, monkeypatch, B, n, d, m, per, form):
    """form: blocked substitution (the bench kernel) / triangular GEMM with L^-1 (what few-task predictions use)."""
    from pygpr_b200 import config
    monkeypatch.setattr(config.options, "predict_form", form)
    x, y, xs, hp = _data(2, B, n, d, m, per_expert=per)
    st = eng.fit(x.cuda(), y.cuda(), hp.cuda(), True)
    torch.cuda.synchronize()
    assert int(st.info.abs().sum()) == 0
    L_ref, alpha_ref = orc.gp_fit(SPEC, hp, x, y)
    L = torch.tril(st.L[:, :n, :n]).cpu()
    assert_close(L.numpy(), L_ref.numpy(), 1e-10, "L")
    assert_close(st.alpha[:, :n].cpu().numpy(), alpha_ref.numpy(), TOL, "alpha")
    assert float(st.alpha[:, n:].abs().max()) == 0.0 if st.npad > n else True
    assert_close(st.nlml().cpu().numpy(), orc.mle_loss(SPEC, hp, x, y).numpy(), TOL, "nlml")
    mean_ref, var_ref = orc.gp_predict(SPEC, hp, x, y, xs, "diag", fit=(L_ref, alpha_ref))
    mean, var = eng.predict(st, xs.cuda(), True)
    assert_close(mean.cpu().numpy(), mean_ref.numpy(), TOL, "mean")
    assert_close(var.cpu().numpy(), var_ref.numpy(), TOL, "var")
    mean2, none = eng.predict(st, xs.cuda(), False)
    assert none is None
    assert_close(mean2.cpu().numpy(), mean_ref.numpy(), TOL, "mean (no var)")


def test_potrf_inplace_and_info(eng, orc):
    x, y, xs, hp = _data(3, 2, 200, 4, 8)
    K = orc.compose_kernel(SPEC, hp, x)
    K.diagonal(dim1=-2, dim2=-1).add_(1e-7)
    npad = 256
    Kp = torch.eye(npad, dtype=torch.float64).repeat(2, 1, 1)
    Kp[:, :200, :200] = K
    Kd = Kp.cuda()
    invdiag, z, ldh, zz, info = eng.potrf_inplace(Kd)
    assert int(info.abs().sum()) == 0
    assert_close(torch.tril(Kd[:, :200, :200]).cpu().numpy(), torch.linalg.cholesky(K).numpy(), 1e-10, "L in place")
    assert_close(ldh.cpu().numpy(), torch.linalg.cholesky(K).diagonal(dim1=-2, dim2=-1).log().sum(-1).numpy(), 1e-10, "logdet")
    # a non positive-definite matrix is reported LAPACK-style in info (first bad leading minor), not as a crash
    Kbad = Kp.clone()
    Kbad[1, 150, 150] = -1.0
    Kb = Kbad.cuda()
    _, _, _, _, info = eng.potrf_inplace(Kb)
    info = info.cpu()
    assert int(info[0]) == 0 and int(info[1]) == 151


def test_grbcm_aggregate(eng, orc):
    g = torch.Generator().manual_seed(5)
    nc, m = 9, 333
    mean_l = torch.randn(nc, m, generator=g, dtype=torch.float64)
    var_l = 0.05 + torch.rand(nc, m, generator=g, dtype=torch.float64)
    mean_g = torch.randn(m, generator=g, dtype=torch.float64)
    var_g = 0.5 + torch.rand(m, generator=g, dtype=torch.float64)
    mu_ref, var_ref, beta_ref, prec_ref = orc.grbcm_aggregate(mean_g, var_g, mean_l, var_l)
    # single shard
    S, br, pr = eng.grbcm_partial(mean_l.cuda(), var_l.cuda(), var_g.cuda(), 0, want_rows=True)
    mu, var, b0, p0 = eng.grbcm_finalize(S, mean_g.cuda(), var_g.cuda(), want_rows=True)
    assert_close(mu.cpu().numpy(), mu_ref.numpy(), 1e-12, "mu_A")
    assert_close(var.cpu().numpy(), var_ref.numpy(), 1e-12, "var_A")
    assert_close(torch.cat((b0[None], br)).cpu().numpy(), beta_ref.numpy(), 1e-12, "beta")
    assert_close(torch.cat((p0[None], pr)).cpu().numpy(), prec_ref.numpy(), 1e-12, "prec")
    # two shards summed = the all-reduce of SURVEY §8(e)
    S1, _, _ = eng.grbcm_partial(mean_l[:4].cuda(), var_l[:4].cuda(), var_g.cuda(), 0)
    S2, _, _ = eng.grbcm_partial(mean_l[4:].cuda(), var_l[4:].cuda(), var_g.cuda(), 4)
    mu2, var2, _, _ = eng.grbcm_finalize(S1 + S2, mean_g.cuda(), var_g.cuda())
    assert_close(mu2.cpu().numpy(), mu_ref.numpy(), 1e-12, "mu_A sharded")
    assert_close(var2.cpu().numpy(), var_ref.numpy(), 1e-12, "var_A sharded")


@pytest.mark.parametrize("mode", ["columns", "dataflow"])
@pytest.mark.parametrize("B,n,d", [(3, 300, 4), (7, 700, 8), (1, 129, 2)])
def test_fit_schedules(eng, orc, monkeypatch, mode, B, n, d):
    """Both schedules of the batched Cholesky (one launch pair per block column with split-K / one dataflow launch with
    per-row progress counters) give the reference's factor, forward solve and NLML."""
    from pygpr_b200 import config
    config.set_potrf_mode(mode)
    x, y, xs, hp = _data(7, B, n, d, 16, per_expert=True)
    st = eng.fit(x.cuda(), y.cuda(), hp.cuda(), True)
    torch.cuda.synchronize()
    assert int(st.info.abs().sum()) == 0
    L_ref, alpha_ref = orc.gp_fit(SPEC, hp, x, y)
    assert_close(torch.tril(st.L[:, :n, :n]).cpu().numpy(), L_ref.numpy(), 1e-10, "L")
    assert_close(st.alpha[:, :n].cpu().numpy(), alpha_ref.numpy(), TOL, "alpha")
    assert_close(st.nlml().cpu().numpy(), orc.mle_loss(SPEC, hp, x, y).numpy(), TOL, "nlml")


def test_dataflow_info_and_repeat(eng, orc, monkeypatch):
    """Dataflow schedule: a non-PD expert is reported in info without disturbing the others or hanging the waits, and
    the progress counters are reset between calls (two calls in a row agree bit for bit)."""
    from pygpr_b200 import config
    config.set_potrf_mode("dataflow")
    x, y, xs, hp = _data(8, 6, 300, 4, 8)
    K = orc.compose_kernel(SPEC, hp, x)
    K.diagonal(dim1=-2, dim2=-1).add_(1e-7)
    npad = 384
    Kp = torch.eye(npad, dtype=torch.float64).repeat(6, 1, 1)
    Kp[:, :300, :300] = K
    Kbad = Kp.clone()
    Kbad[4, 140, 140] = -1.0
    Kb = Kbad.cuda()
    _, _, _, _, info = eng.potrf_inplace(Kb)
    info = info.cpu()
    assert [int(v) for v in info] == [0, 0, 0, 0, 141, 0]
    good = [0, 1, 2, 3, 5]
    assert_close(torch.tril(Kb[good, :300, :300]).cpu().numpy(), torch.linalg.cholesky(K[good]).numpy(), 1e-10, "L of the PD experts")
    a, b = Kp.cuda(), Kp.cuda()
    eng.potrf_inplace(a)
    eng.potrf_inplace(b)
    assert torch.equal(torch.tril(a), torch.tril(b))


@pytest.mark.parametrize("B,n,d", [(1, 1500, 4), (3, 1100, 8), (5, 520, 3), (40, 640, 4)])
def test_backward_solve_forms(eng, orc, monkeypatch, B, n, d):
    """alpha = L^-T z: the cooperative column-owner form (several CTAs per expert, used for few experts) and the
    one-CTA-per-expert form agree with each other and with the reference's cholesky_solve (gpr.py:70-72)."""
    x, y, xs, hp = _data(9, B, n, d, 8, per_expert=True)
    from pygpr_b200 import config
    config.set_trsv_mode("single")
    a1 = eng.fit(x.cuda(), y.cuda(), hp.cuda(), True).alpha.clone()
    config.set_trsv_mode("auto")
    a2 = eng.fit(x.cuda(), y.cuda(), hp.cuda(), True).alpha.clone()
    torch.cuda.synchronize()
    _, alpha_ref = orc.gp_fit(SPEC, hp, x, y)
    assert_close(a1[:, :n].cpu().numpy(), alpha_ref.numpy(), TOL, "alpha (one CTA per expert)")
    assert_close(a2[:, :n].cpu().numpy(), alpha_ref.numpy(), TOL, "alpha (column owners)")
    assert_close(a2.cpu().numpy(), a1.cpu().numpy(), 1e-12, "forms agree")


@pytest.mark.parametrize("nc,nls,ng,d", [(7, 200, 256, 5), (40, 768, 384, 3)])   # column-mode / task-mode L^-T
def test_shared_block_fit_and_gradient(eng, orc, nc, nls, ng, d):
    """Experts that start with the same ng rows under one hyper-parameter vector: factor, NLML and gradient built on top of the
    communication expert's factor / L_g^-T / K_g^-1 equal the stand-alone ones and the reference's (loss.py:92-128)."""
    g = torch.Generator().manual_seed(31)
    xg = torch.rand(ng, d, generator=g, dtype=torch.float64)
    yg = torch.sin(xg.sum(-1)) + 0.1 * torch.randn(ng, generator=g, dtype=torch.float64)
    xl = torch.rand(nc, nls, d, generator=g, dtype=torch.float64)
    yl = torch.sin(xl.sum(-1)) + 0.1 * torch.randn(nc, nls, generator=g, dtype=torch.float64)
    x, y = orc.grbcm_inputs(xl, yl, xg, yg)
    hp = torch.empty(nc, d + 2, dtype=torch.float64)
    hp[:, 0] = 1.1; hp[:, 1:d + 1] = 2.0 / math.sqrt(d); hp[:, d + 1] = 0.3
    st_g = eng.fit(xg[None].cuda(), yg[None].cuda(), hp[:1].cuda(), True)
    eng.kinv(st_g)
    st = eng.fit(x.cuda(), y.cuda(), hp.cuda(), True, shared=st_g)
    assert st.shared_from is st_g and st.shared_upper and int(st.info.abs().sum()) == 0
    st0 = eng.fit(x.cuda(), y.cuda(), hp.cuda(), True)
    n = x.shape[1]
    assert_close(torch.tril(st.L[:, :n, :n]).cpu().numpy(), torch.tril(st0.L[:, :n, :n]).cpu().numpy(), 1e-11, "L")
    assert_close(st.nlml().cpu().numpy(), st0.nlml().cpu().numpy(), 1e-12, "nlml")
    g1, g0 = eng.lml_grad(st).cpu(), eng.lml_grad(st0).cpu()
    assert_close(g1.numpy(), g0.numpy(), 1e-10, "gradient (shared block) == gradient (stand-alone)")
    if n <= 600:
        loss_o, grad_o = orc.mle_loss_and_grad(SPEC, hp, x, y)
        assert_close(st.nlml().cpu().numpy(), loss_o.numpy(), TOL, "nlml vs reference")
        assert_close(g1.numpy(), grad_o.numpy(), TOL, "gradient vs reference")


def test_vt_forms_agree(eng, orc):
    """V^T = K* L^-T kept by the substitution (gpr_predict_keep) and produced by the triangular GEMM with L^-1
    (gpr_predict_gemm) are the same matrix, and both give the reference's mean / variance; the GEMM form's partial sums over
    the first whole blocks are the sums over those training rows."""
    from pygpr_b200 import _lib
    x, y, xs, hp = _data(17, 1, 500, 6, 150)
    st = eng.fit(x.cuda(), y.cuda(), hp.cuda(), True)
    m, mt = xs.shape[0], 2
    xd = xs.cuda()
    lib = _lib.load()
    outs = []
    for form in ("keep", "gemm"):
        mean = torch.empty(1, m, dtype=torch.float64, device="cuda"); var = torch.empty_like(mean)
        vt = torch.zeros(1, mt, 128, st.npad, dtype=torch.float64, device="cuda")
        if form == "keep":
            scratch = torch.empty(int(lib.gpr_predict_scratch_bytes(st.npad)), dtype=torch.uint8, device="cuda")
            _lib.call("gpr_predict_keep", st.L.data_ptr(), st.invdiag.data_ptr(), st.a.data_ptr(), st.nrm.data_ptr(), st.alpha.data_ptr(),
                      st.hp.data_ptr(), st.hp.shape[-1], st.prior.data_ptr(), xd.data_ptr(), 1, st.n, st.npad, st.d, st.dpad, m,
                      mean.data_ptr(), var.data_ptr(), vt.data_ptr(), scratch.data_ptr(), torch.cuda.current_stream().cuda_stream)
        else:
            X = eng.tril_inverse(st)
            mp = torch.empty_like(mean); vp = torch.empty_like(mean)
            scratch = torch.empty(int(lib.gpr_predict_gemm_scratch_bytes(1, st.npad, st.dpad, m)), dtype=torch.uint8, device="cuda")
            _lib.call("gpr_predict_gemm", X.data_ptr(), st.a.data_ptr(), st.nrm.data_ptr(), st.z.data_ptr(), st.hp.data_ptr(),
                      st.hp.shape[-1], st.prior.data_ptr(), xd.data_ptr(), 1, st.n, st.npad, st.d, st.dpad, m, mean.data_ptr(),
                      var.data_ptr(), vt.data_ptr(), scratch.data_ptr(), torch.cuda.current_stream().cuda_stream, 256,
                      mp.data_ptr(), vp.data_ptr())
        torch.cuda.synchronize()
        outs.append((mean.cpu(), var.cpu(), vt.reshape(mt * 128, st.npad)[:m, :st.n].cpu()))
    mean_ref, var_ref = orc.gp_predict(SPEC, hp, x, y, xs, "diag")
    for mean, var, _ in outs:
        assert_close(mean.numpy(), mean_ref.numpy(), TOL, "mean")
        assert_close(var.numpy(), var_ref.numpy(), TOL, "var")
    assert_close(outs[1][2].numpy(), outs[0][2].numpy(), 1e-10, "V^T (GEMM form) == V^T (substitution)")
    V = outs[0][2]
    z = st.z[0, :st.n].cpu()
    assert_close(mp[0].cpu().numpy(), (V[:, :256] @ z[:256]).numpy(), 1e-10, "mean_part")
    assert_close(vp[0].cpu().numpy(), (float(st.prior[0]) - (V[:, :256] ** 2).sum(-1)).numpy(), 1e-10, "var_part")


# ---- emulated FP64 on the INT8 tcgen05 tensor cores (csrc/ozaki_gemm.cu): integer work is checked BIT-EXACT ----------------------
@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (200, 100, 130), (300, 257, 1000), (1, 1, 1)])
@pytest.mark.parametrize("slices,bk", [(1, 0), (3, 32), (6, 0), (6, 32), (7, 64), (8, 0)])
def test_ozaki_digit_planes_and_product_bit_exact(eng, M, N, K, slices, bk):
    """gpr_ozaki_split's digit planes / row scales equal the numpy restatement (oracle/ozaki_ref.py) bit for bit, and
    gpr_ozaki_gemm_nt's C equals the restatement's (exact integer diagonals, same FP64 recombination order) bit for bit."""
    from oracle import ozaki_ref
    from pygpr_b200 import config
    config.set_ozaki_bk(bk)                                  # both k-block variants (64 B / 32 B swizzled rows), and the automatic choice
    g = torch.Generator().manual_seed(3)
    A = torch.randn(M, K, generator=g, dtype=torch.float64) * torch.exp(2.0 * torch.randn(M, 1, generator=g, dtype=torch.float64))
    B = torch.randn(N, K, generator=g, dtype=torch.float64)
    A[0, :] = 0.0 if M > 1 else A[0, :]                      # an all-zero row: scale 0, digits 0
    pa, pb = eng.ozaki_split(A.cuda(), slices), eng.ozaki_split(B.cuda(), slices)
    npa, nsa = ozaki_ref.split(A.numpy(), slices)
    npb, nsb = ozaki_ref.split(B.numpy(), slices)
    assert np.array_equal(pa.planes.cpu().numpy(), npa) and np.array_equal(pb.planes.cpu().numpy(), npb)
    assert np.array_equal(pa.scale.cpu().numpy(), nsa) and np.array_equal(pb.scale.cpu().numpy(), nsb)
    C = eng.ozaki_gemm_nt(pa, pb).cpu().numpy()
    assert np.array_equal(C, ozaki_ref.gemm_nt(npa, nsa, M, npb, nsb, N, slices))
    if slices >= 6:
        ref = (A @ B.T).numpy()
        assert np.abs(C - ref).max() <= (1e-11 if slices == 6 else 1e-13) * np.abs(ref).max()


def test_ozaki_gemm_accumulate_lower_and_batched(eng):
    """beta = 1 accumulation, the lower-trapezoidal B option (zero k-blocks skipped) and the batched launch against FP64 products."""
    g = torch.Generator().manual_seed(4)
    M, N, K = 256, 192, 320
    A = torch.randn(M, K, generator=g, dtype=torch.float64).cuda()
    off = 128
    Bm = torch.randn(N, K, generator=g, dtype=torch.float64)
    Bm = (Bm * (torch.arange(K)[None, :] <= (torch.arange(N)[:, None] + off))).cuda()   # B[j, k] = 0 for k > j + off
    C0 = torch.randn(M, N, generator=g, dtype=torch.float64).cuda()
    pa, pb = eng.ozaki_split(A, 7), eng.ozaki_split(Bm, 7)
    C = eng.ozaki_gemm_nt(pa, pb, out=C0.clone(), beta=1.0, b_lower=off)
    ref = C0 + A @ Bm.T
    assert float((C - ref).abs().max() / ref.abs().max()) <= 1e-13
    assert torch.equal(eng.ozaki_gemm_nt(pa, pb, b_lower=off), eng.ozaki_gemm_nt(pa, pb))   # skipping zeros changes nothing
    # rows of an upper-triangular factor times themselves, lower tiles of C only (K^-1 = U U^T of the emulated gradient)
    n = 384
    U = torch.triu(torch.randn(n, n, generator=g, dtype=torch.float64)).cuda()
    pu = eng.ozaki_split(U, 7)
    Cl = torch.full((n, n), float("nan"), dtype=torch.float64, device="cuda")
    eng.ozaki_gemm_nt(pu, pu, out=Cl, a_upper=0, c_lower=True)
    refu = U @ U.T
    rows, cols = torch.arange(n, device="cuda")[:, None], torch.arange(n, device="cuda")[None, :]
    computed = (cols // 64) * 64 <= (rows // 128) * 128 + 127            # the 128 x 64 tiles that touch the lower triangle
    assert bool(torch.isnan(Cl[~computed]).all()) and not bool(torch.isnan(Cl[computed]).any())
    assert float((Cl[computed] - refu[computed]).abs().max() / refu.abs().max()) <= 1e-13
    assert torch.equal(Cl[computed], eng.ozaki_gemm_nt(pu, pu)[computed])   # skipping zero k-blocks changes nothing
    nb = 3
    Ab = torch.randn(nb, 128, 192, generator=g, dtype=torch.float64).cuda()
    Bb = torch.randn(nb, 256, 192, generator=g, dtype=torch.float64).cuda()
    pab, pbb = eng.ozaki_split(Ab.reshape(nb * 128, 192), 7), eng.ozaki_split(Bb.reshape(nb * 256, 192), 7)
    Cb = eng.ozaki_gemm_nt(pab, pbb, batch=nb)
    refb = torch.bmm(Ab, Bb.transpose(1, 2))
    assert float((Cb - refb).abs().max() / refb.abs().max()) <= 1e-13
    for i in range(nb):   # a batch entry equals the same product launched alone, bit for bit
        assert torch.equal(Cb[i], eng.ozaki_gemm_nt(eng.ozaki_split(Ab[i], 7), eng.ozaki_split(Bb[i], 7)))


@pytest.mark.parametrize("shared", [False, True])
def test_inverse_factors_assembled_from_segments(eng, orc, monkeypatch, shared):
    """L^-T assembled from FP64 inverses of diagonal segments and emulated-FP64 products between them (engine._inverse_factors,
    opt-in) equals the plain FP64 blocked trtri and the oracle's inverse; with and without a shared communication block."""
    from pygpr_b200 import config
    B, n, d = 20, 2048, 4                       # 20 x 15 row tasks: the persistent task schedule of trtri_kernel applies
    x, y, xs, hp = _data(31, B, n, d, 128)
    if shared:
        x[:, :1024] = x[0, :1024]               # every expert starts with the same 1024 points (the communication set)
        y[:, :1024] = y[0, :1024]

    def inverse(slices):
        monkeypatch.setattr(config.options, "emulate_inverse_slices", slices)
        st_g = None
        if shared:
            st_g = eng.fit(x[:1, :1024].contiguous().cuda(), y[:1, :1024].contiguous().cuda(), hp[:1].cuda(), True)
        st = eng.fit(x.cuda(), y.cuda(), hp.cuda(), True, shared=st_g)
        eng.ensure_inverse_factors(st_g, st)
        U = torch.empty(B, n, n, dtype=torch.float64, device="cuda")
        from pygpr_b200 import _lib
        _lib.call("gpr_factor_window", eng._p(st.L), eng._p(st.invdiag), B, st.npad, 0, 0, n, 0, n, eng._p(U), eng._stream())
        return st, U

    st0, U0 = inverse(0)
    st1, U1 = inverse(7)
    assert eng._inverse_plan(st1.npad // 128, 8 if shared else 0, B, 148) > 0
    torch.cuda.synchronize()
    assert float((U1 - U0).abs().max() / U0.abs().max()) <= 1e-11
    L_ref, _ = orc.gp_fit(SPEC, hp[:2], x[:2], y[:2])
    Uref = torch.linalg.inv(L_ref).transpose(1, 2)
    assert_close(U1[:2].cpu().numpy(), Uref.numpy(), 1e-9, "L^-T (assembled) vs the oracle's factor inverted")
    # and the emulated prediction on top of it
    monkeypatch.setattr(config.options, "emulate_inverse_slices", 0)
    mean, var = eng.predict_emulated(st1, xs.cuda(), 7, 4) if not shared else (None, None)
    if not shared:
        mean_o, var_o = orc.gp_predict(SPEC, hp[:2], x[:2], y[:2], xs, "diag")
        assert_close(mean[:2].cpu().numpy(), mean_o.numpy(), TOL, "mean")
        assert_close(var[:2].cpu().numpy(), var_o.numpy(), TOL, "var")


@pytest.mark.parametrize("shared,n", [(False, 1024), (True, 1024), (False, 1100), (True, 900)])
def test_fit_emulated_matches_fp64_fit(eng, orc, monkeypatch, shared, n):
    """Opt-in factorisation with emulated-FP64 panel solves and trailing updates (engine.fit_emulated, 7 digit planes, several panels
    and expert chunks) against the FP64 DMMA Cholesky and the oracle: factor, forward / backward solve, NLML; non-PD reporting."""
    from pygpr_b200 import config
    B, d = 5, 6
    x, y, xs, hp = _data(37, B, n, d, 64)
    hp = hp[:1].expand(B, -1).contiguous()
    st_g0 = st_g1 = None
    if shared:
        x[:, :512] = x[0, :512]; y[:, :512] = y[0, :512]
        st_g0 = eng.fit(x[:1, :512].contiguous().cuda(), y[:1, :512].contiguous().cuda(), hp[:1].cuda(), True)
        st_g1 = eng.fit(x[:1, :512].contiguous().cuda(), y[:1, :512].contiguous().cuda(), hp[:1].cuda(), True)
    st0 = eng.fit(x.cuda(), y.cuda(), hp.cuda(), True, shared=st_g0)
    st1 = eng.fit_emulated(x.cuda(), y.cuda(), hp.cuda(), True, slices=7, shared=st_g1, panel=256, chunk=2)
    torch.cuda.synchronize()
    assert int(st1.info.abs().sum()) == 0
    assert_close(torch.tril(st1.L[:, :n, :n]).cpu().numpy(), torch.tril(st0.L[:, :n, :n]).cpu().numpy(), 1e-10, "L: emulated == FP64")
    assert_close(st1.z[:, :n].cpu().numpy(), st0.z[:, :n].cpu().numpy(), 1e-10, "z")
    assert float(st1.z[:, n:].abs().sum()) == 0.0 and st1.npad == st0.npad
    assert_close(st1.alpha[:, :n].cpu().numpy(), st0.alpha[:, :n].cpu().numpy(), 1e-9, "alpha")
    assert_close(st1.nlml().cpu().numpy(), st0.nlml().cpu().numpy(), 1e-11, "nlml")
    L_ref, alpha_ref = orc.gp_fit(SPEC, hp[:2], x[:2], y[:2])
    assert_close(torch.tril(st1.L[:2, :n, :n]).cpu().numpy(), L_ref.numpy(), 1e-9, "L vs oracle")
    assert_close(st1.alpha[:2, :n].cpu().numpy(), alpha_ref.numpy(), TOL, "alpha vs oracle")
    assert_close(st1.nlml()[:2].cpu().numpy(), orc.mle_loss(SPEC, hp[:2], x[:2], y[:2]).numpy(), TOL, "nlml vs oracle")
    mean, var = eng.predict(st1, xs.cuda(), True) if not shared else eng.predict_pair(st_g1, st1, xs.cuda())[1]
    mean0, var0 = eng.predict(st0, xs.cuda(), True) if not shared else eng.predict_pair(st_g0, st0, xs.cuda())[1]
    assert_close(mean.cpu().numpy(), mean0.cpu().numpy(), 1e-9, "prediction from the emulated factor (mean)")
    assert_close(var.cpu().numpy(), var0.cpu().numpy(), 1e-9, "prediction from the emulated factor (variance)")
    if not shared:   # a non-positive-definite expert: reported in info like the FP64 path, the others unaffected
        xb = x.clone(); xb[3, 700] = xb[3, 100]                  # duplicate point + negative "noise": K loses definiteness
        hpb = hp.clone(); hpb[3, d + 1] = 0.0
        stb = eng.fit_emulated(xb.cuda(), y.cuda(), hpb.cuda(), True, slices=7, panel=256, chunk=2, jitter=-1e-3)
        st_ref = eng.fit(xb.cuda(), y.cuda(), hpb.cuda(), True, jitter=-1e-3)
        torch.cuda.synchronize()
        assert int(stb.info[3]) > 0 and [int(v) for v in stb.info.cpu()] == [int(v) for v in st_ref.info.cpu()]


@pytest.mark.parametrize("B,n,d,m,slices", [(2, 300, 5, 256, 6), (1, 512, 16, 128, 7), (3, 200, 30, 128, 4)])
def test_fused_cross_kernel_digit_planes_bit_exact(eng, B, n, d, m, slices):
    """gpr_se_cross_planes (K* generated straight into its digit planes) gives bit for bit the planes and scales of the two-step
    form gpr_se_cross_ld -> gpr_ozaki_split(bound = sig^2), zeros in the padded k-range included (covar.py:151-160 on the emulated
    prediction path)."""
    from pygpr_b200 import _lib
    x, y, xs, hp = _data(53, B, n, d, m, per_expert=True)
    npad = eng.round_up(n, 128)
    hp_d = hp.cuda()
    a, nrm = eng.scale_inputs(x.cuda(), hp_d, B, npad)
    as_, nrms = eng.scale_inputs(xs.cuda(), hp_d, B, m)
    dpad = a.shape[2]
    assert _lib.load().gpr_se_cross_planes_ok(dpad, slices) == 1
    bound = float((hp[:, 0] * hp[:, 0]).max())
    Ks = torch.zeros(B, m, npad, dtype=torch.float64, device="cuda")
    _lib.call("gpr_se_cross_ld", eng._p(as_), eng._p(nrms), m, eng._p(a), eng._p(nrm), npad, eng._p(hp_d), hp_d.shape[-1], B, m, n, dpad,
              eng._p(Ks), npad, m * npad, eng._stream())
    two = eng.ozaki_split(Ks.view(B * m, npad), slices, bound=bound)
    planes = torch.full((slices, B * m, npad), 77, dtype=torch.int8, device="cuda")
    scale = torch.full((B * m,), -1.0, dtype=torch.float64, device="cuda")
    _lib.call("gpr_se_cross_planes", eng._p(as_), eng._p(nrms), m, eng._p(a), eng._p(nrm), npad, eng._p(hp_d), hp_d.shape[-1], B, n, dpad,
              slices, bound, eng._p(planes), eng._p(scale), eng._stream())
    torch.cuda.synchronize()
    assert torch.equal(planes, two.planes) and torch.equal(scale, two.scale)
    assert int(planes[:, :, n:].abs().sum()) == 0 and int(planes.abs().sum()) > 0
    # ... and both equal the numpy restatement of the split with a known bound (oracle/ozaki_ref.py)
    import numpy as np
    from oracle import ozaki_ref
    npl, nsc = ozaki_ref.split(Ks.view(B * m, npad).cpu().numpy(), slices, bound=bound)
    assert np.array_equal(planes.cpu().numpy(), npl) and np.array_equal(scale.cpu().numpy(), nsc)
