"""Host-side logic that needs no GPU: optimisers (PyGPR/tests/test_opt.py), learn-rate rule, sharding, gloo all-reduce."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class QuadLoss:
    """Random SPD quadratic standing in for a Loss (the reference's test uses Loss(None) the same way)."""

    def __init__(self, dim, seed):
        rng = np.random.RandomState(seed)
        A = rng.rand(dim, dim)
        self.H = A @ A.T + dim * np.eye(dim)
        self.b = rng.rand(dim)
        self.model = None
        self.loss_value = None
        self.grad_value = None

    def loss(self, x):
        return 0.5 * x @ self.H @ x - self.b @ x

    def grad(self, x):
        return self.H @ x - self.b

    def loss_and_grad(self, x):
        return self.loss(x), self.grad(x)


@pytest.mark.parametrize("dim", [2, 3, 5, 7])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_opt_quad(dim, seed, tmp_path, monkeypatch):
    """CG_Quad and BFGS_Quad reach the minimiser of a random SPD quadratic (PyGPR/tests/test_opt.py:20-56)."""
    from pygpr_b200 import BFGS_Quad, CG_Quad, hessian
    monkeypatch.chdir(tmp_path)
    loss = QuadLoss(dim, seed)
    xmin = np.linalg.solve(loss.H, loss.b)
    x0 = np.random.RandomState(seed + 10).rand(dim)
    cg = CG_Quad(loss, gtol=1e-8, max_iter=200, fd_eps=1e-6)
    cg.minimize(x0.copy())
    assert np.allclose(cg.x, xmin, rtol=1e-3)
    bf = BFGS_Quad(loss, gtol=1e-8, max_iter=200)
    bf.minimize(x0.copy(), H0=hessian(x0, loss.grad, 1e-6))
    assert np.allclose(bf.x, xmin, rtol=1e-3)
    assert os.path.exists("opt.dat")


def test_learn_rate_parabola():
    """On an exact quadratic the parabola through f(theta), f(theta -/+ eps J) gives the exact line minimum
    (hp_update.py:14-22): gamma = J.J / J.H.J."""
    from pygpr_b200 import get_learn_rate
    loss = QuadLoss(4, 3)
    theta = torch.tensor([0.3, -0.2, 0.5, 0.1])
    gamma = get_learn_rate(theta, loss, 1e-3)
    J = loss.grad(theta.numpy())
    assert np.isclose(gamma, (J @ J) / (J @ loss.H @ J), rtol=1e-6)


def test_shard_bounds_cover_all_experts():
    from pygpr_b200.dist import shard_bounds
    for nc in (1, 7, 63, 255, 511):
        for ws in (1, 2, 4, 8):
            parts = [shard_bounds(nc, ws, r) for r in range(ws)]
            assert parts[0][0] == 0 and parts[-1][1] == nc
            assert all(parts[i][1] == parts[i + 1][0] for i in range(ws - 1))
            assert max(hi - lo for lo, hi in parts) == -(-nc // ws)
    assert [hi - lo for lo, hi in (shard_bounds(255, 8, r) for r in range(8))] == [32] * 7 + [31]
    assert [hi - lo for lo, hi in (shard_bounds(9, 8, r) for r in range(8))] == [2] + [1] * 7      # balanced: no empty rank while nc >= G
    assert [hi - lo for lo, hi in (shard_bounds(3, 4, r) for r in range(4))] == [1, 1, 1, 0]


def test_unsharded_model_enters_no_collective():
    """Without a process group, and for a model that holds every expert, GRBCM.sharded is False whatever expert_offset / nc_total say
    (the bench constructs its 1-GPU model with them)."""
    import pygpr_b200 as P
    cov = P.Compose([P.Squared_exponential(), P.White_noise()])
    xl, yl, xg, yg = torch.rand(4, 20, 2), torch.rand(4, 20), torch.rand(10, 2), torch.rand(10)
    assert not P.GRBCM(xl, yl, xg, yg, cov).sharded
    assert not P.GRBCM(xl, yl, xg, yg, cov, expert_offset=0, nc_total=4).sharded
    assert not P.GRBCM(xl, yl, xg, yg, cov, shard=True).sharded


def _partial_cpu(mean_l, var_l, var_g, first):
    """numpy restatement of grbcm_partial (include/pygpr_b200.h) used only to exercise the sharded protocol on CPU."""
    p = 1.0 / var_l
    beta = 0.5 * (torch.log(p) - torch.log(1.0 / var_g))
    if first == 0:
        beta[0] = 1.0
    return torch.stack((beta.sum(0), (beta * p).sum(0), (beta * p * mean_l).sum(0)))


def _worker(rank, ws, port, tmp):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    from oracle import pygpr_oracle as orc
    from pygpr_b200 import dist as gdist
    g = torch.Generator().manual_seed(7)
    nc, m = 9, 57
    mean_l = torch.randn(nc, m, generator=g, dtype=torch.float64)
    var_l = 0.05 + torch.rand(nc, m, generator=g, dtype=torch.float64)
    mean_g = torch.randn(m, generator=g, dtype=torch.float64)
    var_g = 0.5 + torch.rand(m, generator=g, dtype=torch.float64)
    lo, hi = gdist.shard_bounds(nc, ws, rank)
    S = _partial_cpu(mean_l[lo:hi], var_l[lo:hi], var_g, lo)
    gdist.all_reduce_sum_(S)                                   # the [3, m] all-reduce of SURVEY §8(e)
    b0 = 1.0 - S[0]
    var_a = 1.0 / (S[1] + b0 / var_g)
    mu_a = var_a * (S[2] + b0 * mean_g / var_g)
    mu_ref, var_ref, beta_ref, _ = orc.grbcm_aggregate(mean_g, var_g, mean_l, var_l)
    ok = torch.allclose(mu_a, mu_ref, rtol=1e-12, atol=1e-14) and torch.allclose(var_a, var_ref, rtol=1e-12, atol=1e-14)
    rows = gdist.all_gather_rows(mean_l[lo:hi].contiguous(), [gdist.shard_bounds(nc, ws, r)[1] - gdist.shard_bounds(nc, ws, r)[0] for r in range(ws)])
    ok = ok and torch.equal(rows, mean_l) and gdist.world() == (rank, ws)
    # which models enter the collectives: only really sharded ones (a replicated GRBCM under an initialised process group must not
    # sum its identical partial sums world-size times, nor wait for ranks that never call predict)
    import pygpr_b200 as P
    cov = P.Compose([P.Squared_exponential(), P.White_noise()])
    xl, yl, xg, yg = torch.rand(4, 20, 2), torch.rand(4, 20), torch.rand(10, 2), torch.rand(10)
    ok = ok and not P.GRBCM(xl, yl, xg, yg, cov).sharded
    sh = P.GRBCM(xl, yl, xg, yg, cov, shard=True)
    ok = ok and sh.sharded and sh.nc == 2 and sh.expert_offset == 2 * rank and sh.nc_total == 4
    ok = ok and P.GRBCM(xl[:2], yl[:2], xg, yg, cov, expert_offset=2 * rank, nc_total=4).sharded
    open(os.path.join(tmp, "ok%d" % rank), "w").write("1" if ok else "0")
    dist.destroy_process_group()


def test_sharded_aggregation_gloo_world2(tmp_path):
    """Two CPU ranks (gloo): shard the experts, all-reduce the partial sums, every rank recovers the oracle's aggregate."""
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert [open(os.path.join(str(tmp_path), "ok%d" % r)).read() for r in range(2)] == ["1", "1"]


def test_bench_reference_arm_contract():
    """bench.py --impl reference (the CPU arm the driver runs beside ours) prints ONE JSON line with the contract's keys;
    the smoke-sized workload keeps it to seconds here."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--small", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=300, cwd=root)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("GRBCM predictions/sec") and d["unit"] == "predictions/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["n_gpus"] == 1 and d["gpu_launches"] == 0
    from oracle import ref_import
    # the unmodified reference when a copy is importable (build container / baseline/_ref), the oracle port otherwise
    assert d["cpu_baseline"]["kind"] == ("reference" if ref_import.available() else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert set(d["config"]) == {"workload", "m_per_step", "l2"}      # same keys as the GPU arm's config (bench.config_of)


def test_cpu_arm_reference_and_port_agree():
    """oracle/cpu_arm.py: the unmodified reference driven through its public API in expert chunks and the oracle port give the
    same per-expert predictions, aggregate, NLML and gradient on a small GRBCM (skipped without a reference copy)."""
    import math
    import pytest
    from oracle import cpu_arm, ref_import
    if not ref_import.available():
        pytest.skip("no reference copy present")
    g = torch.Generator().manual_seed(0)
    M, m0, dd = 6, 200, 4
    x = torch.rand(M, m0, dd, generator=g, dtype=torch.float64)
    y = torch.sin(x.sum(-1)) + 0.1 * torch.randn(M, m0, generator=g, dtype=torch.float64)
    hp = torch.empty(4, dd + 2, dtype=torch.float64)
    hp[:, 0] = 1.0; hp[:, 1:dd + 1] = 2.0 / math.sqrt(dd); hp[:, dd + 1] = 0.3
    xs = torch.rand(50, dd, generator=g, dtype=torch.float64)
    outs = {}
    for use in (True, False):
        cs = cpu_arm.CpuSample(x[1:5], y[1:5], x[0], y[0], hp, hp[0], use_reference=use)
        cs.fit()
        outs[cs.kind] = (cs.predict(xs), cs.nlml(), cs.grad_comm())
    a, b = outs["reference"], outs["port"]
    for k in a[0]:
        assert cpu_arm.rel_err(a[0][k], b[0][k]) <= 1e-13, k
    assert abs(a[1][0] - b[1][0]) <= 1e-12 * abs(b[1][0]) and cpu_arm.rel_err(a[1][1], b[1][1]) <= 1e-13
    assert cpu_arm.rel_err(a[2], b[2]) <= 1e-12


def test_shared_block_detection_and_dense_routing_are_host_logic():
    """Path selection needs no device: GRBCM._shares_hyperparameters (one vector for the whole committee, at least one 128-row
    block in the communication set), engine.fused_dim_ok / covar.fused_spec (which covariances and input widths the on-chip
    tile generation covers; everything else goes the dense path)."""
    import pygpr_b200 as P
    from pygpr_b200 import engine
    from pygpr_b200.covar import fused_spec
    cov = P.Compose([P.Squared_exponential(), P.White_noise()])
    xl, yl = torch.rand(4, 50, 3), torch.rand(4, 50)
    big = P.GRBCM(xl, yl, torch.rand(130, 3), torch.rand(130), cov)
    small = P.GRBCM(xl, yl, torch.rand(100, 3), torch.rand(100), cov)
    assert big._shares_hyperparameters()                 # default init: identical rows (covar.py:96-100,221-225)
    assert not small._shares_hyperparameters()           # communication set smaller than one block
    hp = big.gpl.params.clone(); hp[2, 1] = 1.5
    big.gpl.set_params(hp)
    assert not big._shares_hyperparameters()             # per-expert vectors
    big.gpl.set_params(big.gpg.params.repeat(4, 1))
    assert big._shares_hyperparameters()
    from pygpr_b200 import config
    config.options.share = False
    try:
        assert not big._shares_hyperparameters()
    finally:
        config.options.share = True
    assert engine.fused_dim_ok(16) and engine.fused_dim_ok(44) and not engine.fused_dim_ok(45) and not engine.fused_dim_ok(64)
    assert fused_spec(cov, 16) == (True,) and fused_spec(P.Squared_exponential(), 8) == (False,) and fused_spec(cov, 64) is None
    assert fused_spec(P.Compose([P.Squared_exponential(), P.Squared_exponential(), P.White_noise()]), 4) is None


def test_ozaki_reference_arithmetic():
    """oracle/ozaki_ref.py (the numpy restatement the GPU digit planes are checked against bit for bit): balanced radix-256
    digits in [-128, 127], leading digit within [-64, 64], exact reconstruction of the fixed-point value, known answers, and the
    error of the recombined product falling by 2^-8 per slice."""
    import numpy as np
    from oracle import ozaki_ref
    rng = np.random.default_rng(0)
    X = rng.standard_normal((5, 70)) * np.exp(3 * rng.standard_normal((5, 1)))
    X[3] = 0.0
    for S in (1, 4, 7, 8):
        planes, scale = ozaki_ref.split(X, S)
        assert planes.dtype == np.int8 and planes.shape == (S, 128, 128) and np.abs(planes[0].astype(int)).max() <= 64
        assert scale[3] == 0.0 and not planes[:, 3].any() and not planes[:, 5:].any() and not planes[:, :, 70:].any()
        e = np.log2(scale[:5][scale[:5] > 0]) + 8
        assert np.all(e == np.round(e))                                      # power-of-two scales
        # value = 2^e sum_t q_t 256^-(t+1), exact to half a unit of the last digit
        rec = sum(planes[t, :5, :70].astype(np.float64) * 256.0 ** -(t + 1) for t in range(S)) * (scale[:5, None] * 256.0)
        assert np.all(np.abs(rec - X) <= 0.5 * 256.0 ** -S * (scale[:5, None] * 256.0) * (1 + 1e-12))
    # known answers: 0.5 = 2^2 * 32/256 -> leading digit 32 at exponent 2 (|x| 2^-e < 1/4); -0.75 -> -48
    p, sc = ozaki_ref.split(np.array([[0.5, -0.25, 0.0]]), 2)
    assert sc[0] == 2.0 ** (2 - 8) and p[:, 0, :3].tolist() == [[32, -16, 0], [0, 0, 0]]
    p, sc = ozaki_ref.split(np.array([[-0.75, 0.75 / 256 + 0.75 / 65536]]), 3)
    assert p[:, 0, :2].tolist() == [[-48, 0], [0, 48], [0, 48]]
    A, B = rng.standard_normal((40, 300)), rng.standard_normal((24, 300))
    ref = A @ B.T
    errs = []
    for S in (3, 4, 5, 6, 7):
        pa, sa = ozaki_ref.split(A, S); pb, sb = ozaki_ref.split(B, S)
        errs.append(np.abs(ozaki_ref.gemm_nt(pa, sa, 40, pb, sb, 24, S) - ref).max() / np.abs(ref).max())
    assert errs[-1] < 1e-13 and all(b < a / 50 for a, b in zip(errs, errs[1:]))


def test_emulated_form_heuristics_and_mle_lifetime():
    """Host logic added in round 2: where the emulated factorisation is taken (engine.fit_emulated_pays), and GRBCM_MLE holds no
    reference cycle (a dropped instance — and its HBM-sized factor buffer — must go away without a cyclic GC pass)."""
    import gc
    import weakref
    import torch
    import pygpr_b200 as P
    from pygpr_b200 import engine
    assert engine.fit_emulated_pays(1, 16384) and engine.fit_emulated_pays(1, 8192) and not engine.fit_emulated_pays(1, 4096)
    assert engine.fit_emulated_pays(4, 4096) and not engine.fit_emulated_pays(3, 4096) and not engine.fit_emulated_pays(64, 2048)
    assert engine.can_fit_emulated(1000, None) and not engine.can_fit_emulated(300, None)
    xl, yl = torch.rand(3, 20, 2, dtype=torch.float64), torch.rand(3, 20, dtype=torch.float64)
    xg, yg = torch.rand(10, 2, dtype=torch.float64), torch.rand(10, dtype=torch.float64)
    model = P.GRBCM(xl, yl, xg, yg, P.Compose([P.Squared_exponential(), P.White_noise()]))
    was_enabled = gc.isenabled()
    gc.disable()
    try:
        mle = P.GRBCM_MLE(model, chunk=2)
        assert mle.model is mle                               # the duck type optimisers rely on (loss.model.params / set_params)
        ref = weakref.ref(mle)
        del mle
        assert ref() is None
    finally:
        if was_enabled:
            gc.enable()


def test_digit_planes_without_the_carry_loop():
    """The fused K* builder (se_cross_planes_kernel) forms the balanced radix-256 digits as the bytes of (I + 0x80..80) ^ 0x80..80;
    oracle/ozaki_ref.py restates both forms: they must agree digit for digit, and a known bound must give the planes of split()."""
    import numpy as np
    from oracle import ozaki_ref
    rng = np.random.default_rng(5)
    for slices in (1, 3, 6, 7, 8):
        lim = 1 << (8 * slices - 2)
        I = rng.integers(-lim + 1, lim, size=4000, dtype=np.int64)
        I[:4] = (0, lim - 1, -lim + 1, -1)
        fast = ozaki_ref.digits_by_bias(I, slices)
        J = I.copy()
        for t in range(slices - 1, -1, -1):                   # the carry loop of oz::split_kernel
            dgt = (J & 0xFF).astype(np.uint8).view(np.int8)
            assert np.array_equal(fast[t], dgt)
            J = (J - dgt.astype(np.int64)) >> 8
        assert not J.any()
    X = rng.random((5, 70)) * 1.7                              # 0 <= x < bound
    planes, scale = ozaki_ref.split(X, 6, bound=1.7)
    e = 0 + 3                                                  # ilogb(1.7) + 3
    assert np.all(scale[:5] == np.ldexp(1.0, e - 8)) and np.all(scale[5:] == 0)
    Iref = np.rint(X * np.ldexp(1.0, 48 - e)).astype(np.int64)
    assert np.array_equal(planes[:, :5, :70], ozaki_ref.digits_by_bias(Iref, 6))
    back = sum(planes[t, :5, :70].astype(np.float64) * 256.0 ** -(t + 1) for t in range(6)) * np.ldexp(1.0, e)
    assert np.max(np.abs(back - X)) <= np.ldexp(1.0, e - 48)


def test_tools_and_entry_points_compile():
    """The measurement tools are run by hand on the GPU box: keep them at least syntactically alive on every CPU run."""
    import glob
    import os
    import py_compile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    files = glob.glob(os.path.join(root, "tools", "*.py")) + [os.path.join(root, f) for f in ("bench.py", "__graft_entry__.py")]
    assert len(files) > 10
    for f in files:
        py_compile.compile(f, doraise=True)
