"""Generate golden input/output vectors by running the UNMODIFIED reference (build container).

    python oracle/gen_golden.py            # writes tests/golden/*.npz

TEST INFRASTRUCTURE.  The reference is a Python package that cannot travel to the GPU box, so the
vectors it produces on seeded inputs are committed as small fixtures; ``tests/test_oracle_golden.py``
pins the oracle to them and the ``-m gpu`` tests pin the CUDA path to them.  Inputs are stored in
the fixture (not re-derived from a seed) so RNG differences between boxes cannot matter.
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_import  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def _np(t):
    return t.detach().cpu().numpy().copy() if isinstance(t, torch.Tensor) else np.asarray(t)


def make_data(seed, B, n, d, m, lo=0.0, hi=1.0, noise=0.1):
    g = torch.Generator().manual_seed(seed)
    x = lo + (hi - lo) * torch.rand(B, n, d, generator=g, dtype=torch.float64)
    y = torch.sin(x.sum(-1)) + noise * torch.randn(B, n, generator=g, dtype=torch.float64)
    xs = lo + (hi - lo) * torch.rand(m, d, generator=g, dtype=torch.float64)
    return x, y, xs


def hp_well(B, d, sig_n=0.3, jitter_seed=None):
    hp = torch.empty(B, d + 2, dtype=torch.float64)
    hp[:, 0] = 1.0
    hp[:, 1:d + 1] = 2.0 / np.sqrt(d)
    hp[:, d + 1] = sig_n
    if jitter_seed is not None:
        g = torch.Generator().manual_seed(jitter_seed)
        hp = hp * (0.5 + torch.rand(B, d + 2, generator=g, dtype=torch.float64))
    return hp


ONLY = sys.argv[1:]   # optional name filters: regenerate only the matching fixtures (the others stay byte-identical)


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = ref_import.load()
    warnings.simplefilter("ignore")
    SE, WN, Compose = ref.Squared_exponential, ref.White_noise, ref.Compose

    # ---- covariance kernels: SE, WN, SE+WN, SE+SE+WN; batched and unbatched; cross kernels ----
    for name, B, n, d, m, keep_dk in [("covar_b1_n37_d3", 1, 37, 3, 11, True), ("covar_b4_n130_d5", 4, 130, 5, 29, False),
                                      ("covar_b3_n40_d16", 3, 40, 16, 150, True)]:
        if ONLY and not any(o in name for o in ONLY):
            continue
        x, _, xs = make_data(11, B, n, d, m)
        out = {"x": _np(x), "xs": _np(xs)}
        g = torch.Generator().manual_seed(5)
        for tag, covs in [("se", [SE()]), ("se_wn", [SE(), WN()]), ("se_se_wn", [SE(), SE(), WN()])]:
            cov = Compose(covs)
            xx = x.clone().squeeze(0) if B == 1 else x.clone()
            P = cov.get_params_shape(xx)[-1]
            hp = 0.5 + torch.rand(B, P, generator=g, dtype=torch.float64)
            hpp = hp.clone().squeeze(0) if B == 1 else hp.clone()
            out[tag + "_hp"] = _np(hp)
            out[tag + "_K"] = _np(cov.kernel(hpp.clone(), xx.clone()))
            out[tag + "_Ks"] = _np(cov.kernel(hpp.clone(), xx.clone(), xs.clone()))
            K2, dK = cov.kernel_and_grad(hpp.clone(), xx.clone())
            if keep_dk:
                out[tag + "_dK"] = _np(dK)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)

    # ---- exact GP fit / predict / NLML / gradient / learn rate ----
    cases = [
        # name, B, n, d, m, (lo,hi), hp kind
        ("gp_cfg1_n1000_d1_well", 1, 1000, 1, 256, (0.0, 10.0), "well"),
        ("gp_cfg1_n1000_d1_default", 1, 1000, 1, 256, (0.0, 10.0), "default"),
        ("gp_b1_n300_d4", 1, 300, 4, 77, (0.0, 1.0), "well"),
        ("gp_b5_n200_d3_perexpert", 5, 200, 3, 64, (0.0, 1.0), "jit"),
        ("gp_b3_n384_d8", 3, 384, 8, 130, (0.0, 1.0), "well"),
    ]
    for name, B, n, d, m, (lo, hi), kind in cases:
        if ONLY and not any(o in name for o in ONLY):
            continue
        x, y, xs = make_data(3, B, n, d, m, lo, hi)
        if lo == 0.0 and hi == 10.0:
            y = torch.sin(x.sum(-1))  # config 1: y = sin(x), no noise
        cov = Compose([SE(), WN()])
        if kind == "default":
            hp = cov.init_params(x)
        elif kind == "jit":
            hp = hp_well(B, d, 0.3, jitter_seed=9)
        else:
            hp = hp_well(B, d, 0.3)
        sq = (lambda t: t.clone().squeeze(0)) if B == 1 else (lambda t: t.clone())
        gp = ref.Exact_GP(sq(x), sq(y), cov)
        gp.set_params(sq(hp))
        mean, var = gp.predict(xs.clone(), var="diag")
        out = {"x": _np(x), "y": _np(y), "xs": _np(xs), "hp": _np(hp),
               "mean": _np(mean), "var": _np(var), "alpha": _np(gp.wt), "L": _np(gp.krnchd) if n <= 300 else np.zeros(0)}
        if n <= 300:
            mean_f, cov_f = gp.predict(xs.clone(), var="full")
            out["cov_full"] = _np(cov_f)
        mle = ref.MLE(gp)
        out["nlml"] = np.array(mle.loss(_np(sq(hp)).copy()))
        l2, g2 = mle.loss_and_grad(_np(sq(hp)).copy())
        out["nlml2"] = np.array(l2); out["grad"] = np.array(g2)
        out["grad_only"] = np.array(mle.grad(_np(sq(hp)).copy()))
        import PyGPR.hp_update as ref_hpu
        gamma = ref_hpu.get_learn_rate(sq(hp), mle, 1e-4)
        out["gamma"] = np.array(gamma)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)

    # ---- GRBCM ----
    for name, nc, nls, ng, d, m, kind in [("grbcm_nc4_nls64_ng32_d3", 4, 64, 32, 3, 50, "well"),
                                          ("grbcm_nc7_nls100_ng60_d5_perexpert", 7, 100, 60, 5, 40, "jit"),
                                          ("grbcm_nc3_nls200_ng184_d16", 3, 200, 184, 16, 33, "well"),
                                          # tile-aligned communication sets + one shared hyper-parameter vector: the cases
                                          # the shared-leading-block path of pygpr_b200 takes (gpr_potrf_batched_shared)
                                          ("grbcm_nc5_nls150_ng128_d4_sharedhp", 5, 150, 128, 4, 70, "well"),
                                          ("grbcm_nc6_nls200_ng256_d8_sharedhp", 6, 200, 256, 8, 45, "well")]:
        if ONLY and not any(o in name for o in ONLY):
            continue
        xl, yl, xs = make_data(21, nc, nls, d, m)
        xg, yg, _ = make_data(22, 1, ng, d, 1)
        xg, yg = xg[0], yg[0]
        cov = Compose([SE(), WN()])
        model = ref.GRBCM(xl.clone(), yl.clone(), xg.clone(), yg.clone(), cov)
        hp_l = hp_well(nc, d, 0.3, jitter_seed=4 if kind == "jit" else None)
        hp_g = hp_well(1, d, 0.3)[0]
        model.gpl.set_params(hp_l)
        model.gpg.set_params(hp_g)
        mu, var = model.predict(xs.clone(), var="diag")
        out = {"xl": _np(xl), "yl": _np(yl), "xg": _np(xg), "yg": _np(yg), "xs": _np(xs),
               "hp_l": _np(hp_l), "hp_g": _np(hp_g), "mu": _np(mu), "var": _np(var),
               "beta": _np(model.beta), "prec": _np(model.prec)}
        mu_f, cov_f = model.predict(xs.clone(), var="full")
        out["mu_full"] = _np(mu_f); out["cov_full"] = _np(cov_f)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)

    # ---- samplers / partitioning (sampler.py): UNIFORM, MATERN1.sample / cluster_samples / partition, euclidean_dist, cluster_samples ----
    name = "sampler_seed3_seed5"
    if not ONLY or any(o in name for o in ONLY):
        mins, maxs = torch.zeros(2), torch.ones(2) * 2.0
        xu = ref.UNIFORM(3).sample(500, mins, maxs)
        mat = ref.MATERN1(5)
        parts, centres = mat.partition(6, 20, mins, maxs)
        dist = ref.euclidean_dist(xu, centres)
        # cluster_samples needs equal-sized cells: a 2 x 2 grid of centres and 25 points per cell drawn around them
        g = torch.Generator().manual_seed(9)
        grid = torch.tensor([[0.5, 0.5], [0.5, 1.5], [1.5, 0.5], [1.5, 1.5]])
        pts = (grid[:, None, :] + 0.4 * (torch.rand(4, 25, 2, generator=g, dtype=torch.float64) - 0.5)).reshape(100, 2)
        pts = pts[torch.randperm(100, generator=g)]
        cells = ref.cluster_samples(pts, grid)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), mins=_np(mins), maxs=_np(maxs), uniform=_np(xu), parts=_np(parts),
                            centres=_np(centres), min_dist=np.float64(float(mat.min_dist)), dist=_np(dist), grid=_np(grid), pts=_np(pts),
                            cells=_np(cells))

    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
