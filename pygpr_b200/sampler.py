"""Data generation and partitioning with the reference's API (PyGPR/sampler.py) — SURVEY §8(f) N3.

Random draws come from torch's CPU generator seeded exactly like the reference (``tc.manual_seed(seed)`` then
``tc.rand``), so the same seed gives the same points.  The nearest-centre assignment (the O(n nc d) part) runs on the
GPU through ``gpr_sq_dist`` (GEMM-form distance, cross term on FP64 tensor cores) — the same tile as the kernel matrix.
"""
from __future__ import annotations

import torch

from . import _lib, engine
from .covar import _back


def euclidean_dist(x, y):
    """Squared Euclidean distances [n, nc] in GEMM form (sampler.py:94-100)."""
    xd, yd = engine._dev(x), engine._dev(y)
    n, m, d = xd.shape[0], yd.shape[0], xd.shape[1]
    ones = torch.ones(1, d + 1, dtype=torch.float64, device=xd.device)   # [sig, ls...] = 1: plain distances
    ax, nx = engine.scale_inputs(xd, ones, 1)
    ay, ny = engine.scale_inputs(yd, ones, 1)
    D = torch.empty(n, m, dtype=torch.float64, device=xd.device)
    _lib.call("gpr_sq_dist", engine._p(ax), engine._p(nx), ax.shape[1], engine._p(ay), engine._p(ny), ay.shape[1],
              n, m, ax.shape[2], engine._p(D), engine._stream())
    engine._count(1)
    return _back(D, x)


def _nearest_centre(x, xc):
    return torch.argmin(euclidean_dist(x, xc), 1)


def cluster_samples(x, xc):
    """Partition x [n, d] by nearest centre xc [nc, d] into [nc, n/nc, d]; needs equal-sized cells (sampler.py:103-119)."""
    n, nc = x.shape[-2], xc.shape[-2]
    assert n % nc == 0
    idx = _nearest_centre(x, xc)
    xpart = torch.empty([nc, n // nc, x.shape[-1]], dtype=x.dtype, device=x.device)
    for i in range(nc):
        xpart[i, :, :] = x[idx == i, :]
    return xpart


class UNIFORM(object):
    """Uniform points in a box (sampler.py:6-14)."""

    def __init__(self, seed):
        self.seed = seed

    def sample(self, n, mins, maxs):
        torch.manual_seed(self.seed)
        return mins + torch.rand(n, len(mins)).mul_(maxs - mins)


class MATERN1(UNIFORM):
    """Matern type-I hard-core (repulsive) centres and nearest-centre cells around them (sampler.py:17-91)."""

    def __init__(self, seed):
        super().__init__(seed)
        self.min_dist = None
        self.max_count = 5000

    def sample_repulsion(self, mins, maxs, min_dist):
        """Sequential rejection: keep a uniform draw only if it is farther than min_dist from every kept point; stop
        after max_count points or max_count consecutive rejections (sampler.py:23-47).  Inherently serial and tiny —
        it stays on the host with the reference's RNG stream."""
        torch.manual_seed(self.seed)
        dim = len(mins)
        kept = torch.empty([self.max_count, dim])
        kept[0, :] = mins + torch.rand(1, dim).mul_(maxs - mins)
        k, misses = 1, 1
        while k < self.max_count and misses < self.max_count:
            cand = mins + torch.rand(1, dim).mul_(maxs - mins)
            gap = (kept[:k, :] - cand).square_().sum(1).sqrt_()
            if torch.all(gap.sub_(min_dist) > 1e-5):
                kept[k, :] = cand
                k += 1
                misses = 0
            misses += 1
        return kept[:k, :]

    def sample(self, n, mins, maxs):
        vol = torch.prod(maxs - mins)
        min_dist = (vol / n) ** (1 / len(mins))
        xc = self.sample_repulsion(mins, maxs, min_dist)
        while xc.shape[0] < n:                      # shrink the exclusion radius until n centres fit (sampler.py:56-58)
            min_dist *= 0.9
            xc = self.sample_repulsion(mins, maxs, min_dist)
        self.min_dist = min_dist
        return xc[:n, :]

    def cluster_samples(self, xc, ns, mins, maxs):
        """ns points per centre: draw 10 ns nc uniform points, keep the first ns of each nearest-centre cell."""
        torch.manual_seed(self.seed)
        nc, dim = xc.shape
        x = mins + torch.rand(10 * ns * nc, dim).mul_(maxs - mins)
        idx = _nearest_centre(x, xc)
        xpart = torch.empty([nc, ns, dim])
        for i in range(nc):
            xpart[i, :, :] = x[idx == i][:ns, :]
        return xpart

    def partition(self, nc, ns, mins, maxs):
        xc = self.sample(nc, mins, maxs)
        return self.cluster_samples(xc, ns, mins, maxs), xc


def sample_gp(x, cov, hp=None, mean=None, **kwargs):
    """One draw f ~ N(mean, K(x,x)) via the GPU Cholesky.  The reference's version (sampler.py:122-137) calls ``cov`` as a
    function and cannot run; this is the working equivalent on the Covar protocol (``hp`` defaults to
    ``cov.init_params(x)``, a 1e-7 jitter keeps K factorisable as in gpr.py:68)."""
    if hp is None:
        hp = cov.init_params(x)
    K = engine._dev(cov.kernel(hp, x)).reshape(-1, x.shape[-2], x.shape[-2]).clone()
    K.diagonal(dim1=-2, dim2=-1).add_(engine.JITTER)
    n = K.shape[-1]
    npad = engine.round_up(n, engine.TILE)
    buf = torch.zeros(K.shape[0], npad, npad, dtype=torch.float64, device=K.device)
    buf[:, :n, :n] = K
    if npad > n:
        i = torch.arange(n, npad, device=K.device)
        buf[:, i, i] = 1.0
    _, _, _, _, info = engine.potrf_inplace(buf)
    if int(info.abs().sum()):
        raise torch.linalg.LinAlgError("sample_gp: covariance is not positive-definite")
    L = torch.tril(buf[:, :n, :n])
    noise = torch.randn(n).to(L.device, torch.float64)
    f = torch.matmul(L, noise).squeeze(0)
    if mean is not None:
        f = f + engine._dev(mean)
    return _back(f, x)
