"""Generalized Robust Bayesian Committee Machine with the reference's API (PyGPR/gr_bcm.py)."""
from __future__ import annotations

import torch

from . import dist as gdist
from . import engine
from .covar import _back
from .gpr import GPR, Exact_GP


class GRBCM(GPR):
    """GRBCM (arXiv:1806.00720): ``nc`` local experts, each trained on the communication set (xg, yg) plus its own
    block (gr_bcm.py:19-29), and one communication expert on (xg, yg) alone; predictions are combined with
    the precision weights of gr_bcm.py:116-149.

    Same constructor and attributes as the reference (``gpg, gpl, nc, nsc, ng, dim, beta, prec``).  Extra,
    optional keyword arguments drive expert sharding over GPUs (one process per GPU):

    * ``shard=True``      — ``xl, yl`` hold ALL experts on every rank; this rank keeps its contiguous block;
    * ``expert_offset=k, nc_total=N`` — ``xl, yl`` already hold only this rank's block, which starts at global
      expert ``k`` of ``N``;
    * ``group``           — the torch.distributed process group (default: world).

    ``predict`` then all-reduces the ``[3, m]`` partial sums (NCCL) and returns identical results on all ranks.
    """

    def __init__(self, xl, yl, xg, yg, cov, hp=None, **kargs):
        self.group = kargs.get("group", None)
        rank, ws = gdist.world(self.group)
        nc_given = xl.shape[0]
        if kargs.get("shard", False) and ws > 1:
            lo, hi = gdist.shard_bounds(nc_given, ws, rank)
            xl, yl = xl[lo:hi], yl[lo:hi]
            self.expert_offset, self.nc_total = lo, nc_given
        else:
            self.expert_offset = int(kargs.get("expert_offset", 0))
            self.nc_total = int(kargs.get("nc_total", nc_given))
        self.sharded = self.nc_total != xl.shape[0]

        nc = xl.shape[0]
        nls = xl.shape[1]
        ng = xg.shape[0]
        dim = xg.shape[1]

        # communication set FIRST, then the local block (gr_bcm.py:19-26)
        x = torch.cat((xg.to(xl.device)[None].expand(nc, ng, dim), xl), dim=1)
        y = torch.cat((yg.to(yl.device)[None].expand(nc, ng), yl), dim=1)

        self.gpg = Exact_GP(xg, yg, cov)                      # gr_bcm.py:28
        self.gpl = Exact_GP(x, y, cov)                        # gr_bcm.py:29

        self.nc = nc
        self.nsc = nls
        self.ng = ng
        self.dim = dim
        self.beta = NotImplemented
        self.prec = NotImplemented

    # ---------------------------------------------------------------------------------------------
    def aggregate(self, ys_g, covars_g, ys_l, covars_l, var="diag"):
        """Precision-weighted combination (gr_bcm.py:116-149).  With sharding the three sums over the expert
        axis are formed per rank and all-reduced; ``beta``/``prec`` then hold row 0 + this rank's rows."""
        if var != "diag":
            return self._aggregate_full(ys_g, covars_g, ys_l, covars_l)
        out_like = ys_g
        mean_g, var_g = engine._dev(ys_g).reshape(-1), engine._dev(covars_g).reshape(-1)
        m = mean_g.shape[0]
        mean_l = engine._dev(ys_l).reshape(-1, m)
        var_l = engine._dev(covars_l).reshape(-1, m)
        S, brow, prow = engine.grbcm_partial(mean_l, var_l, var_g, self.expert_offset, want_rows=True)
        gdist.all_reduce_sum_(S, self.group)                  # the only data-path collective: [3, m] doubles
        mu, v, b0, p0 = engine.grbcm_finalize(S, mean_g, var_g, want_rows=True)
        self.beta = _back(torch.cat((b0[None], brow)), out_like)   # gr_bcm.py:135-136
        self.prec = _back(torch.cat((p0[None], prow)), out_like)
        return _back(mu, out_like), _back(v, out_like)

    def _aggregate_full(self, ys_g, covars_g, ys_l, covars_l):
        return engine.grbcm_aggregate_full(self, ys_g, covars_g, ys_l, covars_l)

    def predict(self, xs, var="diag"):
        ys_g, covars_g = self.gpg.predict(xs, var=var)        # gr_bcm.py:152
        ys_l, covars_l = self.gpl.predict(xs, var=var)        # gr_bcm.py:153
        if self.nc == 1:
            ys_l = ys_l.reshape(1, -1)
        return self.aggregate(ys_g, covars_g, ys_l, covars_l, var=var)  # gr_bcm.py:155

    def gather_weights(self):
        """(beta, prec) for ALL experts [nc_total+1, m] when sharded (all-gather of the local rows)."""
        if not self.sharded:
            return self.beta, self.prec
        _, ws = gdist.world(self.group)
        counts = [gdist.shard_bounds(self.nc_total, ws, r) for r in range(ws)]
        counts = [hi - lo for lo, hi in counts]
        out = []
        for t in (self.beta, self.prec):
            td = engine._dev(t)
            out.append(_back(torch.cat((td[:1], gdist.all_gather_rows(td[1:].contiguous(), counts, self.group))), t))
        return tuple(out)


def log_likelihood_batched(x, y, hp, cov, **kwargs):
    """Batched NLML of experts (x:[nc,n,d], y:[nc,n], hp:[nc,P]) -> numpy [nc].

    The reference's function of this name (gr_bcm.py:158-176) cannot run (it calls ``cov`` as a function and
    drops two of the three terms); this is the quantity it was meant to return, computed like MLE.loss.
    """
    from .loss import MLE
    model = Exact_GP(x, y, cov)
    return MLE(model).loss(hp.detach().cpu().numpy() if isinstance(hp, torch.Tensor) else hp)
