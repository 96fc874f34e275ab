"""CPU arm of bench.py: the reference's own implementation of the path on the host cores — TEST INFRASTRUCTURE.

Used only by ``bench.py --impl reference`` and by the ``cpu_baseline`` / ``parity`` legs of the default arm.  When an unmodified
copy of the reference is importable (``oracle/ref_import.py``: /root/reference in the build container, ``baseline/_ref`` on the
GPU box) it is driven through its public API, exactly as a user would (kind = "reference"):

    GRBCM(xl, yl, xg, yg, cov)                     gr_bcm.py:12-34
    model.gpl.set_params / model.gpg.set_params    gpr.py:28-31
    model.gpl.predict(xs, var="diag")              gpr.py:76-106        (first call factorises, gpr.py:65-74)
    model.aggregate(ys_g, cov_g, ys_l, cov_l)      gr_bcm.py:116-149
    MLE(model.gpl).loss / loss_and_grad            loss.py:35-57, 92-128

Otherwise the oracle port (``oracle/pygpr_oracle.py``) does the same work (kind = "port").  The reference needs O(nc N^2)
memory and minutes for the whole job, so the local experts are taken in chunks of two (a GRBCM of one expert raises upstream)
and only a bounded sample is run; experts are equal-sized and independent, so the whole-job rate follows from the N^2 work
ratio.
"""
from __future__ import annotations

import math
import time
from typing import Dict

import numpy as np
import torch

from . import pygpr_oracle as orc
from . import ref_import

F64 = torch.float64
SPEC = ["se", "wn"]


class CpuSample:
    """Fitted sample (nc_s local experts + communication expert) that can predict repeatedly with cached factors."""

    def __init__(self, xl, yl, xg, yg, hp_l, hp_g, use_reference: bool = True, chunk: int = 2):
        self.kind = "reference" if (use_reference and ref_import.available()) else "port"
        self.xl, self.yl, self.xg, self.yg, self.hp_l, self.hp_g = xl, yl, xg, yg, hp_l, hp_g
        self.nc_s = xl.shape[0]
        self.chunk = max(2, chunk)
        assert self.nc_s % self.chunk == 0
        self.fit_s = None
        if self.kind == "reference":
            R = ref_import.load()
            self.R = R
            cov = R.Compose([R.Squared_exponential(), R.White_noise()])
            self.models = []
            for s in range(0, self.nc_s, self.chunk):
                m = R.GRBCM(xl[s:s + self.chunk].clone(), yl[s:s + self.chunk].clone(), xg.clone(), yg.clone(), cov)
                m.gpl.set_params(hp_l[s:s + self.chunk])
                m.gpg.set_params(hp_g)
                self.models.append(m)
            # aggregate() sizes its weight arrays by self.nc (gr_bcm.py:125-126): an un-fitted GRBCM over all sample experts
            # (the constructor only concatenates inputs) does the aggregation of the chunked predictions
            self.agg = R.GRBCM(xl.clone(), yl.clone(), xg.clone(), yg.clone(), cov)

    def fit(self) -> float:
        """The factorisations GRBCM.predict triggers on first use (gr_bcm.py:152-153 -> gpr.py:65-74); seconds."""
        t0 = time.perf_counter()
        if self.kind == "reference":
            self.models[0].gpg.update()
            for m in self.models:
                m.gpl.update()
        else:
            self.fits = orc.grbcm_fit(SPEC, self.hp_g, self.hp_l, self.xl, self.yl, self.xg, self.yg, expert_chunk=self.chunk)
        self.fit_s = time.perf_counter() - t0
        return self.fit_s

    def predict(self, xs) -> Dict[str, torch.Tensor]:
        """One step: both predictions with resident factors + the aggregation over the sample's experts."""
        if self.kind == "reference":
            ys_g, cov_g = self.models[0].gpg.predict(xs, var="diag")
            outs = [m.gpl.predict(xs, var="diag") for m in self.models]
            ys_l = torch.cat([o[0].reshape(self.chunk, -1) for o in outs])
            cov_l = torch.cat([o[1].reshape(self.chunk, -1) for o in outs])
            mean_g, var_g = ys_g.clone(), cov_g.clone()      # aggregate() overwrites its covars_g argument (gr_bcm.py:128)
            mu, var = self.agg.aggregate(ys_g, cov_g, ys_l, cov_l, var="diag")
            return {"mean_g": mean_g, "var_g": var_g, "mean_l": ys_l.clone(), "var_l": cov_l.clone(), "mu": mu.clone(), "var": var.clone()}
        mu, var, parts = orc.grbcm_predict(SPEC, self.hp_g, self.hp_l, self.xl, self.yl, self.xg, self.yg, xs,
                                           expert_chunk=self.chunk, fits=self.fits)
        return {"mean_g": parts["mean_g"], "var_g": parts["var_g"], "mean_l": parts["mean_l"], "var_l": parts["var_l"], "mu": mu, "var": var}

    def nlml(self):
        """NLML of the communication expert and of every sample expert (loss.py:35-57); refits, like the reference."""
        if self.kind == "reference":
            R = self.R
            g = float(R.MLE(self.models[0].gpg).loss(self.hp_g.numpy()))
            l = np.concatenate([np.atleast_1d(R.MLE(m.gpl).loss(self.hp_l[i * self.chunk:(i + 1) * self.chunk].numpy()))
                                for i, m in enumerate(self.models)])
            return g, l
        x, y = orc.grbcm_inputs(self.xl, self.yl, self.xg, self.yg)
        l = torch.cat([orc.mle_loss(SPEC, self.hp_l[s:s + 1], x[s:s + 1], y[s:s + 1]) for s in range(self.nc_s)])
        g = orc.mle_loss(SPEC, self.hp_g[None], self.xg[None], self.yg[None])
        return float(g[0]), l.numpy()

    def loss_and_grad_sum(self):
        """GRBCM training objective at the shared vector: NLML_g + sum_i NLML_i and its gradient (loss.py:92-128 per model; the sum
        is what gr_bcm.py:36-97 was written to minimise) — for the config-5 parity at reduced N."""
        if self.kind == "reference":
            R = self.R
            f, j = R.MLE(self.models[0].gpg).loss_and_grad(self.hp_g.numpy())
            f, j = float(f), np.asarray(j, dtype=np.float64).reshape(-1).copy()
            for i, m in enumerate(self.models):
                l, g = R.MLE(m.gpl).loss_and_grad(self.hp_l[i * self.chunk:(i + 1) * self.chunk].numpy())
                f += float(np.sum(l))
                j += np.asarray(g, dtype=np.float64).reshape(self.chunk, -1).sum(0)
            return f, j
        x, y = orc.grbcm_inputs(self.xl, self.yl, self.xg, self.yg)
        l, g = orc.mle_loss_and_grad_chunked(SPEC, self.hp_l, x, y, chunk=1)
        lg, gg = orc.mle_loss_and_grad(SPEC, self.hp_g[None], self.xg[None], self.yg[None])
        return float(l.sum() + lg.sum()), (g.sum(0) + gg[0]).numpy()

    def grad_comm(self):
        """NLML gradient of the communication expert (loss.py:92-128) — order ng <= 4096 keeps the reference's dK in RAM."""
        if self.kind == "reference":
            R = self.R
            _, j = R.MLE(self.models[0].gpg).loss_and_grad(self.hp_g.numpy())
            return np.asarray(j, dtype=np.float64).reshape(-1)
        _, j = orc.mle_loss_and_grad(SPEC, self.hp_g[None], self.xg[None], self.yg[None])
        return j[0].numpy()


def rel_err(a, b) -> float:
    """max|a - b| / max|b| — the parity metric of SURVEY 8(d)."""
    a = torch.as_tensor(np.asarray(a), dtype=F64).reshape(-1)
    b = torch.as_tensor(np.asarray(b), dtype=F64).reshape(-1)
    return float((a - b).abs().max() / b.abs().max())
