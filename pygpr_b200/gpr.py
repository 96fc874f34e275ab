"""GPR models with the reference's API (PyGPR/gpr.py), running on the sm_100a library."""
from __future__ import annotations

from typing import Optional, Sequence

import torch
from torch import Tensor

from . import engine
from .covar import Covar, fused_spec, _back


class GPR:
    """Base class for Gaussian process regression models (gpr.py:13-43)."""

    def __init__(self, x: Tensor, y: Tensor, cov: Covar) -> None:
        self.x: Tensor = x
        self.y: Tensor = y
        self.cov = cov
        self.params: Tensor = NotImplemented
        self.need_upd: bool = True

    def set_params(self, params: Tensor) -> None:
        self.params = torch.clone(params)                    # gpr.py:28-31
        self.need_upd = True

    def update(self) -> None:
        raise NotImplementedError

    def predict(self, xp: Tensor, var: str) -> Sequence[Tensor]:
        raise NotImplementedError

    def predict_var(self, xp: Tensor, **kwargs: Tensor) -> Tensor:
        raise NotImplementedError

    def predict_covar(self, xp: Tensor, **kwargs: Tensor) -> Tensor:
        raise NotImplementedError


def _unsupported(cov):
    return NotImplementedError(
        "pygpr_b200: the fused GPU path supports Squared_exponential or Compose([Squared_exponential(), "
        "White_noise()]); got %r. Kernel-level calls (kernel / kernel_and_grad) work for any composition." % (cov,))


class Exact_GP(GPR):
    """Exact GP (gpr.py:46-120).  ``x:[n,d]`` or ``[nc,n,d]`` (one expert per batch entry), ``y`` alike.

    ``update()`` factorises on the GPU (K is generated tile by tile inside the Cholesky, never stored);
    ``krnchd`` (lower factor), ``wt`` (= K^-1 y) and ``krn`` keep the reference's names and shapes and are
    materialised lazily on the device of ``x``.
    """

    def __init__(self, x: Tensor, y: Tensor, cov: Covar) -> None:
        super().__init__(x, y, cov)
        self.params: Tensor = cov.init_params(x)
        self._state: Optional[engine.FitState] = None
        self._krn = NotImplemented
        self._wt = NotImplemented
        self._krnchd = NotImplemented
        self.need_upd = True

    # -- reference attributes (gpr.py:57-59), filled lazily from the device state -------------------
    @property
    def batched(self) -> bool:
        return self.x.dim() > 2

    def _shape_out(self, t: Tensor) -> Tensor:
        return _back(t if self.batched else t.squeeze(0), self.x)

    @property
    def krnchd(self):
        if self._krnchd is NotImplemented and self._state is not None and not self.need_upd:
            n = self._state.n
            self._krnchd = self._shape_out(torch.tril(self._state.L[:, :n, :n]))
        return self._krnchd

    @krnchd.setter
    def krnchd(self, v):
        self._krnchd = v

    @property
    def wt(self):
        if self._wt is NotImplemented and self._state is not None and not self.need_upd:
            self._wt = self._shape_out(self._state.alpha[:, :self._state.n])
        return self._wt

    @wt.setter
    def wt(self, v):
        self._wt = v

    @property
    def krn(self):
        if self._krn is NotImplemented and self._state is not None and not self.need_upd:
            k = self.cov.kernel(self.params, self.x).clone()
            k.diagonal(dim1=-2, dim2=-1).add_(engine.JITTER)  # gpr.py:68
            self._krn = k
        return self._krn

    @krn.setter
    def krn(self, v):
        self._krn = v

    # -----------------------------------------------------------------------------------------------
    def _device_inputs(self):
        x = engine._dev(self.x)
        x = x.reshape(-1, x.shape[-2], x.shape[-1])
        y = engine._dev(self.y).reshape(-1, self.y.shape[-1])
        hp = engine._dev(self.params).reshape(-1, self.params.shape[-1])
        B = max(x.shape[0], hp.shape[0])
        if hp.shape[0] != B:
            hp = hp.expand(B, -1).contiguous()
        if x.shape[0] != B:
            x = x.expand(B, -1, -1).contiguous()
            y = y.expand(B, -1).contiguous()
        return x, y, hp

    def _launch_update(self, ws_tag: str = "", shared=None) -> None:
        """Queue the factorisation on the current stream without reading anything back (GRBCM overlaps two of them)."""
        x, y, hp = self._device_inputs()
        spec = fused_spec(self.cov, x.shape[-1])
        assert hp.shape[-1] == self.cov.get_params_shape(self.x)[-1]
        if spec is None:
            # any other Covar composition: K materialised by the kernel-level API, factorised in place (engine.fit_dense)
            K = self.cov.kernel(hp, x).reshape(-1, x.shape[1], x.shape[1]).clone()
            K.diagonal(dim1=-2, dim2=-1).add_(engine.JITTER)           # gpr.py:68
            self._state = None
            self._state = engine.fit_dense(K, y)
            self._dense_inputs = (x, hp)
            return
        # refit in place: the old factor's buffer (it can fill most of HBM) is handed back to fit()
        old, self._state = self._state, None
        npad = engine.round_up(x.shape[1], engine.TILE)
        L_buf = old.L if old is not None and tuple(old.L.shape) == (x.shape[0], npad, npad) and old.L.device == x.device else None
        del old
        if shared is not None and not engine.can_share(shared, x.shape[1], x.shape[2], spec[0]):
            shared = None
        self._state = engine.fit(x, y, hp, spec[0], L_out=L_buf, ws_tag=ws_tag, shared=shared)   # gpr.py:67-72

    def _finish_update(self) -> None:
        self._state.check()                                   # tc.cholesky raises on a non-PD matrix, gpr.py:69
        self._krn = self._wt = self._krnchd = NotImplemented
        self.need_upd = False

    def update(self) -> None:
        if self.need_upd:
            self._launch_update()
            self._finish_update()
        return None

    def _predict_dense(self, xs: Tensor, full: bool):
        x, hp = self._dense_inputs
        B, n, m = x.shape[0], x.shape[1], xs.shape[0]
        Ks = self.cov.kernel(hp, x, xs).reshape(B, m, n)                                    # gpr.py:79
        if full:
            Kss = self.cov.kernel(hp, xs.expand(B, -1, -1) if B > 1 else xs).reshape(B, m, m)  # gpr.py:110
            prior = torch.diagonal(Kss, dim1=-2, dim2=-1).contiguous()
        else:
            # gpr.py:98 builds the whole [m, m] matrix for its diagonal; here the diagonal is taken block by block (128 test
            # points at a time), PER TEST POINT: the Covar protocol also admits non-stationary kernels
            Kss = None
            prior = torch.empty(B, m, dtype=Ks.dtype, device=Ks.device)
            for s in range(0, m, 128):
                blk = xs[s:s + 128]
                kb = self.cov.kernel(hp, blk.expand(B, -1, -1) if B > 1 else blk).reshape(B, blk.shape[0], blk.shape[0])
                prior[:, s:s + 128] = torch.diagonal(kb, dim1=-2, dim2=-1)
        return engine.predict_dense(self._state, Ks, prior, Kss)

    def predict(self, xp: Tensor, var: str = "full") -> Sequence[Tensor]:
        self.update()
        xs = engine._dev(xp).reshape(-1, xp.shape[-1])
        if isinstance(self._state, engine.DenseState):
            mean, v, c = self._predict_dense(xs, var == "full")
            if var == "full":
                covars = _back(c.squeeze(0) if not self.batched else c, xp)
            elif var == "diag":
                covars = self._shape_out(v)
            else:
                covars = NotImplemented
            return [_back(mean.squeeze(), self.x), covars]
        if var == "full":
            mean, covars = self._predict_full(xs)
            covars = _back(covars.squeeze(0) if not self.batched else covars, xp)
        elif var == "diag":
            mean, v = engine.predict(self._state, xs, True)  # gpr.py:79-90
            covars = self._shape_out(v)
        else:
            mean, _ = engine.predict(self._state, xs, False)
            covars = NotImplemented                          # gpr.py:91-92
        ys = _back(mean.squeeze(), self.x)                   # gpr.py:85: squeeze_() drops every size-1 axis
        return [ys, covars]

    def predict_var(self, xp: Tensor, **kwargs: Tensor) -> Tensor:
        """Predictive variance at xp (gpr.py:96-106).  ``krns`` is accepted for signature parity and ignored:
        the cross-covariance tiles are regenerated on chip."""
        self.update()
        xs = engine._dev(xp).reshape(-1, xp.shape[-1])
        if isinstance(self._state, engine.DenseState):
            return self._shape_out(self._predict_dense(xs, False)[1])
        _, v = engine.predict(self._state, xs, True)
        return self._shape_out(v)

    def predict_covar(self, xp: Tensor, **kwargs: Tensor) -> Tensor:
        """Full predictive covariance (gpr.py:108-120)."""
        self.update()
        xs = engine._dev(xp).reshape(-1, xp.shape[-1])
        if isinstance(self._state, engine.DenseState):
            c = self._predict_dense(xs, True)[2]
            return _back(c.squeeze(0) if not self.batched else c, xp)
        _, c = self._predict_full(xs)
        return _back(c.squeeze(0) if not self.batched else c, xp)

    def _predict_full(self, xs: Tensor):
        return engine.predict_full(self._state, xs, self.cov)
