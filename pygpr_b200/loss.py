"""Model-selection losses with the reference's API (PyGPR/loss.py)."""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch
from numpy import ndarray

from . import engine
from .covar import fused_spec
from .gpr import GPR, _unsupported


class Loss:
    """Base class for cost functions for GP model selection (loss.py:10-28)."""

    def __init__(self, model: GPR) -> None:
        self.model: GPR = model
        self.loss_value: float = NotImplemented
        self.grad_value: ndarray = NotImplemented

    def loss(self, params: ndarray) -> float:
        raise NotImplementedError

    def grad(self, params: ndarray) -> ndarray:
        raise NotImplementedError

    def loss_and_grad(self, params: ndarray) -> Tuple[float, ndarray]:
        raise NotImplementedError


class MLE(Loss):
    """Negative log marginal likelihood and its hyper-parameter gradient (loss.py:31-128).

    Takes and returns numpy (scipy drives the optimisers on the host, loss.py:37,57,90,128).  One
    factorisation per call, like the reference, but the gradient never forms dK: K^-1 is assembled tile
    by tile from L^-T and reduced against K_se generated on chip (N^3 flops instead of (1/3+2P) N^3).

    Memory: every evaluation needs one [B, npad, npad] factor buffer.  The instance keeps its buffer between calls (an
    optimiser makes hundreds of evaluations on the same shapes); when a fresh buffer would not fit beside the model's own fitted
    factor (config 4 on one GPU: 137 GB each), the model's buffer is borrowed instead and the model marked stale — its next
    predict refits, exactly as after set_params (gpr.py:28-31).  Otherwise the model's fitted state is not touched.
    """

    _L_buf = None

    def _factor_buffer(self, B: int, npad: int, dev) -> "torch.Tensor | None":
        """A [B, npad, npad] buffer to factorise into: this instance's own from the previous call, else the model's (borrowed)
        when there is no room for another one, else None (engine.fit allocates; the result is kept for the next call)."""
        shape = (B, npad, npad)
        buf = self._L_buf
        if buf is not None and tuple(buf.shape[1:]) == shape[1:] and buf.shape[0] >= B and buf.device == dev:
            return buf[:B]                                   # a smaller batch (the last chunk of GRBCM_MLE) uses the front
        self._L_buf = buf = None
        need = B * npad * npad * 8
        free, _ = torch.cuda.mem_get_info(dev)
        st = getattr(self.model, "_state", None)
        if need + (2 << 30) > free and isinstance(st, engine.FitState) and tuple(st.L.shape) == shape and st.L.device == dev:
            self.model._state = None                         # the model's factor is overwritten: it refits on its next use
            self.model.need_upd = True
            self._L_buf = st.L
            return st.L
        return None

    def _fit(self, params, want_alpha: bool, check: bool = True):
        """One factorisation at `params` (numpy like the reference's API, or a device tensor: the batched training loop keeps its
        iterates in HBM).  check=False leaves `info` on the device (no host synchronisation; the caller masks failed experts)."""
        xs, ys = self.model.x, self.model.y
        spec = fused_spec(self.model.cov, xs.shape[-1])
        x = engine._dev(xs).reshape(-1, xs.shape[-2], xs.shape[-1])
        y = engine._dev(ys).reshape(-1, ys.shape[-1])
        if isinstance(params, torch.Tensor):
            hp = engine._dev(params.detach()).to(torch.float64).contiguous()
        else:
            hp = engine._dev(torch.from_numpy(np.ascontiguousarray(params, dtype=np.float64)))
        self._batched = hp.dim() > 1 or xs.dim() > 2
        hp = hp.reshape(-1, hp.shape[-1])
        B = max(hp.shape[0], x.shape[0])
        if hp.shape[0] != B:
            hp = hp.expand(B, -1).contiguous()
        if x.shape[0] != B:
            x = x.expand(B, -1, -1).contiguous()
            y = y.expand(B, -1).contiguous()
        assert hp.shape[-1] == self.model.cov.get_params_shape(xs)[-1]
        if spec is None:
            # any other Covar composition: materialised K (and dK for the gradient, like the reference, loss.py:94)
            n = x.shape[1]
            if want_alpha:
                K, dK = self.model.cov.kernel_and_grad(hp, x)
                self._dK = dK.reshape(B, -1, n, n)
            else:
                K = self.model.cov.kernel(hp, x)
            K = K.reshape(B, n, n).clone()
            K.diagonal(dim1=-2, dim2=-1).add_(engine.JITTER)           # loss.py:38,63,96
            st = engine.fit_dense(K, y, want_alpha=want_alpha)
            if check:
                st.check()
            return st
        shared = getattr(self, "shared_state", None)        # set by GRBCM_MLE: the communication expert fitted at these params
        if shared is not None and not engine.can_share(shared, x.shape[1], x.shape[2], spec[0]):
            shared = None
        npad = engine.round_up(x.shape[1], engine.TILE)
        st = engine.fit(x, y, hp, spec[0], want_alpha=want_alpha, shared=shared, L_out=self._factor_buffer(B, npad, x.device))
        if self._L_buf is None or st.L.numel() > self._L_buf.numel():
            self._L_buf = st.L
        if check:
            st.check()                                       # tc.cholesky raises on a non-PD matrix, loss.py:39
        return st

    def _grad(self, st) -> torch.Tensor:
        if isinstance(st, engine.DenseState):
            return engine.lml_grad_dense(st, self._dK)
        return engine.lml_grad(st)

    def _out(self, t: torch.Tensor) -> ndarray:
        t = t if self._batched else t.squeeze(0)
        return t.cpu().numpy()

    def loss(self, params: ndarray) -> float:
        st = self._fit(params, want_alpha=False)
        self.loss_value = self._out(st.nlml())               # loss.py:49-57
        return self.loss_value

    def grad(self, params: ndarray) -> ndarray:
        st = self._fit(params, want_alpha=True)
        self.grad_value = self._out(self._grad(st))          # loss.py:77-90
        return self.grad_value

    def loss_and_grad(self, params: ndarray) -> Tuple[float, ndarray]:
        st = self._fit(params, want_alpha=True)
        self.loss_value = self._out(st.nlml())
        self.grad_value = self._out(self._grad(st))          # loss.py:107-128
        return (self.loss_value, self.grad_value)
