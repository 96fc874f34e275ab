"""Covariance kernels with the reference's API (PyGPR/covar.py), computed by the sm_100a library.

Same classes, signatures, hyper-parameter layout and shape conventions as the reference:
``hp = [sig, ls_1..ls_d]`` for ``Squared_exponential`` (amplitude and noise enter SQUARED, the
length-scales MULTIPLY the inputs, no 1/2 factor — covar.py:86,137-149), ``[sig_n]`` for
``White_noise``, concatenation along the last axis for ``Compose``; a leading batch axis is the
expert axis and a batch of one is squeezed away (covar.py:162-166).

Tensors may live on the CPU (like the reference) or on the GPU; results are returned on the device
of ``x``.  The arithmetic always runs on the GPU — there is no CPU implementation here.
"""
from __future__ import annotations

from typing import List, Optional, Protocol, Sequence

import torch
from torch import Tensor

from . import engine

# the reference flips the process-wide default dtype to float64 at import (covar.py:5); drop-in users rely on it
torch.set_default_dtype(torch.float64)


class Covar(Protocol):
    """Structural type of a covariance kernel (covar.py:9-25)."""

    def get_params_shape(self, x: Tensor) -> List[int]: ...

    def init_params(self, x: Tensor) -> Tensor: ...

    def kernel(self, params: Tensor, x: Tensor, xp: Tensor = None) -> Tensor: ...

    def kernel_and_grad(self, params: Tensor, x: Tensor) -> List[Tensor]: ...


def _b3(x: Tensor) -> Tensor:
    return engine._dev(x).reshape(-1, x.shape[-2], x.shape[-1])


def _b2(hp: Tensor) -> Tensor:
    return engine._dev(hp).reshape(-1, hp.shape[-1])


def _expand(t: Tensor, B: int) -> Tensor:
    return t if t.shape[0] == B else t.expand(B, *t.shape[1:]).contiguous()


def _back(t: Tensor, like: Tensor) -> Tensor:
    return t if like.is_cuda else t.cpu()


def _params_shape(x: Tensor, width: int) -> List[int]:
    shape = list(x.shape)
    shape[-1] = width
    shape.pop(-2)
    return shape


class Squared_exponential:
    """K(x,x') = sig^2 exp(-|(x-x') * ls|^2)   (covar.py:84-206)."""

    def get_params_shape(self, x: Tensor) -> List[int]:
        return _params_shape(x, x.shape[-1] + 1)            # covar.py:89-94

    def init_params(self, x: Tensor) -> Tensor:
        return torch.ones(self.get_params_shape(x), dtype=torch.float64, device=x.device)  # covar.py:96-100

    def kernel(self, hp: Tensor, x: Tensor, xp: Optional[Tensor] = None) -> Tensor:
        assert hp.shape[-1] == self.get_params_shape(x)[-1]  # covar.py:131
        hpd, xd = _b2(hp), _b3(x)
        if xp is None:
            B = max(hpd.shape[0], xd.shape[0])
            K = engine.se_kernel_matrix(_expand(hpd, B), _expand(xd, B))
        else:
            xpd = _b3(xp)
            B = max(hpd.shape[0], xd.shape[0], xpd.shape[0])
            K = engine.se_cross_matrix(_expand(hpd, B), xd, xpd)
        return _back(K.squeeze(0), x)                        # covar.py:162-166

    def kernel_and_grad(self, hp: Tensor, x: Tensor) -> List[Tensor]:
        assert hp.shape[-1] == self.get_params_shape(x)[-1]  # covar.py:171
        hpd, xd = _b2(hp), _b3(x)
        B = max(hpd.shape[0], xd.shape[0])
        hpd, xd = _expand(hpd, B), _expand(xd, B)
        K = engine.se_kernel_matrix(hpd, xd)
        dK = engine.se_kernel_grad(hpd, xd, K)
        return [_back(K.squeeze(0), x), _back(dK.squeeze(0), x)]


class White_noise:
    """K = sig_n^2 I; cross-covariance is the scalar 0 (covar.py:209-269)."""

    def get_params_shape(self, x: Tensor) -> List[int]:
        return _params_shape(x, 1)                           # covar.py:214-219

    def init_params(self, x: Tensor) -> Tensor:
        return 1e-4 * torch.ones(self.get_params_shape(x), dtype=torch.float64, device=x.device)  # covar.py:221-225

    def kernel(self, hp: Tensor, x: Tensor, xp: Optional[Tensor] = None) -> Tensor:
        if xp is not None:
            return torch.tensor(0)                           # covar.py:242-243 (0-d int64, as the reference)
        hpd = _b2(hp)
        n = x.shape[-2]
        K = torch.eye(n, dtype=torch.float64, device=hpd.device)[None] * hpd[:, 0, None, None].square()  # covar.py:237-238
        return _back(K.squeeze(0), x)

    def kernel_and_grad(self, hp: Tensor, x: Tensor) -> List[Tensor]:
        hpd = _b2(hp)
        n = x.shape[-2]
        eye = torch.eye(n, dtype=torch.float64, device=hpd.device)[None]
        K = eye * hpd[:, 0, None, None].square()
        dK = (eye * (2.0 * hpd[:, 0, None, None])).unsqueeze(1)  # covar.py:262-264: K * 2/sig_n
        return [_back(K.squeeze(0), x), _back(dK.squeeze(0), x)]


class Compose:
    """Sum of kernels; hyper-parameters concatenated along the last axis (covar.py:28-81)."""

    def __init__(self, covars: Sequence[Covar]) -> None:
        self.covars = covars

    def get_params_shape(self, x: Tensor) -> List[int]:
        nparams = sum(c.get_params_shape(x)[-1] for c in self.covars)
        return _params_shape(x, nparams)                     # covar.py:38-43

    def init_params(self, x: Tensor) -> Tensor:
        return torch.cat([c.init_params(x) for c in self.covars], dim=-1)  # covar.py:45-48

    def _split(self, hp: Tensor, x: Tensor):
        assert hp.shape[-1] == self.get_params_shape(x)[-1]  # covar.py:52,66
        chunks = [c.get_params_shape(x)[-1] for c in self.covars]
        return hp.split(chunks, dim=-1)                      # covar.py:54-55

    def kernel(self, hp: Tensor, x: Tensor, xp: Optional[Tensor] = None) -> Tensor:
        params = self._split(hp, x)
        krn = self.covars[0].kernel(params[0], x, xp)
        for cov, p in zip(self.covars[1:], params[1:]):
            krn = krn + cov.kernel(p, x, xp).to(krn.device)  # covar.py:59-60
        return krn

    def kernel_and_grad(self, hp: Tensor, x: Tensor) -> List[Tensor]:
        params = self._split(hp, x)
        krn, dk = None, []
        for cov, p in zip(self.covars, params):
            k_i, dk_i = cov.kernel_and_grad(p, x)            # (the reference evaluates later children twice, covar.py:76-77)
            krn = k_i if krn is None else krn + k_i
            dk.append(dk_i)
        return [krn, torch.cat(dk, dim=-3)]                  # covar.py:79

    def fused_spec(self):
        """(has_noise,) if this composition is the one the fused CUDA path handles:
        Compose([Squared_exponential()]) or Compose([Squared_exponential(), White_noise()]); else None."""
        kinds = [type(c) for c in self.covars]
        if kinds == [Squared_exponential]:
            return (False,)
        if kinds == [Squared_exponential, White_noise]:
            return (True,)
        return None


def fused_spec(cov, d: Optional[int] = None):
    """has_noise flag tuple for covariances the fused fit/predict/LML kernels support, else None (-> dense path).
    d: input dimension, when known — beyond what the on-chip tile generation holds the answer is None as well."""
    if d is not None and not engine.fused_dim_ok(d):
        return None
    if isinstance(cov, Squared_exponential):
        return (False,)
    if isinstance(cov, Compose):
        return cov.fused_spec()
    return None
