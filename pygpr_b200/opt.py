"""Host-side optimisers with the reference's API (PyGPR/opt.py).

These are scalar drivers over ``[nhp]`` numpy vectors; every function/gradient evaluation is one pass
through the GPU hot path via ``Loss``.  Behaviour kept from the reference: scipy CG / Nelder-Mead with
the same default options, a trace written to ``./opt.dat``, ``model.set_params`` with the result.
"""
from __future__ import annotations

from typing import Callable

import numpy as np
import scipy.optimize as scopt
import torch
from numpy import ndarray

from .loss import Loss


class Opt:
    """Base class for optimisers (opt.py:11-26)."""

    def __init__(self, loss: Loss, par: ndarray = None) -> None:
        self.loss: Loss = loss
        self.args: dict = {}
        self.x: ndarray = NotImplemented

    def minimize(self):
        raise NotImplementedError

    def step(self):
        raise NotImplementedError


def _start_point(loss: Loss) -> ndarray:
    return loss.model.params.detach().cpu().clone().numpy()


class _ScipyOpt(Opt):
    method: str = ""
    use_jac: bool = False

    def __init__(self, loss: Loss) -> None:
        super().__init__(loss)
        self.res: scopt.OptimizeResult = NotImplemented

    def minimize(self) -> None:
        x0 = _start_point(self.loss)
        with open("opt.dat", "w") as self.fstr:              # opt.py:48,101
            fun = self.loss.loss_and_grad if self.use_jac else self.loss.loss
            self.res = scopt.minimize(fun, x0, method=self.method, jac=True if self.use_jac else None,
                                      callback=self.callback, options=self.args)
        return self._finish()

    def step(self):
        raise NotImplementedError


class CG(_ScipyOpt):
    """Nonlinear conjugate gradient through scipy (opt.py:29-78)."""
    method = "CG"
    use_jac = True

    def __init__(self, loss: Loss) -> None:
        super().__init__(loss)
        self.args = {"gtol": 1e-4, "maxiter": 1000, "disp": True, "return_all": True}  # opt.py:37-42

    def _finish(self) -> None:
        self.loss.model.set_params(torch.from_numpy(self.res.x))  # set in both outcomes, opt.py:61-65
        if self.res.success is not True:
            print("Optimizer Failed")

    def callback(self, params: ndarray) -> None:
        print(*params, self.loss.loss_value, np.linalg.norm(self.loss.grad_value), file=self.fstr)


class Nelder_Mead(_ScipyOpt):
    """Derivative-free simplex search through scipy (opt.py:81-122)."""
    method = "Nelder-Mead"
    use_jac = False

    def __init__(self, loss: Loss) -> None:
        super().__init__(loss)
        self.args = {"fatol": 1e-4, "maxiter": 1000, "disp": True, "return_all": True}  # opt.py:89-94

    def _finish(self) -> None:
        if self.res.success is True:                          # only on success, opt.py:111-114
            self.loss.model.set_params(torch.from_numpy(self.res.x))
        else:
            print("Optimizer Failed")

    def callback(self, params: ndarray) -> None:
        print(*params, self.loss.loss_value, file=self.fstr)


def hessian(x: np.ndarray, jac: Callable[..., np.ndarray], eps: float) -> np.ndarray:
    """Forward-difference Hessian, one gradient pair per column (opt.py:125-136)."""
    dim = x.shape[-1]
    hess = np.empty([dim, dim])
    for i in range(dim):
        shifted = np.copy(x)
        shifted[i] += eps
        hess[:, i] = (jac(shifted) - jac(x)) / eps
    return hess


class _QuadOpt(Opt):
    def __init__(self, loss: Loss, gtol: float = 1e-4, max_iter: int = 100, fd_eps: float = 1e-5) -> None:
        super().__init__(loss)
        self.x: ndarray = NotImplemented
        self.r: ndarray = NotImplemented
        self.eps = fd_eps
        self.max_iter = max_iter
        self.gtol = gtol

    def _iterate(self) -> int:
        k = 0
        with open("opt.dat", "w") as fstr:                    # opt.py:196,282
            while np.linalg.norm(self.r) > self.gtol and k < self.max_iter:
                self.step()
                k += 1
                print(k, np.linalg.norm(self.r), file=fstr)
        if self.loss.model is not None:
            self.loss.model.set_params(torch.tensor(self.x))
        return k


class CG_Quad(_QuadOpt):
    """Linear conjugate gradient on the local quadratic model; Hessian-vector products by forward
    differences of the gradient, i.e. two ``loss.grad`` calls per step (opt.py:139-208)."""

    def __init__(self, loss: Loss, gtol: float = 1e-4, max_iter: int = 100, fd_eps: float = 1e-5) -> None:
        super().__init__(loss, gtol, max_iter, fd_eps)
        self.p: ndarray = NotImplemented

    def hessian_product(self, par: ndarray, v: ndarray, eps: float) -> ndarray:
        return (self.loss.grad(par + eps * v) - self.loss.grad(par)) / eps  # opt.py:161-163

    def step(self) -> None:
        r, p, x = np.copy(self.r), np.copy(self.p), np.copy(self.x)
        Hp = self.hessian_product(x, p, eps=self.eps)
        rr = np.dot(r, r)
        alp = rr / np.dot(p, Hp)
        x = x + alp * p
        r = r + alp * Hp
        p = (np.dot(r, r) / rr) * p - r
        self.r, self.p, self.x = r, p, x

    def minimize(self, par: ndarray = None) -> int:
        self.x = _start_point(self.loss) if par is None else par
        self.r = self.loss.grad(self.x)
        self.p = -1.0 * self.r
        return self._iterate()


class BFGS_Quad(_QuadOpt):
    """Full-step BFGS with the inverse-Hessian update; one ``loss.grad`` per step (opt.py:211-295)."""

    def __init__(self, loss: Loss, gtol: float = 1e-4, max_iter: int = 100, fd_eps: float = 1e-5) -> None:
        super().__init__(loss, gtol, max_iter, fd_eps)
        self.HI: ndarray = NotImplemented

    def hessian_inv_update(self, HI: ndarray, s: ndarray, y: ndarray) -> ndarray:
        eye = np.identity(HI.shape[-1])
        rho = 1 / np.dot(y, s)
        left = eye - rho * np.outer(s, y)
        right = eye - rho * np.outer(y, s)
        return left @ HI @ right + rho * np.outer(s, s)       # opt.py:233-246

    def step(self):
        x_old, r_old = np.copy(self.x), np.copy(self.r)
        x = x_old - self.HI @ r_old                            # full quasi-Newton step, opt.py:256-258
        r = self.loss.grad(x)
        self.HI = self.hessian_inv_update(np.copy(self.HI), x - x_old, r - r_old)
        self.x, self.r = x, r

    def minimize(self, par: ndarray = None, H0: ndarray = None):
        self.x = _start_point(self.loss) if par is None else par
        self.r = self.loss.grad(self.x)
        self.HI = np.identity(self.x.shape[-1]) if H0 is None else np.linalg.inv(H0)
        return self._iterate()
