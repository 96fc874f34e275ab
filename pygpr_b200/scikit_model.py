"""scikit-learn estimator shell with the reference's API (PyGPR/scikit_model.py)."""
from typing import Any

from sklearn.base import BaseEstimator, RegressorMixin
from torch import Tensor

from .gpr import GPR


class SK_WRAP(RegressorMixin, BaseEstimator):
    """Wraps a GPR model as an sklearn regressor (scikit_model.py:15-35).

    Deliberate fix: ``fit`` marks the wrapped model dirty.  The reference sets ``need_upd`` on the wrapper
    instead of the model (scikit_model.py:33), so a second ``fit`` with new data silently keeps the old
    factorisation; first-fit behaviour is identical.
    """

    def __init__(self, model: GPR) -> None:
        self.model: GPR = model

    def fit(self, x: Tensor, y: Tensor) -> Any:
        print("Fitting", x.shape, y.shape)
        self.model.x = x
        self.model.y = y
        self.model.need_upd = True
        return self

    def predict(self, xp: Tensor) -> Tensor:
        print("Predicting", xp.shape)
        yp, _ = self.model.predict(xp, var="none")           # mean only, scikit_model.py:34
        return yp
