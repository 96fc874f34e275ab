"""Quadratic learn-rate rule with the reference's API (PyGPR/hp_update.py)."""
from torch import Tensor

from .loss import Loss


def get_learn_rate(current_param: Tensor, loss_new: Loss, eps: float) -> float:
    """Step size along -J from a parabola through the loss at theta and theta -/+ eps J (hp_update.py:6-28).

    One ``loss_and_grad`` and two ``loss`` evaluations; batched parameters give one gamma per expert.
    """
    theta = current_param.detach().cpu().clone().numpy()
    f0, J = loss_new.loss_and_grad(theta)
    f_minus = loss_new.loss(theta - eps * J)                 # hp_update.py:16
    f_plus = loss_new.loss(theta + eps * J)                  # hp_update.py:17
    slope = (f_minus - f_plus) / (2.0 * eps)                 # C1, hp_update.py:19
    curv = (f_minus + f_plus - 2 * f0) / (2.0 * eps ** 2)    # C2, hp_update.py:20
    return -0.5 * (slope / curv)                             # hp_update.py:22
