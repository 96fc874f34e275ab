"""pygpr_b200 — B200-native (sm_100a) drop-in for PyGPR's data-parallel hot path.

Re-exports the reference's public names (PyGPR/__init__.py:1-7).  Importing sets torch's default dtype to
float64, as the reference does (covar.py:5).  The CUDA library is loaded on first use and there is no CPU
fallback: without ``libpygpr_b200.so`` or without a GPU the compute calls raise.
"""
from .gpr import GPR, Exact_GP
from .covar import Squared_exponential, Covar, Compose, White_noise
from .loss import Loss, MLE
from .opt import Opt, CG, Nelder_Mead, BFGS_Quad, CG_Quad, hessian
from .gr_bcm import GRBCM, GRBCM_MLE, log_likelihood_batched
from .sampler import UNIFORM, MATERN1, sample_gp, cluster_samples, euclidean_dist
from .hp_update import get_learn_rate
from .scikit_model import SK_WRAP

__all__ = [
    "GPR", "Exact_GP", "Squared_exponential", "Covar", "Compose", "White_noise", "Loss", "MLE", "Opt", "CG",
    "Nelder_Mead", "BFGS_Quad", "CG_Quad", "hessian", "GRBCM", "log_likelihood_batched", "get_learn_rate", "SK_WRAP",
    "UNIFORM", "MATERN1", "sample_gp", "cluster_samples", "euclidean_dist", "GRBCM_MLE",
]
