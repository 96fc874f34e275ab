"""Small end-to-end case for compute-sanitizer (memcheck / racecheck / synccheck): every kernel family once — general and
shared-block GRBCM fit + predict (dataflow and per-column Cholesky, both backward solves, GEMM-form communication expert),
NLML gradient (general and shared), full covariance, kernel builders."""
import math, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import pygpr_b200 as P
from pygpr_b200 import engine as eng

torch.manual_seed(0)
d = 5
cov = P.Compose([P.Squared_exponential(), P.White_noise()])
hp1 = torch.tensor([1.0] + [2.0 / math.sqrt(d)] * d + [0.3], dtype=torch.float64)
out = {}
for nc, nls, ng, tag in ((7, 150, 128, "shared"), (3, 100, 60, "general")):
    xl = torch.rand(nc, nls, d, dtype=torch.float64); yl = torch.sin(xl.sum(-1))
    xg = torch.rand(ng, d, dtype=torch.float64); yg = torch.sin(xg.sum(-1))
    xs = torch.rand(70, d, dtype=torch.float64)
    model = P.GRBCM(xl, yl, xg, yg, cov)
    model.gpl.set_params(hp1.repeat(nc, 1)); model.gpg.set_params(hp1)
    mu, var = model.predict(xs, var="diag")
    out[tag] = (float(mu.sum()), float(var.sum()), model.gpl._state.shared_from is not None)
    loss = P.GRBCM_MLE(model, chunk=4)
    f, g = loss.loss_and_grad(hp1.numpy())
    out[tag + "_mle"] = (float(f), float(np.abs(g).sum()))
    muf, covf = model.predict(xs[:20], var="full")
    out[tag + "_full"] = float(covf.diagonal().sum())
gp = P.Exact_GP(torch.rand(300, d, dtype=torch.float64), torch.rand(300, dtype=torch.float64), cov)
gp.set_params(hp1)
ys, cs = gp.predict(torch.rand(33, d, dtype=torch.float64), var="diag")
l, j = P.MLE(gp).loss_and_grad(hp1.numpy())
K, dK = cov.kernel_and_grad(hp1, gp.x[:40])
torch.cuda.synchronize()
print("sanitize_case ok", out, float(ys.sum()), float(l), float(dK.sum()))
