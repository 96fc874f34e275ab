"""Round-2 evidence case for ncu (one launch of each kernel that had no counters in profiles/ after round 1):
shared-block fit (32 experts x 8192 on a 4096 communication block, the 8-GPU share of config 4) -> trtri_kernel and the
shared-block lauum_grad_kernel; the stand-alone se_tile_kernel; the one-CTA-per-expert trsv_backward_kernel (160 experts x
2048); grbcm_partial_kernel at 255 x 4096; plus the emulated-FP64 pieces (split_kernel, gemm_kernel<6>, vt_reduce_kernel).

    python tools/profile_case2.py
"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pygpr_b200 as P  # noqa: E402
from pygpr_b200 import config, engine as eng  # noqa: E402


def hp_of(B, d):
    hp = torch.empty(B, d + 2, dtype=torch.float64)
    hp[:, 0] = 1.0; hp[:, 1:d + 1] = 2.0 / math.sqrt(d); hp[:, d + 1] = 0.3
    return hp


def main():
    g = torch.Generator().manual_seed(0)
    nc, m0, d, m = 32, 4096, 16, 1024
    x = torch.rand(nc + 1, m0, d, generator=g, dtype=torch.float64)
    y = torch.sin(x.sum(-1)) + 0.1 * torch.randn(nc + 1, m0, generator=g, dtype=torch.float64)
    xs = torch.rand(m, d, generator=g, dtype=torch.float64)
    cov = P.Compose([P.Squared_exponential(), P.White_noise()])
    model = P.GRBCM(x[1:].cuda(), y[1:].cuda(), x[0].cuda(), y[0].cuda(), cov)
    model.gpl.set_params(hp_of(nc, d)); model.gpg.set_params(hp_of(1, d)[0])
    mu, var = model.predict(xs.cuda(), var="diag")                 # gpr_potrf_grbcm, predict_kernel<shared>, grbcm_partial/finalize
    grad = eng.lml_grad(model.gpl._state)                         # trtri_kernel + lauum_grad_kernel (shared block) + grad_finalize
    mu_e = mu
    if "noemu" not in sys.argv:
        config.options.emulate_slices = 6
        mu_e, var_e = model.predict(xs.cuda(), var="diag")         # split_kernel, gemm_kernel<6>, vt_reduce_kernel, tril_rows_from_upper
        config.options.emulate_slices = 0
    K = eng.se_kernel_matrix(hp_of(4, d)[:, :d + 1].contiguous().cuda(), x[:4].cuda())   # se_tile_kernel, 4 x 4096^2
    # one CTA per expert backward solve: experts >= SMs
    B2, n2 = 160, 2048
    x2 = torch.rand(B2, n2, 8, generator=g, dtype=torch.float64).cuda()
    y2 = torch.sin(x2.sum(-1))
    st2 = eng.fit(x2, y2, hp_of(B2, 8).cuda(), True)               # trsv_backward_kernel (single-CTA form)
    ml = torch.rand(255, 4096, generator=g, dtype=torch.float64).cuda()
    vl = 0.5 + torch.rand(255, 4096, generator=g, dtype=torch.float64).cuda()
    vg = 1.0 + torch.rand(4096, generator=g, dtype=torch.float64).cuda()
    S, _, _ = eng.grbcm_partial(ml, vl, vg, 0)                    # grbcm_partial_kernel at the bench's shape
    torch.cuda.synchronize()
    print("ok", float(mu.sum()), float((mu - mu_e).abs().max()), float(grad.sum()), float(K.sum()), float(st2.alpha.sum()), float(S.sum()))


if __name__ == "__main__":
    main()
