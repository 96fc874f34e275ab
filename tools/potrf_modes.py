"""A/B of the two batched-Cholesky schedules (PYGPR_POTRF_MODE=columns|dataflow): results and device time."""
import math, os, sys
import torch
sys.path.insert(0, ".")
from pygpr_b200 import engine as eng
from pygpr_b200 import config


def data(B, n, d, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, n, d, generator=g, dtype=torch.float64).cuda()
    y = (torch.sin(x.sum(-1)) + 0.1 * torch.randn(B, n, dtype=torch.float64, device="cuda"))
    hp = torch.empty(B, d + 2, dtype=torch.float64)
    hp[:, 0] = 1.0; hp[:, 1:d + 1] = 2.0 / math.sqrt(d); hp[:, d + 1] = 0.3
    return x, y, hp.cuda()


def fit(mode, x, y, hp, L=None):
    config.set_potrf_mode(mode)
    return eng.fit(x, y, hp, True, L_out=L)


def check(B, n, d):
    x, y, hp = data(B, n, d)
    a = fit("columns", x, y, hp); b = fit("dataflow", x, y, hp)
    torch.cuda.synchronize()
    nt = a.npad // 128
    tril = torch.tril(torch.ones(a.npad, a.npad, dtype=torch.bool, device="cuda"))
    dL = float(((a.L - b.L) * tril).abs().max())
    out = dict(B=B, n=n, dL=dL, dz=float((a.z - b.z).abs().max()), dalpha=float((a.alpha - b.alpha).abs().max()),
               dlogdet=float((a.logdet_half - b.logdet_half).abs().max()), dzz=float((a.zz - b.zz).abs().max()),
               dinv=float((a.invdiag - b.invdiag).abs().max()), info=(int(a.info.abs().sum()), int(b.info.abs().sum())))
    print(out, flush=True)
    # in-place (gen=0) path
    K = eng.se_kernel_matrix(hp, x)
    buf = torch.zeros(B, a.npad, a.npad, dtype=torch.float64, device="cuda")
    buf[:, :n, :n] = K + torch.diag_embed((hp[:, d + 1] ** 2 + 1e-7)[:, None].expand(B, n))
    idx = torch.arange(n, a.npad, device="cuda"); buf[:, idx, idx] = 1.0
    b1, b2 = buf.clone(), buf.clone()
    config.set_potrf_mode("columns"); r1 = eng.potrf_inplace(b1, None)
    config.set_potrf_mode("dataflow"); r2 = eng.potrf_inplace(b2, None)
    torch.cuda.synchronize()
    print("   inplace dL", float(((b1 - b2) * tril).abs().max()), "vs gen", float(((b1 - a.L) * tril).abs().max()), flush=True)


def bench(B, n, d, reps=3):
    x, y, hp = data(B, n, d)
    npad = eng.round_up(n, 128)
    L = torch.empty(B, npad, npad, dtype=torch.float64, device="cuda")
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    for mode in ("columns1", "columns", "dataflow"):
        fit(mode, x, y, hp, L); torch.cuda.synchronize()
        best = 1e9
        for _ in range(reps):
            e0.record(); st = fit(mode, x, y, hp, L); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print(f"fit {mode:9s} B={B} n={n}: {best:9.2f} ms {B * n ** 3 / 3.0 / best * 1e-9:6.2f} TFLOP/s info={int(st.info.abs().sum())}", flush=True)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("all", "check"):
        for c in [(3, 300, 4), (5, 1500, 8), (1, 128, 3), (2, 129, 16), (40, 1024, 16)]:
            check(*c)
    if what == "small":
        for c in [(1, 16384, 8), (1, 4096, 16), (1, 8192, 16), (2, 8192, 8), (4, 8192, 16), (1, 1024, 4), (3, 2048, 8)]:
            bench(*c)
    if what == "border":
        for c in [(2, 16384, 8), (4, 8192, 16), (4, 16384, 8), (16, 2048, 8), (8, 4096, 8), (16, 4096, 8), (148, 1024, 8), (2, 8192, 8)]:
            bench(*c)
    if what in ("all", "bench"):
        for c in [(63, 4096, 8), (32, 8192, 16), (255, 8192, 16), (8, 16384, 8), (1, 16384, 8)]:
            bench(*c)
