"""Random SPD matrices through gpr_potrf_batched (in place, gen = 0) + gpr_potri_batched against torch.linalg on the same GPU:
every schedule (dataflow / three-stream per-column / single-stream), ragged orders, 1..12 matrices."""
import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from pygpr_b200 import engine as eng
from pygpr_b200 import config

g = torch.Generator(device="cuda").manual_seed(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
worst = 0.0
for case in range(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    B = int(torch.randint(1, 13, (1,)).item())
    n = int(torch.randint(1, 2300, (1,)).item())
    A = torch.randn(B, n, n + 8, generator=g, dtype=torch.float64, device="cuda")
    K = A @ A.transpose(1, 2) / (n + 8) + 0.5 * torch.eye(n, dtype=torch.float64, device="cuda")
    Lref = torch.linalg.cholesky(K)
    for mode in ("auto", "columns", "columns1", "dataflow"):
        config.set_potrf_mode(mode)
        npad = eng.round_up(n, 128)
        buf = torch.zeros(B, npad, npad, dtype=torch.float64, device="cuda")
        buf[:, :n, :n] = K
        idx = torch.arange(n, npad, device="cuda"); buf[:, idx, idx] = 1.0
        invdiag, _, ldh, _, info = eng.potrf_inplace(buf)
        assert int(info.abs().sum()) == 0, (case, mode, B, n)
        e = float((torch.tril(buf[:, :n, :n]) - Lref).abs().max() / Lref.abs().max())
        e2 = float((ldh - Lref.diagonal(dim1=-2, dim2=-1).log().sum(-1)).abs().max() / max(1.0, float(ldh.abs().max())))
        worst = max(worst, e, e2)
        assert e < 1e-10 and e2 < 1e-12, (case, mode, B, n, e, e2)
    config.set_potrf_mode("auto")
    Kinv = eng.spd_inverse(K)
    e3 = float((Kinv @ K - torch.eye(n, dtype=torch.float64, device="cuda")).abs().max())
    worst = max(worst, e3 * 1e-3)
    assert e3 < 1e-8, (case, B, n, e3)
print("fuzz_potrf: ok, worst rel err %.2e" % worst)
