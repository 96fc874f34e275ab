"""One fit + predict + LML-gradient pass at a size small enough for ncu replay (used for profiles/)."""
import math, sys
import torch
sys.path.insert(0, ".")
from pygpr_b200 import engine as eng

def main():
    B, n, d, m = (int(v) for v in (sys.argv[1:5] if len(sys.argv) >= 5 else (37, 4096, 16, 512)))
    g = torch.Generator().manual_seed(0)
    x = torch.rand(B, n, d, generator=g, dtype=torch.float64).cuda()
    y = torch.sin(x.sum(-1)) + 0.1 * torch.randn(B, n, dtype=torch.float64, device="cuda")
    xs = torch.rand(m, d, generator=g, dtype=torch.float64).cuda()
    hp = torch.empty(B, d + 2, dtype=torch.float64)
    hp[:, 0] = 1.0; hp[:, 1:d + 1] = 2.0 / math.sqrt(d); hp[:, d + 1] = 0.3
    hp = hp.cuda()
    for it in range(2):
        st = eng.fit(x, y, hp, True)
        mean, var = eng.predict(st, xs)
        grad = eng.lml_grad(st)
        torch.cuda.synchronize()
    print("ok", float(mean.sum()), float(var.sum()), float(grad.sum()), int(st.info.abs().sum()))

if __name__ == "__main__":
    main()
