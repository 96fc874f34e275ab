"""Race / determinism soak: the same fit + predict + gradient repeated many times must reproduce bit for bit (every reduction in
the library has a fixed order; a data race between CTAs, a missed wait on a progress counter or a stale L1 line would show
up as run-to-run differences)."""
import math, os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import pygpr_b200 as P
from pygpr_b200 import engine as eng
from pygpr_b200 import config

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 60
bad = 0
for (nc, nls, ng, d, m) in [(40, 640, 384, 6, 300), (9, 1500, 256, 16, 129), (3, 3000, 1100, 8, 64), (64, 256, 128, 4, 1000)]:
    g = torch.Generator().manual_seed(nc)
    xl = torch.rand(nc, nls, d, generator=g, dtype=torch.float64)
    yl = torch.sin(xl.sum(-1)) + 0.1 * torch.randn(nc, nls, generator=g, dtype=torch.float64)
    xg = torch.rand(ng, d, generator=g, dtype=torch.float64)
    yg = torch.sin(xg.sum(-1)) + 0.1 * torch.randn(ng, generator=g, dtype=torch.float64)
    xs = torch.rand(m, d, generator=g, dtype=torch.float64)
    hp = torch.empty(nc, d + 2, dtype=torch.float64)
    hp[:, 0] = 1.0; hp[:, 1:d + 1] = 2.0 / math.sqrt(d); hp[:, d + 1] = 0.3
    cov = P.Compose([P.Squared_exponential(), P.White_noise()])
    for share in ("1", "0"):
        config.options.share = share != "0"
        ref = None
        for r in range(reps):
            model = P.GRBCM(xl, yl, xg, yg, cov)
            model.gpl.set_params(hp); model.gpg.set_params(hp[0])
            mu, var = model.predict(xs, var="diag")
            grad = eng.lml_grad(model.gpl._state)
            cur = (mu.clone(), var.clone(), grad.cpu(), model.gpl._state.alpha.cpu(), model.gpl._state.logdet_half.cpu())
            if ref is None:
                ref = cur
            elif not all(torch.equal(a, b) for a, b in zip(ref, cur)):
                bad += 1
                print("MISMATCH case nc=%d share=%s rep=%d: %s" % (nc, share, r, [float((a - b).abs().max()) for a, b in zip(ref, cur)]), flush=True)
        print("nc=%d n=%d share=%s: %d repetitions %s" % (nc, ng + nls, share, reps, "bit-identical" if not bad else "DIFFER"), flush=True)
print("determinism_check:", "ok" if not bad else "FAILED (%d)" % bad)
sys.exit(1 if bad else 0)
