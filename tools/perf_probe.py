"""Quick device-timed probe of fit (batched Cholesky) and predict throughput. Not the bench."""
import math, sys, time
import torch
sys.path.insert(0, ".")
from pygpr_b200 import engine as eng

def run(B, n, d, m, reps=2):
    g = torch.Generator().manual_seed(0)
    x = torch.rand(B, n, d, generator=g, dtype=torch.float64).cuda()
    y = (torch.sin(x.sum(-1)) + 0.1 * torch.randn(B, n, dtype=torch.float64, device="cuda"))
    xs = torch.rand(m, d, generator=g, dtype=torch.float64).cuda()
    hp = torch.empty(B, d + 2, dtype=torch.float64)
    hp[:, 0] = 1.0; hp[:, 1:d + 1] = 2.0 / math.sqrt(d); hp[:, d + 1] = 0.3
    hp = hp.cuda()
    L = torch.empty(B, eng.round_up(n, 128), eng.round_up(n, 128), dtype=torch.float64, device="cuda")
    st = eng.fit(x, y, hp, True, L_out=L); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(reps):
        e0.record(); st = eng.fit(x, y, hp, True, L_out=L); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    fit_flops = B * n ** 3 / 3.0
    print(f"fit     B={B} n={n} d={d}: {best:9.2f} ms  {fit_flops / best * 1e-9:6.2f} TFLOP/s (N^3/3)  info={int(st.info.abs().sum())}", flush=True)
    mean, var = eng.predict(st, xs); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0.record(); mean, var = eng.predict(st, xs); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    pf = B * float(n) ** 2 * m
    print(f"predict B={B} n={n} m={m}: {best:9.2f} ms  {pf / best * 1e-9:6.2f} TFLOP/s (N^2 m)  {m / best * 1e3:9.1f} pred/s", flush=True)

if __name__ == "__main__":
    cfgs = [(32, 2048, 8, 1024), (32, 4096, 16, 1024), (37, 8192, 16, 512), (148, 8192, 16, 1024)]
    if len(sys.argv) > 1:
        cfgs = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
    for c in cfgs:
        run(*c)
