import sys, math, torch, json
sys.path.insert(0, ".")
from pygpr_b200 import engine as eng
def hp_well(B, d):
    hp = torch.empty(B, d + 2, dtype=torch.float64); hp[:, 0] = 1.0; hp[:, 1:d + 1] = 2.0 / math.sqrt(d); hp[:, d + 1] = 0.3
    return hp.cuda()
def timed(fn, reps=2):
    fn(); torch.cuda.synchronize(); best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
g = torch.Generator().manual_seed(0)
for B, n in ((1, 16384), (1, 8192), (1, 4096), (2, 16384), (4, 8192), (8, 8192)):
    d = 8
    x = torch.rand(B, n, d, generator=g, dtype=torch.float64).cuda()
    y = torch.sin(x.sum(-1)) + 0.1 * torch.randn(B, n, generator=g, dtype=torch.float64).cuda()
    hp = hp_well(B, d)
    L = torch.empty(B, n, n, dtype=torch.float64, device="cuda")
    rec = {"B": B, "n": n, "fp64_ms": timed(lambda: eng.fit(x, y, hp, True, L_out=L))}
    for panel in (512, 1024, 2048, 4096):
        if panel * 2 <= n:
            rec["emu7_panel%d_ms" % panel] = timed(lambda: eng.fit_emulated(x, y, hp, True, slices=7, L_out=L, panel=panel))
    flops = B * n ** 3 / 3
    rec["fp64_tflops"] = flops / rec["fp64_ms"] * 1e-9
    rec["best_emu_tflops_equiv"] = flops / min(v for k, v in rec.items() if k.startswith("emu7")) * 1e-9
    print(json.dumps(rec), flush=True)
    del x, y, L
