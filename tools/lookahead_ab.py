"""A/B of the panel look-ahead of engine.fit_emulated (config.options.emulate_fit_lookahead) on one matrix of order 16384 and on
8 matrices of order 8192: alternating runs in one process, every time listed (the side stream shares the GPU with the trailing
update's tensor-core launch: the interest is in the spread as much as in the best time)."""
import json, math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pygpr_b200 import config, engine as eng

def hp_well(B, d):
    hp = torch.empty(B, d + 2, dtype=torch.float64); hp[:, 0] = 1.0; hp[:, 1:d + 1] = 2.0 / math.sqrt(d); hp[:, d + 1] = 0.3
    return hp.cuda()

g = torch.Generator().manual_seed(0)
for B, n in ((1, 16384), (8, 8192)):
    d = 8
    x = torch.rand(B, n, d, generator=g, dtype=torch.float64).cuda()
    y = torch.sin(x.sum(-1)) + 0.1 * torch.randn(B, n, generator=g, dtype=torch.float64).cuda()
    hp = hp_well(B, d)
    L = torch.empty(B, n, n, dtype=torch.float64, device="cuda")
    times = {True: [], False: []}
    for rep in range(9):
        for la in (True, False):
            config.options.emulate_fit_lookahead = la
            hp2 = hp * (1.0 + 0.01 * rep)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); eng.fit_emulated(x, y, hp2, True, slices=7, L_out=L); e1.record(); torch.cuda.synchronize()
            if rep:
                times[la].append(round(e0.elapsed_time(e1), 2))
    print(json.dumps({"B": B, "n": n, "lookahead_ms": times[True], "inline_ms": times[False]}), flush=True)
    del x, y, L
