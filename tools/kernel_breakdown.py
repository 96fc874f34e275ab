"""Per-kernel device time of one call of the path (CUPTI through torch.profiler — every kernel of the process is seen, also the
ones launched through the C-ABI).  Shares, not benchmark numbers: the profiler adds launch overhead.

    python tools/kernel_breakdown.py fit1 [n=16384] [panel=1024] [slices=7]     one matrix, emulated factorisation vs FP64
    python tools/kernel_breakdown.py predict [experts=16] [slices=6]            shared-block GRBCM prediction, emulated
    python tools/kernel_breakdown.py timeline [n=4096] [first=1500] [count=40]  start / duration of consecutive kernels of engine.fit, B=1
"""
import json
import math
import os
import sys
from collections import defaultdict

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pygpr_b200 import engine as eng  # noqa: E402


def breakdown(fn, label):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    wall = e0.elapsed_time(e1)
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        fn(); torch.cuda.synchronize()
    agg = defaultdict(lambda: [0.0, 0])
    for ev in prof.events():
        if ev.device_type.name == "CUDA":
            name = ev.name.split("(")[0].replace("void ", "")
            agg[name][0] += ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
            agg[name][1] += 1
    rows = sorted(((v[0] * 1e-3, v[1], k) for k, v in agg.items()), reverse=True)
    total = sum(r[0] for r in rows)
    print(json.dumps({"case": label, "unprofiled_ms": wall, "kernel_sum_ms": total,
                      "kernels": [{"name": k[:70], "ms": round(ms, 3), "launches": c} for ms, c, k in rows[:14]]}), flush=True)


def hp_well(B, d):
    hp = torch.empty(B, d + 2, dtype=torch.float64)
    hp[:, 0] = 1.0; hp[:, 1:d + 1] = 2.0 / math.sqrt(d); hp[:, d + 1] = 0.3
    return hp.cuda()


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "fit1"
    nums = [int(a) for a in sys.argv[2:]]
    g = torch.Generator().manual_seed(0)
    if what == "timeline":
        n = nums[0] if nums else 4096
        first = nums[1] if len(nums) > 1 else 1500
        count = nums[2] if len(nums) > 2 else 40
        d = 8
        x = torch.rand(1, n, d, generator=g, dtype=torch.float64).cuda()
        y = (torch.sin(x.sum(-1)) + 0.1 * torch.randn(1, n, generator=g, dtype=torch.float64).cuda())
        hp = hp_well(1, d)
        eng.fit(x, y, hp, True); torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            eng.fit(x, y, hp, True); torch.cuda.synchronize()
        evs = sorted((ev for ev in prof.events() if ev.device_type.name == "CUDA"), key=lambda ev: ev.time_range.start)
        t0 = evs[0].time_range.start
        sel = [ev for ev in evs if ev.time_range.start - t0 >= first][:count]
        for ev in sel:
            print("%9.1f us  +%7.1f us  %s" % (ev.time_range.start - t0, ev.time_range.end - ev.time_range.start,
                                              ev.name.split("(")[0].replace("void ", "")[:60]), flush=True)
        return
    if what == "fit1":
        n = nums[0] if nums else 16384
        panel = nums[1] if len(nums) > 1 else 1024
        S = nums[2] if len(nums) > 2 else 7
        d = 8
        x = torch.rand(1, n, d, generator=g, dtype=torch.float64).cuda()
        y = (torch.sin(x.sum(-1)) + 0.1 * torch.randn(1, n, generator=g, dtype=torch.float64).cuda())
        hp = hp_well(1, d)
        L = torch.empty(1, n, n, dtype=torch.float64, device="cuda")
        breakdown(lambda: eng.fit(x, y, hp, True, L_out=L), "fp64 fit n=%d B=1" % n)
        breakdown(lambda: eng.fit_emulated(x, y, hp, True, slices=S, L_out=L, panel=panel), "emulated fit n=%d B=1 panel=%d S=%d" % (n, panel, S))
    else:
        B = nums[0] if nums else 16
        S = nums[1] if len(nums) > 1 else 6
        m0, d, m = 4096, 16, 4096
        xg = torch.rand(1, m0, d, generator=g, dtype=torch.float64)
        yg = torch.sin(xg.sum(-1))
        xl = torch.rand(B, m0, d, generator=g, dtype=torch.float64)
        yl = torch.sin(xl.sum(-1))
        x = torch.cat((xg.expand(B, -1, -1), xl), 1).cuda()
        y = torch.cat((yg.expand(B, -1), yl), 1).cuda()
        hp = hp_well(B, d)
        st_g = eng.fit(xg.cuda(), yg.cuda(), hp[:1], True)
        st_l = eng.fit(x, y, hp, True, shared=st_g)
        xs = torch.rand(m, d, generator=g, dtype=torch.float64).cuda()
        breakdown(lambda: eng.predict_shared_emulated(st_g, st_l, xs, S, 8), "emulated shared-block prediction B=%d S=%d" % (B, S))


if __name__ == "__main__":
    main()
