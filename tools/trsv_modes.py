"""Device time of the backward solve alpha = L^-T z in its two forms (PYGPR_TRSV_MODE=single|auto)."""
import math, os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from pygpr_b200 import engine as eng, _lib
from pygpr_b200 import config
from potrf_modes import data

for B, n, d in [(1, 16384, 8), (1, 4096, 16), (8, 8192, 16), (32, 8192, 16), (63, 4096, 8), (120, 4096, 8)]:
    x, y, hp = data(B, n, d)
    st = eng.fit(x, y, hp, True)
    torch.cuda.synchronize()
    alpha = torch.empty_like(st.z)
    work = torch.empty(int(_lib.load().gpr_trsv_work_bytes(B)), dtype=torch.uint8, device="cuda")
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    for mode in ("single", "auto"):
        config.set_trsv_mode(mode)
        best = 1e9
        for _ in range(4):
            e0.record()
            _lib.call("gpr_trsv_backward", st.L.data_ptr(), st.invdiag.data_ptr(), B, st.npad, st.z.data_ptr(), alpha.data_ptr(),
                      work.data_ptr(), torch.cuda.current_stream().cuda_stream)
            e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        gb = B * st.npad ** 2 * 4 / 1e9
        print(f"trsv {mode:6s} B={B} n={n}: {best:8.3f} ms  {gb / best:7.1f} GB/s (4 n^2 B per expert)  err={float((alpha - st.alpha).abs().max()):.1e}", flush=True)
