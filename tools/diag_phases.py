"""Clock stamps of the phases of ONE diagonal-tile task (potrf_diag_kernel, expert 0, last block column) from a developer build:

    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -shared -Xcompiler -fPIC --expt-relaxed-constexpr -DGPR_DIAG_STAMPS \
         -o pygpr_b200/libpygpr_b200_dev.so pygpr_b200/csrc/gpr_kernels.cu pygpr_b200/csrc/ozaki_gemm.cu
    PYGPR_B200_LIB=pygpr_b200/libpygpr_b200_dev.so python tools/diag_phases.py [n=2048]
"""
import ctypes
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pygpr_b200 import _lib, engine as eng  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
d = 8
g = torch.Generator().manual_seed(0)
x = torch.rand(1, n, d, generator=g, dtype=torch.float64).cuda()
y = torch.sin(x.sum(-1)) + 0.1 * torch.randn(1, n, generator=g, dtype=torch.float64).cuda()
hp = torch.empty(1, d + 2, dtype=torch.float64)
hp[:, 0] = 1.0; hp[:, 1:d + 1] = 2.0 / math.sqrt(d); hp[:, d + 1] = 0.3
hp = hp.cuda()
names = ["start", "product (last block column)", "", "tile assembled", "potf2", "L written + log", "trtri", "inverse written", "z block", "flag"]
for gen in (True, False):
    for it in range(2):
        if gen:
            eng.fit(x, y, hp, True)
        else:
            K = eng.se_matrix(hp, x) if hasattr(eng, "se_matrix") else None
            buf = torch.zeros(1, n, n, dtype=torch.float64, device="cuda")
            a, nrm = eng.scale_inputs(x, hp, 1, n)
            dadd, _ = eng.split_hp(hp, d, True, eng.JITTER)
            _lib.call("gpr_se_build", eng._p(a), eng._p(nrm), eng._p(hp), hp.shape[-1], eng._p(dadd), 1, n, n, a.shape[2], eng._p(buf), n, 1, 1, eng._stream())
            eng.potrf_inplace(buf, y.contiguous())
        torch.cuda.synchronize()
    st = (ctypes.c_longlong * 16)()
    lib = _lib.load()
    lib.gpr_debug_diag_stamps.argtypes = [ctypes.POINTER(ctypes.c_longlong)]
    assert lib.gpr_debug_diag_stamps(st) == 0
    t = list(st)
    out = {"kernel": "potrf_diag_kernel<%s>, n=%d, last column" % ("GEN" if gen else "in place", n), "total_clocks": t[9] - t[0]}
    prev = t[0]
    for i in (1, 3, 4, 5, 6, 7, 8, 9):
        out[names[i]] = t[i] - prev
        prev = t[i]
    print(json.dumps(out), flush=True)
