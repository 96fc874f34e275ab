"""One emulated-FP64 GEMM (csrc/ozaki_gemm.cu) for ncu:  python tools/ozaki_profile.py M N K slices [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pygpr_b200 import engine as eng  # noqa: E402

M, N, K, s = (int(v) for v in sys.argv[1:5])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
A = torch.randn(M, K, generator=g, dtype=torch.float64).to(dev)
B = torch.randn(N, K, generator=g, dtype=torch.float64).to(dev)
pa, pb = eng.ozaki_split(A, s), eng.ozaki_split(B, s)
C = torch.empty(M, N, dtype=torch.float64, device=dev)
for _ in range(reps):
    eng.ozaki_gemm_nt(pa, pb, out=C)
torch.cuda.synchronize()
ref = A @ B.T
print("rel err", float((C - ref).abs().max() / ref.abs().max()))
