"""Emulated-FP64 GEMM on the INT8 tcgen05 tensor cores (csrc/ozaki_gemm.cu): correctness and rate against FP64 DGEMM.

    python tools/ozaki_check.py [--quick] [--out gpurun_out/ozaki.jsonl]

1. digit planes / scales bit-exact against the numpy restatement (oracle/ozaki_ref.py), ragged shapes;
2. C against the numpy model (same integer sums, same recombination order) and against the FP64 product;
3. error and TFLOP/s-equivalent (2MNK / time) for slices 3..8 on the trailing-update shape and on the prediction shape with
   REAL operands (K* and L^-1 of a GP at the bench's conditioning), next to cuBLAS DGEMM timed in the same run.
"""
import argparse
import json
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from pygpr_b200 import engine as eng  # noqa: E402
from oracle import ozaki_ref  # noqa: E402  (checker only)


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--out", default=None)
    ap.add_argument("--bk", type=int, default=0, help="force the GEMM's k-block (32 / 64); 0 = automatic")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    from pygpr_b200 import config
    config.set_ozaki_bk(args.bk)
    out = []

    def emit(rec):
        out.append(rec)
        print(json.dumps(rec), flush=True)

    g = torch.Generator().manual_seed(0)
    # ---- 1/2: exactness on small and ragged shapes ----------------------------------------------------------------------------
    for (M, N, K) in ((128, 64, 64), (128, 64, 256), (200, 100, 130), (300, 257, 1000)):
        A = torch.randn(M, K, generator=g, dtype=torch.float64) * torch.exp(torch.randn(M, 1, generator=g, dtype=torch.float64))
        B = torch.randn(N, K, generator=g, dtype=torch.float64)
        ref = A @ B.T
        for s in (1, 3, 6, 7, 8):
            pa = eng.ozaki_split(A.to(dev), s)
            pb = eng.ozaki_split(B.to(dev), s)
            npa, nsa = ozaki_ref.split(A.numpy(), s)
            npb, nsb = ozaki_ref.split(B.numpy(), s)
            planes_ok = bool((pa.planes.cpu().numpy() == npa).all() and (pb.planes.cpu().numpy() == npb).all())
            scales_ok = bool((pa.scale.cpu().numpy() == nsa).all() and (pb.scale.cpu().numpy() == nsb).all())
            C = eng.ozaki_gemm_nt(pa, pb).cpu()
            model = torch.from_numpy(ozaki_ref.gemm_nt(npa, nsa, M, npb, nsb, N, s))
            emit({"case": "exact", "M": M, "N": N, "K": K, "slices": s, "planes_bit_exact": planes_ok, "scales_bit_exact": scales_ok,
                  "max_abs_vs_model": float((C - model).abs().max()), "bit_exact_vs_model": bool((C == model).all()),
                  "rel_vs_fp64": rel(C, ref)})
    # beta = 1 accumulation and the lower-triangular B option
    M, N, K = 256, 192, 192
    A = torch.randn(M, K, generator=g, dtype=torch.float64).to(dev)
    Bm = torch.tril(torch.randn(N, K, generator=g, dtype=torch.float64)).to(dev)
    C0 = torch.randn(M, N, generator=g, dtype=torch.float64).to(dev)
    pa, pb = eng.ozaki_split(A, 7), eng.ozaki_split(Bm, 7)
    C = eng.ozaki_gemm_nt(pa, pb, out=C0.clone(), beta=1.0, b_lower=True)
    emit({"case": "beta1_lower", "rel_vs_fp64": rel(C, C0 + A @ Bm.T)})

    # ---- 3: rate and error at scale ---------------------------------------------------------------------------------------------
    shapes = [("square_4096", 4096, 4096, 4096)] if args.quick else [("square_4096", 4096, 4096, 4096), ("update_8192", 8192, 8192, 8192),
                                                                     ("predict_4096x4096x8192", 4096, 4096, 8192)]
    for name, M, N, K in shapes:
        A = torch.randn(M, K, generator=g, dtype=torch.float64).to(dev)
        B = torch.randn(N, K, generator=g, dtype=torch.float64).to(dev)
        ref = torch.empty(M, N, dtype=torch.float64, device=dev)
        t_dgemm = timed(lambda: torch.matmul(A, B.T, out=ref))
        flops = 2.0 * M * N * K
        emit({"case": name, "impl": "cublas_dgemm", "ms": t_dgemm, "tflops": flops / t_dgemm * 1e-9})
        C = torch.empty(M, N, dtype=torch.float64, device=dev)
        for s in (3, 4, 5, 6, 7, 8):
            t_split = timed(lambda: (eng.ozaki_split(A, s), eng.ozaki_split(B, s)))
            pa, pb = eng.ozaki_split(A, s), eng.ozaki_split(B, s)
            t_gemm = timed(lambda: eng.ozaki_gemm_nt(pa, pb, out=C))
            emit({"case": name, "impl": "ozaki_i8_tcgen05", "slices": s, "plane_products": s * (s + 1) // 2, "gemm_ms": t_gemm,
                  "split_ms": t_split, "tflops_equiv_gemm": flops / t_gemm * 1e-9, "tflops_equiv_with_split": flops / (t_gemm + t_split) * 1e-9,
                  "int8_tops": flops * (s * (s + 1) // 2) / t_gemm * 1e-9, "speedup_vs_dgemm": t_dgemm / t_gemm,
                  "rel_vs_fp64": rel(C, ref)})
            del pa, pb
        del A, B, ref, C

    # ---- real operands: K* (test x train cross-covariances) times L^-1 of a GP at the bench's conditioning --------------------------
    n, d, m = (2048, 16, 1024) if args.quick else (8192, 16, 4096)
    x = torch.rand(1, n, d, generator=g, dtype=torch.float64)
    y = torch.sin(x.sum(-1)) + 0.1 * torch.randn(1, n, generator=g, dtype=torch.float64)
    hp = torch.empty(1, d + 2, dtype=torch.float64)
    hp[:, 0] = 1.0; hp[:, 1:d + 1] = 2.0 / math.sqrt(d); hp[:, d + 1] = 0.3
    xs = torch.rand(m, d, generator=g, dtype=torch.float64)
    st = eng.fit(x.to(dev), y.to(dev), hp.to(dev), True)
    X = eng.tril_inverse(st)[0]                                   # L^-1 [n, n], lower triangular
    Ks = eng.se_cross_matrix(hp.to(dev), x.to(dev), xs.to(dev))[0]  # K* [m, n]
    ref = Ks @ X.T                                                # V^T = K* L^-T  (FP64, cuBLAS: checker)
    t_dgemm = timed(lambda: torch.matmul(Ks, X.T))
    flops = 2.0 * m * n * n
    emit({"case": "real_KsLinvT", "impl": "cublas_dgemm", "m": m, "n": n, "ms": t_dgemm, "tflops_dense_count": flops / t_dgemm * 1e-9})
    var_ref = (ref * ref).sum(1)
    for s in (3, 4, 5, 6, 7, 8):
        pa, pb = eng.ozaki_split(Ks, s), eng.ozaki_split(X, s)
        C = torch.empty(m, n, dtype=torch.float64, device=dev)
        t_full = timed(lambda: eng.ozaki_gemm_nt(pa, pb, out=C))
        e_full = rel(C, ref)
        t_tri = timed(lambda: eng.ozaki_gemm_nt(pa, pb, out=C, b_lower=True))
        emit({"case": "real_KsLinvT", "impl": "ozaki_i8_tcgen05", "slices": s, "ms_full": t_full, "ms_lower": t_tri,
              "tflops_equiv_full": flops / t_full * 1e-9, "tflops_equiv_lower_dense_count": flops / t_tri * 1e-9,
              "speedup_vs_dgemm_full": t_dgemm / t_full, "rel_V": e_full, "rel_V_lower": rel(C, ref),
              "rel_sumV2": rel((C * C).sum(1), var_ref)})
    if args.out:
        with open(args.out, "w") as f:
            for r in out:
                f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
