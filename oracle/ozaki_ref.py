"""CPU restatement (numpy) of the digit-plane arithmetic of pygpr_b200/csrc/ozaki_gemm.cu — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's checker legs may import this module; the product never does.

The reference (sarath-srinivas/PyGPR) has no counterpart: its products are FP64 `tc.matmul` / `tc.cholesky_solve`
(gpr.py:80,100).  This file pins the INTEGER part of the emulated-FP64 path bit for bit — the digit planes and row
scales `gpr_ozaki_split` must produce, the int32 diagonals `D_g`, and the FP64 recombination in the same operation order as
the kernel's epilogue — so that the GPU result can be checked exactly (digits, scales) and to the last rounding (C).
"""
from __future__ import annotations

import numpy as np

BM, BK = 128, 64


def round_up(v: int, m: int) -> int:
    return (v + m - 1) // m * m


def digits_by_bias(I: np.ndarray, slices: int) -> np.ndarray:
    """Balanced radix-256 digits of int64 I (|I| < 2^(8 slices - 2)) WITHOUT the carry loop, as se_cross_planes_kernel forms them:
    with J = I + sum_t 128 * 256^t the bytes of J are d_t + 128, so the digits are the bytes of J ^ 0x80..80.  Returns
    [slices, ...] int8, most significant digit first (plane 0), like split()."""
    bias = np.uint64(sum(0x80 << (8 * t) for t in range(slices)))
    with np.errstate(over="ignore"):                          # modulo 2^64, like the kernel's unsigned long long
        W = (np.asarray(I, dtype=np.int64).astype(np.uint64) + bias) ^ bias
    return np.stack([((W >> np.uint64(8 * (slices - 1 - t))) & np.uint64(0xFF)).astype(np.uint8).view(np.int8) for t in range(slices)])


def split(X: np.ndarray, slices: int, bound: float = 0.0):
    """planes [slices, rpad, kpad] int8, scale [rpad] float64 — split_kernel.  bound > 0: every row is scaled by that bound instead of
    its own maximum (the caller vouches |x| <= bound; gpr_ozaki_split(bound), gpr_se_cross_planes)."""
    X = np.asarray(X, dtype=np.float64)
    rows, k = X.shape
    rpad, kpad = round_up(rows, BM), round_up(k, BK)
    planes = np.zeros((slices, rpad, kpad), dtype=np.int8)
    scale = np.zeros(rpad, dtype=np.float64)
    amax = np.abs(X).max(axis=1) if not bound > 0.0 else np.full(rows, float(bound))
    for r in range(rows):
        a = amax[r]
        if not (a > 0.0 and np.isfinite(a)):
            continue
        e = int(np.floor(np.log2(a)))
        # ilogb: exact exponent (guard against log2 rounding at powers of two)
        while np.ldexp(1.0, e + 1) <= a:
            e += 1
        while np.ldexp(1.0, e) > a:
            e -= 1
        e += 3
        scale[r] = np.ldexp(1.0, e - 8)
        I = np.rint(X[r] * np.ldexp(1.0, 8 * slices - e)).astype(np.int64)
        for t in range(slices - 1, -1, -1):
            dgt = (I & 0xFF).astype(np.uint8).view(np.int8)
            planes[t, r, :k] = dgt
            I = (I - dgt.astype(np.int64)) >> 8
        assert not I.any()
    return planes, scale


def gemm_nt(pa, sa, M, pb, sb, N, slices):
    """C [M, N] from digit planes: exact int64 plane products (the kernel's int32 sums never overflow for K*slices < 2^17),
    Horner recombination from the smallest diagonal as in the kernel's epilogue."""
    acc = np.zeros((M, N), dtype=np.float64)
    A = pa[:, :M].astype(np.int64)
    B = pb[:, :N].astype(np.int64)
    for g in range(slices - 1, -1, -1):
        D = np.zeros((M, N), dtype=np.int64)
        for t in range(g + 1):
            D += A[t] @ B[g - t].T
        assert np.abs(D).max() < 2 ** 31
        acc = acc * 0.00390625 + D.astype(np.float64)
    return acc * sa[:M, None] * sb[None, :N]
