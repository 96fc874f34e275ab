// pygpr_b200 — factorisation and inversion of ONE 128x128 diagonal tile inside a CTA: the serial link of the Cholesky's
// dependency chain (tc.cholesky, gpr.py:69), which bounds the factorisation of one or a few matrices (BASELINE configs 1-3 and the
// communication expert) while the tensor pipe idles.
//
// v3 of the tile routines (v2 = the 8x8 register blocks of gpr_core.cuh: 16 block steps, each an in-thread 8x8 Cholesky with its
// DSQRT / DDIV chain, 136 of 256 lanes busy and fewer as the trailing matrix shrinks).  Here the tile is cut into four 32-wide
// block columns:
//   chol32_warp   one warp, lane = row, the row in registers; right-looking, one pivot per step: the pivot is broadcast by shuffle,
//                 1/sqrt(p) is MUFU.RSQ64H + two Newton steps (no DSQRT, no DDIV), the new column goes through shared memory so the
//                 rank-1 update is broadcast LDS.128 + DFMA; the lane's own diagonal entry is carried separately so that the next
//                 pivot never waits for that round trip                                          (~130 clocks per column)
//   inv32_warp    same warp, lane = column of inv(L_pp), right-looking forward substitution                       (~1.3 k clocks)
//   panel         rows below:  P = A inv(L_pp)^T           DMMA, every warp owns 8 rows
//   trailing      T[below, below] -= P P^T (lower blocks)  DMMA, (8 rows x 32 columns) units dealt round-robin to the 8 warps
// and the inverse of the tile is assembled from the four diagonal inverses by DMMA products, block column by block column in place.
#pragma once
#include "gpr_core.cuh"

namespace gpr {

constexpr int DT_NB = 32;                                  // panel width
constexpr int XD_LD = 36;                                  // row stride of the scratch blocks (36 mod 16 = 4: conflict-free fragments)
// scratch (doubles) behind the tile: 4 diagonal inverses [32][XD_LD], the columns of the current L_pp [32][32], 1/L_kk [128],
// the exchange block of trtri_tile [32][XD_LD] (overlays the column buffer: never live together)
constexpr int DT_XD = 0;
constexpr int DT_LT = DT_XD + 4 * DT_NB * XD_LD;
constexpr int DT_DINV = DT_LT + DT_NB * XD_LD;
constexpr int DT_DOUBLES = DT_DINV + TILE;
static_assert(DT_DOUBLES <= BM_DOUBLES, "the tile scratch must fit the exchange buffer behind the tile");

// y ~ p^-1/2 to FP64 rounding: hardware seed (relative error 2^-22) and two Newton steps  y <- y + (y/2)(1 - p y^2)
__device__ __forceinline__ double rsqrt_nr(double p) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(p));
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const double t = p * y;
        const double e = fma(-t, y, 1.0);
        y = fma(0.5 * y, e, y);
    }
    return y;
}

// Cholesky of the 32x32 block at T[r0.., r0..] (lower part read), by ONE warp.  On return the block's lower triangle holds L_pp,
// LT[k][r] = L_pp[r][k] (r >= k; rows of 32 doubles), dinv[k] = 1 / L_pp[k][k].  Returns the first 1-based column with a
// non-positive pivot (0 = none).
__device__ __forceinline__ int chol32_warp(double* T, int r0, double* LT, double* dinv) {
    const int lane = threadIdx.x & 31;
    double R[DT_NB];
    double* row = T + (r0 + lane) * XLD + r0;
#pragma unroll
    for (int c = 0; c < DT_NB; c += 2) {
        const double2 v = *reinterpret_cast<const double2*>(row + c);
        R[c] = v.x;
        R[c + 1] = v.y;
    }
    double dg = row[lane];                                 // this lane's diagonal entry, updated alongside R[lane]
    int fail = 0;
#pragma unroll
    for (int k = 0; k < DT_NB; ++k) {
        const double p = __shfl_sync(0xffffffffu, dg, k);
        if (!(p > 0.0) && fail == 0) fail = k + 1;
        const double y = rsqrt_nr(p);
        const double l = R[k] * y;                         // lane k: p * y = sqrt(p); lanes above k: never read again
        R[k] = l;
        dg = fma(-l, l, dg);
        LT[k * DT_NB + lane] = l;
        if (lane == 0) dinv[k] = y;
        __syncwarp();
        // rank-1 update of the columns to the right: R[c] -= l * L[c][k]
        if (k + 1 < DT_NB) {
            const int c0 = (k + 1) & ~1;
#pragma unroll
            for (int c = 0; c < DT_NB; c += 2) {
                if (c >= c0) {
                    const double2 lc = *reinterpret_cast<const double2*>(LT + k * DT_NB + c);
                    if (c > k) R[c] = fma(-l, lc.x, R[c]);
                    R[c + 1] = fma(-l, lc.y, R[c + 1]);
                }
            }
        }
    }
    __syncwarp();
#pragma unroll
    for (int c = 0; c < DT_NB; c += 2) *reinterpret_cast<double2*>(row + c) = make_double2(R[c], R[c + 1]);
    return fail;
}

// X = inv(L_pp) [32][XD_LD] (zeros above the diagonal) from LT / dinv of chol32_warp, by the same warp: lane = column of X.
__device__ __forceinline__ void inv32_warp(const double* LT, const double* dinv, double* X) {
    const int lane = threadIdx.x & 31;
    double s[DT_NB];
#pragma unroll
    for (int m = 0; m < DT_NB; ++m) s[m] = (m == lane) ? 1.0 : 0.0;
#pragma unroll
    for (int m = 0; m < DT_NB; ++m) {
        const double x = s[m] * dinv[m];
        X[m * XD_LD + lane] = x;
        const int r0 = (m + 1) & ~1;
#pragma unroll
        for (int r = 0; r < DT_NB; r += 2) {
            if (r >= r0) {
                const double2 lr = *reinterpret_cast<const double2*>(LT + m * DT_NB + r);   // L[r][m], L[r+1][m]
                if (r > m) s[r] = fma(-lr.x, x, s[r]);
                s[r + 1] = fma(-lr.y, x, s[r + 1]);
            }
        }
    }
}

// Warp-level DMMA helpers on shared-memory operands.  Lane (g = lane / 4, t = lane % 4) holds A[g][t], B[t][g], C[g][2t .. 2t+1].
// C(8 x 8 NBLK) += A(8 x K) * Bt(8 NBLK x K)^T      A rows at A + g lda, Bt rows at Bt + (8 n + g) ldb, k contiguous in both
template <int NBLK, int K, bool NEG_A, bool B_LOWER>
__device__ __forceinline__ void warp_mma_nt(double (&c)[NBLK][2], const double* A, int lda, const double* Bt, int ldb, int nmax = NBLK) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    double a[K / 4];
#pragma unroll
    for (int k = 0; k < K / 4; ++k) {
        const double v = A[g * lda + 4 * k + t];
        a[k] = NEG_A ? -v : v;
    }
#pragma unroll
    for (int n = 0; n < NBLK; ++n) {
        if (n < nmax) {
#pragma unroll
            for (int k = 0; k < K / 4; ++k)
                if (!B_LOWER || 4 * k < 8 * (n + 1)) dmma(c[n], a[k], Bt[(8 * n + g) * ldb + 4 * k + t]);   // Bt[j][k] = 0 for k > j
        }
    }
}
// C(8 x 8 NBLK) += A(8 x K) * B(K x 8 NBLK)          A rows at A + g lda (k contiguous), B rows (k) at B + k ldb
template <int NBLK, bool NEG_A>
__device__ __forceinline__ void warp_mma_nn(double (&c)[NBLK][2], const double* A, int lda, const double* B, int ldb, int K) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll 4
    for (int k = 0; k < K; k += 4) {
        const double v = A[g * lda + k + t];
        const double a = NEG_A ? -v : v;
#pragma unroll
        for (int n = 0; n < NBLK; ++n) dmma(c[n], a, B[(k + t) * ldb + 8 * n + g]);
    }
}

// T: symmetric tile (lower part used) -> L in the lower triangle (the strict upper part is left as it was / as scratch values).
// S: DT_DOUBLES of scratch; on return S + DT_XD holds the four diagonal-block inverses for trtri_tile.
// *fail = first 1-based column with a non-positive pivot (atomicMin; unchanged if none).
__device__ __forceinline__ void potf2_tile(double* T, double* S, int* fail) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    double* LT = S + DT_LT;
    double* dinv = S + DT_DINV;
#pragma unroll 1
    for (int p = 0; p < TILE / DT_NB; ++p) {
        const int r0 = p * DT_NB;
        double* Xp = S + DT_XD + p * DT_NB * XD_LD;
        if (warp == 0) {
            const int f = chol32_warp(T, r0, LT, dinv + r0);
            if (f && lane == 0) atomicMin(fail, r0 + f);
            inv32_warp(LT, dinv + r0, Xp);
        }
        __syncthreads();
        const int t0 = r0 + DT_NB, nb = TILE - t0;        // rows below the diagonal block
        if (nb == 0) break;
        // panel: P = A inv(L_pp)^T, in place; a warp owns whole rows, so only its own reads precede its writes
        for (int rg = warp; rg < nb / 8; rg += NTHREADS / 32) {
            double* Arow = T + (t0 + rg * 8) * XLD + r0;
            double c[4][2];
#pragma unroll
            for (int n = 0; n < 4; ++n) c[n][0] = c[n][1] = 0.0;
            warp_mma_nt<4, DT_NB, false, true>(c, Arow, XLD, Xp, XD_LD);
            __syncwarp();
#pragma unroll
            for (int n = 0; n < 4; ++n) *reinterpret_cast<double2*>(Arow + g * XLD + 8 * n + 2 * t) = make_double2(c[n][0], c[n][1]);
        }
        __syncthreads();
        // trailing update of the lower blocks: unit (i, J) = rows t0 + 8 i .., columns t0 + 32 J .. (J <= i / 4)
        int u = 0;
        for (int i = 0; i < nb / 8; ++i)
            for (int J = 0; J <= i / 4; ++J, ++u) {
                if ((u & 7) != warp) continue;
                const int nmax = (J == i / 4) ? (i & 3) + 1 : 4;     // 8-column blocks on / below the diagonal
                double* Crow = T + (t0 + 8 * i) * XLD + t0 + DT_NB * J;
                double c[4][2];
#pragma unroll
                for (int n = 0; n < 4; ++n) {
                    const double2 v = *reinterpret_cast<const double2*>(Crow + g * XLD + 8 * n + 2 * t);
                    c[n][0] = v.x;
                    c[n][1] = v.y;
                }
                warp_mma_nt<4, DT_NB, true, false>(c, T + (t0 + 8 * i) * XLD + r0, XLD, T + (t0 + DT_NB * J) * XLD + r0, XLD, nmax);
#pragma unroll
                for (int n = 0; n < 4; ++n)
                    if (n < nmax) *reinterpret_cast<double2*>(Crow + g * XLD + 8 * n + 2 * t) = make_double2(c[n][0], c[n][1]);
            }
        __syncthreads();
    }
}

// X = L^-1 of the tile factorised by potf2_tile (T = L in the lower triangle, S = its scratch).  On return T holds X in the lower
// triangle and zeros above it.  Block column j, rows i > j:  X_ij = -X_ii (sum_{j <= k < i} L_ik X_kj)  — the sum is ONE product
// over the k-range [32 j, 32 i) because the blocks are adjacent; X_ij then replaces L_ij, which no later product reads.
__device__ __forceinline__ void trtri_tile(double* T, double* S) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    __syncthreads();                                       // the caller may still be reading L (its diagonal, for the log-determinant)
    for (int id = threadIdx.x; id < TILE * TILE; id += NTHREADS) {
        const int r = id >> 7, c = id & 127;
        if ((c >> 5) == (r >> 5)) T[r * XLD + c] = S[DT_XD + ((r >> 5) * DT_NB + (r & 31)) * XD_LD + (c & 31)];   // X_pp (zeros above)
        else if (c > r) T[r * XLD + c] = 0.0;
    }
    __syncthreads();
    double* E = S + DT_LT;                                 // exchange block [32][XD_LD]
    const int rw = (warp & 3) * 8, cw = (warp >> 2) * 16;  // this warp's 8 rows x 16 columns of a 32 x 32 block
#pragma unroll 1
    for (int j = 0; j < TILE / DT_NB - 1; ++j)
#pragma unroll 1
        for (int i = j + 1; i < TILE / DT_NB; ++i) {
            double c[2][2];
            c[0][0] = c[0][1] = c[1][0] = c[1][1] = 0.0;
            warp_mma_nn<2, false>(c, T + (DT_NB * i + rw) * XLD + DT_NB * j, XLD, T + (DT_NB * j) * XLD + DT_NB * j + cw, XLD, DT_NB * (i - j));
#pragma unroll
            for (int n = 0; n < 2; ++n) *reinterpret_cast<double2*>(E + (rw + g) * XD_LD + cw + 8 * n + 2 * t) = make_double2(c[n][0], c[n][1]);
            __syncthreads();
            c[0][0] = c[0][1] = c[1][0] = c[1][1] = 0.0;
            warp_mma_nn<2, true>(c, T + (DT_NB * i + rw) * XLD + DT_NB * i, XLD, E + cw, XD_LD, DT_NB);
#pragma unroll
            for (int n = 0; n < 2; ++n)
                *reinterpret_cast<double2*>(T + (DT_NB * i + rw + g) * XLD + DT_NB * j + cw + 8 * n + 2 * t) = make_double2(c[n][0], c[n][1]);
            __syncthreads();
        }
}

}  // namespace gpr
