// Round-1 tile routines (8x8 register blocks per thread), kept beside the microbenchmark as the baseline the round-2 routines of
// pygpr_b200/csrc/gpr_diag_tile.cuh are measured against; not part of the library any more.
#pragma once
#include "gpr_core.cuh"

namespace gpr {

// ---------------------------------------------------------------------------------------------
// Register-resident factorisation of one 128x128 diagonal tile.
// The lower triangle is cut into 136 blocks of 8x8; thread q < 136 owns block (bi, bj) in 64 registers (potf2_blk /
// trtri_blk below).  History (r01): an smem version took ~0.5 ms per tile, a fully unrolled/predicated register version
// 0.36 ms (240 KB of code, 53 % barrier stalls), 128 rolled rank-1 steps with one barrier each ~0.1 ms, the blocked form
// below ~0.07 ms; what is left is phase latency (in-thread 8x8 Cholesky with its sqrt/divide chain) and idle lanes as the
// trailing matrix shrinks.
// ---------------------------------------------------------------------------------------------
struct DiagBlock {
    int bi, bj;       // block row / column (bi >= bj); bi = -1 for idle threads
    double R[8][8];
};

__device__ __forceinline__ void diag_block_coords(DiagBlock& d) {
    const int q = threadIdx.x;
    d.bi = -1;
    d.bj = 0;
    if (q < 136) {
        int bi = 0, rem = q;
        while (rem >= bi + 1) { rem -= bi + 1; ++bi; }
        d.bi = bi;
        d.bj = rem;
    }
}

// ---------------------------------------------------------------------------------------------
// Blocked factorisation / inversion (v2).  v1 ran 128 rank-1 steps (one CTA barrier and 64 DFMA per thread each, ~780 clocks
// per step); here 16 block steps over the 8x8 register blocks: the owner of the diagonal block factorises and
// inverts it in registers, the panel blocks multiply by that inverse, the trailing blocks take one 8x8x8 product
// (512 DFMA behind 32 LDS.128) — 2 barriers per block step instead of 8, and 8x more independent work between them.
// T is the exchange medium: every block is written at its final position.  invblk[kb] = inv(L_kk) (8x8, row major).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int chol8(double (&R)[8][8], double (&dinv)[8]) {
    int fail = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const double p = R[k][k];
        if (!(p > 0.0) && fail == 0) fail = k + 1;
        const double sq = sqrt(p);
        const double inv = 1.0 / sq;
        R[k][k] = sq;
        dinv[k] = inv;
#pragma unroll
        for (int r = k + 1; r < 8; ++r) R[r][k] *= inv;
#pragma unroll
        for (int c = k + 1; c < 8; ++c)
#pragma unroll
            for (int r = c; r < 8; ++r) R[r][c] -= R[r][k] * R[c][k];
    }
    return fail;
}

__device__ __forceinline__ void load_block8(double (&R)[8][8], const double* T, int bi, int bj) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; c += 2) {
            const double2 v = *reinterpret_cast<const double2*>(T + (bi * 8 + r) * XLD + bj * 8 + c);
            R[r][c] = v.x;
            R[r][c + 1] = v.y;
        }
}
__device__ __forceinline__ void store_block8(const double (&R)[8][8], double* T, int bi, int bj) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; c += 2)
            *reinterpret_cast<double2*>(T + (bi * 8 + r) * XLD + bj * 8 + c) = make_double2(R[r][c], R[r][c + 1]);
}

// Exchange buffer of the blocked routines: the 136 lower-triangular 8x8 blocks, block major, 66 doubles apart.  In the
// row-major tile (stride 132 doubles) blocks that differ by a multiple of 8 rows sit on the same banks and the panel /
// trailing phases ran 8- to 16-way conflicted (clock64 build: 2.8 k + 3.6 k clocks per block step for 288 + 512 DFMA); with
// 528-byte block spacing consecutive block ids rotate through the eight 16-byte bank groups.
constexpr int BLK_LD = 66;
__device__ __forceinline__ int blk_id(int bi, int bj) { return bi * (bi + 1) / 2 + bj; }
__device__ __forceinline__ void store_blk(const double (&R)[8][8], double* Bm, int bi, int bj) {
    double* p = Bm + blk_id(bi, bj) * BLK_LD;
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; c += 2) *reinterpret_cast<double2*>(p + r * 8 + c) = make_double2(R[r][c], R[r][c + 1]);
}
// R -= A * B^T (TRANS_B) or R -= A * B, A = block (ai, aj), B = block (bi, bj) of Bm; two k at a time
template <bool TRANS_B>
__device__ __forceinline__ void block8_mulsub(double (&R)[8][8], const double* Bm, int ai, int aj, int bi, int bj) {
    const double* pa = Bm + blk_id(ai, aj) * BLK_LD;
    const double* pb = Bm + blk_id(bi, bj) * BLK_LD;
#pragma unroll
    for (int m = 0; m < 8; m += 2) {
        double2 av[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) av[r] = *reinterpret_cast<const double2*>(pa + r * 8 + m);  // A[r][m..m+1]
        if (TRANS_B) {
            double2 bv[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) bv[c] = *reinterpret_cast<const double2*>(pb + c * 8 + m);  // B[c][m..m+1]
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int c = 0; c < 8; ++c) R[r][c] -= av[r].x * bv[c].x + av[r].y * bv[c].y;
        } else {
            double2 b0[4], b1[4];  // B[m][0..7], B[m+1][0..7]
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                b0[c] = *reinterpret_cast<const double2*>(pb + m * 8 + 2 * c);
                b1[c] = *reinterpret_cast<const double2*>(pb + (m + 1) * 8 + 2 * c);
            }
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    R[r][2 * c] -= av[r].x * b0[c].x + av[r].y * b1[c].x;
                    R[r][2 * c + 1] -= av[r].x * b0[c].y + av[r].y * b1[c].y;
                }
        }
    }
}

// T: symmetric tile (lower part used) -> L in the lower triangle (strict upper of the diagonal 8x8 blocks zeroed; blocks above
// the block diagonal untouched).  Bm: BM_DOUBLES of scratch, left holding every strictly-lower block of L for trtri_blk.
// invblk: 16 x 64 doubles.  *fail = first 1-based column with a non-positive pivot (unchanged if none).
__device__ __forceinline__ void potf2_blk(double* T, double* Bm, double* invblk, int* fail, DiagBlock& d) {
    const bool live = d.bi >= 0;
    if (live) load_block8(d.R, T, d.bi, d.bj);
#pragma unroll 1
    for (int kb = 0; kb < TILE / 8; ++kb) {
        if (live && d.bi == kb && d.bj == kb) {
            double dinv[8];
            const int f = chol8(d.R, dinv);
            if (f) atomicMin(fail, kb * 8 + f);
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int c = r + 1; c < 8; ++c) d.R[r][c] = 0.0;
            // Y = inv(L_kk), column by column: Y[c][c] = 1/L[c][c], Y[r][c] = -(sum_{c<=m<r} L[r][m] Y[m][c]) / L[r][r]
            double* Y = invblk + kb * 64;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                double y[8];
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    if (r < c) {
                        y[r] = 0.0;
                    } else if (r == c) {
                        y[r] = dinv[c];
                    } else {
                        double s = 0.0;
#pragma unroll
                        for (int m = c; m < r; ++m) s += d.R[r][m] * y[m];
                        y[r] = -s * dinv[r];
                    }
                    Y[r * 8 + c] = y[r];
                }
            }
        }
        __syncthreads();
        if (live && d.bj == kb && d.bi > kb) {  // panel: R <- R inv(L_kk)^T, in place from the last column down
            const double* Y = invblk + kb * 64;
#pragma unroll
            for (int c = 7; c >= 0; --c) {
                double yc[8];
#pragma unroll
                for (int m = 0; m <= c; ++m) yc[m] = Y[c * 8 + m];
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    double s = d.R[r][c] * yc[c];
#pragma unroll
                    for (int m = 0; m < c; ++m) s += d.R[r][m] * yc[m];
                    d.R[r][c] = s;
                }
            }
            store_blk(d.R, Bm, d.bi, kb);
        }
        __syncthreads();
        if (live && d.bj > kb) block8_mulsub<true>(d.R, Bm, d.bi, kb, d.bj, kb);  // trailing: R -= L[bi,kb] L[bj,kb]^T
    }
    if (live) store_block8(d.R, T, d.bi, d.bj);  // every thread holds its final block of L
    __syncthreads();
}

// X = L^-1 (Bm = the strictly-lower blocks of L from potf2_blk, invblk = its diagonal-block inverses).  On return T holds X in
// the lower triangle and zeros above it.
__device__ __forceinline__ void trtri_blk(double* T, double* Bm, const double* invblk, DiagBlock& d) {
    const bool live = d.bi >= 0;
    for (int id = threadIdx.x; id < TILE * TILE; id += NTHREADS) {  // blocks above the block diagonal
        const int r = id >> 7, c = id & 127;
        if ((c >> 3) > (r >> 3)) T[r * XLD + c] = 0.0;
    }
    if (live) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c < 8; ++c) d.R[r][c] = (d.bi == d.bj && r == c) ? 1.0 : 0.0;
    }
#pragma unroll 1
    for (int kb = 0; kb < TILE / 8; ++kb) {
        if (live && d.bi == kb) {  // block row kb of X is final: X[kb, bj] = inv(L_kk) R[kb, bj], in place from the last row up
            const double* Y = invblk + kb * 64;
#pragma unroll
            for (int r = 7; r >= 0; --r) {
                double yr[8];
#pragma unroll
                for (int m = 0; m <= r; ++m) yr[m] = Y[r * 8 + m];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    double s = d.R[r][c] * yr[r];
#pragma unroll
                    for (int m = 0; m < r; ++m) s += yr[m] * d.R[m][c];
                    d.R[r][c] = s;
                }
            }
            store_blk(d.R, Bm, kb, d.bj);  // over L[kb, bj], which was last read at block step bj < kb
        }
        __syncthreads();
        if (live && d.bi > kb && d.bj <= kb) block8_mulsub<false>(d.R, Bm, d.bi, kb, kb, d.bj);  // R -= L[bi,kb] X[kb,bj]
        __syncthreads();  // row kb+1 is overwritten next; its L blocks were just read by the whole row
    }
    if (live) store_block8(d.R, T, d.bi, d.bj);  // final X blocks (diagonal blocks carry their zeros above the diagonal)
    __syncthreads();
}


}  // namespace gpr
