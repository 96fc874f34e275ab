// pygpr_b200 — shared device code: the FP64 tensor-core (DMMA) tile engine.
//
// Every O(N^3) step of the hot path (Cholesky trailing products, the triangular solves of the
// predictive variance, L^-T and K^-1 = U U^T for the LML gradient) is expressed as the SAME
// 128x128 "NT" tile product
//        acc[r][c] += sum_k A[r][k] * B[c][k]        A, B row-major, k contiguous
// so one software-pipelined mainloop serves all of them.  FP64 on sm_100a has no tcgen05 kind;
// the tensor path is warp-level mma.sync (SASS DMMA.8x8x4), measured at 37.0 TFLOP/s on B200
// (profiles/r01_fp64_issue_rates.log) = 1 DMMA / 16 clk / SM sub-partition, so the design goal is
// simply to keep one DMMA issuing every 16 clocks per scheduler: 8 warps x (32x64) register tiles,
// 12 LDS.64 per 32 DMMA, cp.async (LDGSTS, L2-only) 4-stage pipeline, conflict-free padded smem.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gpr {

constexpr int TILE = 128;                      // CTA tile edge
constexpr int KC = 16;                         // k-chunk (doubles) per pipeline stage = 128 B rows
constexpr int SLD = 20;                        // smem row stride of a chunk tile; 20 mod 16 = 4 -> conflict-free 64-bit fragment loads
#ifndef GPR_STAGES
#define GPR_STAGES 4
#endif
constexpr int STAGES = GPR_STAGES;
constexpr int NTHREADS = 256;                  // 8 warps: 4 (rows) x 2 (cols), warp tile 32 x 64
constexpr int STAGE_DOUBLES = 2 * TILE * SLD;  // A chunk tile + B chunk tile
constexpr int MAIN_SMEM_BYTES = STAGES * STAGE_DOUBLES * 8;  // 163,840 B
constexpr int XLD = 132;                       // row stride of a full 128x128 smem tile (132 mod 16 = 4)
static_assert((KC / 4) % 2 == 0, "fragment double buffering across chunks needs an even number of k4 blocks");
static_assert(TILE * XLD * 8 <= MAIN_SMEM_BYTES, "full tile must overlay the pipeline stages");

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

// ---- mbarrier pipeline state -----------------------------------------------------------------------------------
// full[s]  : completes when all 256 threads' cp.async copies of the chunk in stage s have landed
//            (cp.async.mbarrier.arrive.noinc, one per thread and phase)
// empty[s] : completes when all 8 warps have finished reading stage s (one arrive per warp)
// A running chunk counter g gives stage = g % STAGES and parity = (g / STAGES) & 1 across mainloop calls.
struct Pipe {
    unsigned bars;   // shared-space address of 2*STAGES mbarriers: full[0..STAGES), empty[0..STAGES)
    unsigned g;      // chunks consumed so far by this CTA (identical in every thread)
};
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive(unsigned bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
// Call once per kernel (all threads), before the first mainloop.
__device__ __forceinline__ void pipe_init(Pipe& p, unsigned long long* bars_smem) {
    p.bars = (unsigned)__cvta_generic_to_shared(bars_smem);
    p.g = 0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(p.bars + 8 * s, NTHREADS);
            mbar_init(p.bars + 8 * (STAGES + s), NTHREADS / 32);
        }
    }
    __syncthreads();
}

// D(8x8) += A(8x4, row) * B(4x8, col):  lane (g=lane/4, t=lane%4) holds A[g][t], B[t][g], C[g][2t..2t+1]
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c[0]), "+d"(c[1])
        : "d"(a), "d"(b));
}

// One contiguous run of k-chunks: A rows at A + r*lda, B rows at B + c*ldb, both advancing 16 doubles per chunk.
struct Seg {
    const double* A;
    long long lda;
    const double* B;
    long long ldb;
    int nchunks;
};

struct NoHook {
    __device__ __forceinline__ void pre(int) {}
    __device__ __forceinline__ void post(const double*, int) {}
};

__device__ __forceinline__ void zero_acc(double (&acc)[4][8][2]) {
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 8; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
}

// acc += A * B^T over the chunks of s0 then s1.  SAME_AB: the B tile is the A tile (diagonal blocks) and is
// loaded once.  HOOK gets every landed A chunk tile (used to fuse the forward solve into the Cholesky).
// Ends with a __syncthreads() after the last read, so the caller may reuse `smem` immediately.
//
// Pipeline (r01 ncu history): v1 issued 8 LDGSTS per thread in a burst after a CTA barrier per chunk -> the LSU queue
// backed up and the tensor pipe idled ~12 % of every chunk.  v2 sliced the copies between the k4 blocks and double
// buffered the fragments (DMMA active 78 -> 86 %).  v3 (this): no CTA-wide barrier in the loop at all — per-stage
// full/empty mbarriers; chunk c+2 is issued during chunk c into the stage freed at chunk c-2, so a warp only ever waits
// for a stage that every warp left a whole chunk ago, and warps drift apart instead of hitting LDS/LDGSTS in lockstep.
constexpr int PREFETCH = STAGES - 2;

// XPF: fetch the first fragments of chunk c+1 during the last k4 block of chunk c (measured per kernel, see the XPF_* constants).
template <bool SAME_AB, bool XPF = false, typename HOOK>
__device__ __forceinline__ void mainloop(double (&acc)[4][8][2], const Seg& s0, const Seg& s1, double* smem, HOOK& hook,
                                         Pipe& pipe) {
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp & 3, wn = warp >> 2;
    const int n0 = s0.nchunks;
    const int total = n0 + s1.nchunks;
    const unsigned g0 = pipe.g;
    // per-thread copy coordinates: slice q moves rows (tid>>3) + 32 q, 16-byte column (tid&7)
    const int crow = tid >> 3, ccol = (tid & 7) * 2;
    const long long offA0 = crow * s0.lda + ccol, offB0 = crow * s0.ldb + ccol;
    const long long offA1 = crow * s1.lda + ccol, offB1 = crow * s1.ldb + ccol;
    const int soff = crow * SLD + ccol;

    auto issue_slice = [&](int c, int q) {  // one quarter (32 rows) of chunk c; q == 0 first waits for the stage to be free
        if (c < total) {
            const unsigned gc = g0 + c;
            const unsigned st = gc % STAGES;
            if (q == 0) mbar_wait(pipe.bars + 8 * (STAGES + st), ((gc / STAGES) & 1) ^ 1);
            const bool first = c < n0;
            const int lc = first ? c : c - n0;
            double* sA = smem + st * STAGE_DOUBLES + soff + q * 32 * SLD;
            const double* gA = (first ? s0.A + offA0 + (long long)q * 32 * s0.lda : s1.A + offA1 + (long long)q * 32 * s1.lda) + (long long)lc * KC;
            cp_async16(sA, gA);
            if (!SAME_AB) {
                const double* gB = (first ? s0.B + offB0 + (long long)q * 32 * s0.ldb : s1.B + offB1 + (long long)q * 32 * s1.ldb) + (long long)lc * KC;
                cp_async16(sA + TILE * SLD, gB);
            }
            if (q == 3) cp_async_mbar_arrive(pipe.bars + 8 * st);
        }
    };

#pragma unroll
    for (int c = 0; c < PREFETCH; ++c) {
#pragma unroll
        for (int q = 0; q < 4; ++q) issue_slice(c, q);
    }

    const int aoff = (wm * 32 + g) * SLD + t;
    const int boff = (wn * 64 + g) * SLD + t;
    if constexpr (XPF) {
        // Fragments are double buffered in registers ACROSS chunk boundaries (v4): the k4 block 0 of chunk c+1 is fetched
        // during block 3 of chunk c (after checking that chunk c+1 has landed, which it has two chunks ahead of need), so a
        // warp never sits between the barrier wait and its first LDS with an empty tensor pipe behind it.
        double fa[2][4], fb[2][8];
        if (total > 0) {
            mbar_wait(pipe.bars + 8 * (g0 % STAGES), (g0 / STAGES) & 1);
            const double* sA = smem + (g0 % STAGES) * STAGE_DOUBLES;
            const double* sB = SAME_AB ? sA : sA + TILE * SLD;
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) fa[0][mi] = sA[aoff + mi * 8 * SLD];
#pragma unroll
            for (int ni = 0; ni < 8; ++ni) fb[0][ni] = sB[boff + ni * 8 * SLD];
        }
        for (int c = 0; c < total; ++c) {
            const unsigned gc = g0 + c;
            const unsigned st = gc % STAGES;
            const double* sA = smem + st * STAGE_DOUBLES;
            const double* sB = SAME_AB ? sA : sA + TILE * SLD;
            hook.pre(c);
#pragma unroll
            for (int kk = 0; kk < KC / 4; ++kk) {
                const int cur = kk & 1, nxt = cur ^ 1;
                if (kk + 1 < KC / 4) {
#pragma unroll
                    for (int mi = 0; mi < 4; ++mi) fa[nxt][mi] = sA[aoff + mi * 8 * SLD + (kk + 1) * 4];
#pragma unroll
                    for (int ni = 0; ni < 8; ++ni) fb[nxt][ni] = sB[boff + ni * 8 * SLD + (kk + 1) * 4];
                } else if (c + 1 < total) {
                    const unsigned gn = gc + 1;
                    mbar_wait(pipe.bars + 8 * (gn % STAGES), (gn / STAGES) & 1);
                    const double* nA = smem + (gn % STAGES) * STAGE_DOUBLES;
                    const double* nB = SAME_AB ? nA : nA + TILE * SLD;
#pragma unroll
                    for (int mi = 0; mi < 4; ++mi) fa[nxt][mi] = nA[aoff + mi * 8 * SLD];
#pragma unroll
                    for (int ni = 0; ni < 8; ++ni) fb[nxt][ni] = nB[boff + ni * 8 * SLD];
                }
                issue_slice(c + PREFETCH, kk);
#pragma unroll
                for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                    for (int ni = 0; ni < 8; ++ni) dmma(acc[mi][ni], fa[cur][mi], fb[cur][ni]);
            }
            hook.post(sA, c);
            __syncwarp();
            if (lane == 0) mbar_arrive(pipe.bars + 8 * (STAGES + st));
        }
    } else {
        for (int c = 0; c < total; ++c) {
            const unsigned gc = g0 + c;
            const unsigned st = gc % STAGES;
            mbar_wait(pipe.bars + 8 * st, (gc / STAGES) & 1);
            const double* sA = smem + st * STAGE_DOUBLES;
            const double* sB = SAME_AB ? sA : sA + TILE * SLD;
            hook.pre(c);
            double fa[2][4], fb[2][8];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) fa[0][mi] = sA[aoff + mi * 8 * SLD];
#pragma unroll
            for (int ni = 0; ni < 8; ++ni) fb[0][ni] = sB[boff + ni * 8 * SLD];
#pragma unroll
            for (int kk = 0; kk < KC / 4; ++kk) {
                const int cur = kk & 1, nxt = cur ^ 1;
                if (kk + 1 < KC / 4) {
#pragma unroll
                    for (int mi = 0; mi < 4; ++mi) fa[nxt][mi] = sA[aoff + mi * 8 * SLD + (kk + 1) * 4];
#pragma unroll
                    for (int ni = 0; ni < 8; ++ni) fb[nxt][ni] = sB[boff + ni * 8 * SLD + (kk + 1) * 4];
                }
                issue_slice(c + PREFETCH, kk);
#pragma unroll
                for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                    for (int ni = 0; ni < 8; ++ni) dmma(acc[mi][ni], fa[cur][mi], fb[cur][ni]);
            }
            hook.post(sA, c);
            __syncwarp();
            if (lane == 0) mbar_arrive(pipe.bars + 8 * (STAGES + st));
        }
    }
    pipe.g = g0 + total;
    __syncthreads();
}

// Row / column of accumulator element (mi, ni, e) inside the 128x128 tile.
__device__ __forceinline__ int acc_row(int mi) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    return (warp & 3) * 32 + mi * 8 + (lane >> 2);
}
__device__ __forceinline__ int acc_col(int ni, int e) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    return (warp >> 2) * 64 + ni * 8 + 2 * (lane & 3) + e;
}

// smem row stride for a [128 x dpad] scaled-input tile: stride mod 16 in {4, 12} keeps fragment loads conflict-free
__host__ __device__ __forceinline__ int feat_sld(int dpad) { return (dpad % 8 == 0) ? dpad + 4 : dpad; }

// Load a [128 x dpad] tile of scaled inputs (global row stride dpad) into smem with stride sld (cp.async; caller waits).
__device__ __forceinline__ void load_feat_tile(double* s, int sld, const double* g, int dpad) {
    const int per_row = dpad / 2;  // 16-byte chunks per row
    for (int id = threadIdx.x; id < TILE * per_row; id += NTHREADS) {
        const int row = id / per_row, cc = (id % per_row) * 2;
        cp_async16(s + row * sld + cc, g + (long long)row * dpad + cc);
    }
}

// exp(-d) for the squared-exponential epilogue, d >= 0 up to rounding (covar.py:147-149: mul_(-1).exp_()).
// libdevice's exp() costs ~65 SASS instructions per value with its range handling; 64 values per thread per 128x128 tile
// made the K / K* tile generation ~5 % of the prediction kernel.  Here: n = round(-d log2 e) by the 1.5*2^52 shifter,
// r = -d - n ln2 (two-piece ln2), degree-12 Taylor on |r| <= ln2/2 (truncation 1.7e-16), 2^n added into the exponent
// field.  Relative error < 4e-16 over the whole range (tests/test_gpu_engine.py::test_se_kernels holds K to 1e-13
// against the CPU oracle); arguments above 690 (results below 2^-995) are flushed to 0.
__device__ __forceinline__ double exp_neg(double d) {
    const double t = d * -1.4426950408889634074;
    const double shifter = 6755399441055744.0;
    const double tn = t + shifter;
    const int n = __double2loint(tn);
    const double nf = tn - shifter;
    double r = fma(nf, -6.93147180369123816490e-01, -d);
    r = fma(nf, -1.90821492927058770002e-10, r);
    double p = 2.08767569878680989792e-09;           // 1/12!
    p = fma(p, r, 2.50521083854417187751e-08);       // 1/11!
    p = fma(p, r, 2.75573192239858906526e-07);       // 1/10!
    p = fma(p, r, 2.75573192239858906526e-06);       // 1/9!
    p = fma(p, r, 2.48015873015873015873e-05);       // 1/8!
    p = fma(p, r, 1.98412698412698412698e-04);       // 1/7!
    p = fma(p, r, 1.38888888888888888889e-03);       // 1/6!
    p = fma(p, r, 8.33333333333333333333e-03);       // 1/5!
    p = fma(p, r, 4.16666666666666666667e-02);       // 1/4!
    p = fma(p, r, 1.66666666666666666667e-01);       // 1/3!
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const double v = __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
    return d > 690.0 ? 0.0 : v;   // 2^-995: far below anything a kernel matrix entry can contribute; also guards the int range of n
}

// Squared-exponential tile in GEMM form, covar.py:102-149:  sqd = (|aR|^2 + |aC|^2) - 2 aR.aC ;  k = sig2 exp(-sqd)
// The cross term runs on DMMA.  f(mi, ni, e, kval) is called for every accumulator element of the thread.
// RAW: hand the squared distance itself to f (no exponential).
template <bool RAW = false, typename F>
__device__ __forceinline__ void gen_se_tile(const double* sR, const double* sC, int sld, const double* nR, const double* nC,
                                            int dpad, double sig2, F&& f) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp & 3, wn = warp >> 2;
    const double* pa = sR + (wm * 32 + g) * sld + t;
    double nr[4];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) nr[mi] = nR[wm * 32 + mi * 8 + g];
#pragma unroll
    for (int ni = 0; ni < 8; ++ni) {
        double G[4][2];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) G[mi][0] = G[mi][1] = 0.0;
        const double* pb = sC + (wn * 64 + ni * 8 + g) * sld + t;
        for (int k = 0; k < dpad; k += 4) {
            const double b = pb[k];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) dmma(G[mi], pa[mi * 8 * sld + k], b);
        }
        const int c0 = wn * 64 + ni * 8 + 2 * t;
        const double nc0 = nC[c0], nc1 = nC[c0 + 1];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) {
            const double d0 = -2.0 * G[mi][0] + (nr[mi] + nc0), d1 = -2.0 * G[mi][1] + (nr[mi] + nc1);
            f(mi, ni, 0, RAW ? d0 : sig2 * exp_neg(d0));
            f(mi, ni, 1, RAW ? d1 : sig2 * exp_neg(d1));
        }
    }
}

// The same tile, produced INTO the accumulators (which must be zero): the cross term runs like one mainloop chunk — 12 LDS.64 +
// 32 independent DMMA per k4 step instead of gen_se_tile's per-column-block chains of 4 dependent DMMAs behind their loads — and
// every element is then replaced by f(mi, ni, e, k).  Used to seed a product  acc = -K* + sum V L^T  (predict_kernel).
template <typename F>
__device__ __forceinline__ void se_seed_acc(double (&acc)[4][8][2], const double* sR, const double* sC, int sld, const double* nR,
                                            const double* nC, int dpad, double sig2, F&& f) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp & 3, wn = warp >> 2;
    const double* pa = sR + (wm * 32 + g) * sld + t;
    const double* pb = sC + (wn * 64 + g) * sld + t;
    for (int k = 0; k < dpad; k += 4) {
        double fa[4], fb[8];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) fa[mi] = pa[mi * 8 * sld + k];
#pragma unroll
        for (int ni = 0; ni < 8; ++ni) fb[ni] = pb[ni * 8 * sld + k];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < 8; ++ni) dmma(acc[mi][ni], fa[mi], fb[ni]);
    }
    double nr[4];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) nr[mi] = nR[wm * 32 + mi * 8 + g];
#pragma unroll
    for (int ni = 0; ni < 8; ++ni) {
        const int c0 = wn * 64 + ni * 8 + 2 * t;
        const double nc0 = nC[c0], nc1 = nC[c0 + 1];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) {
            const double d0 = -2.0 * acc[mi][ni][0] + (nr[mi] + nc0), d1 = -2.0 * acc[mi][ni][1] + (nr[mi] + nc1);
            acc[mi][ni][0] = f(mi, ni, 0, sig2 * exp_neg(d0));
            acc[mi][ni][1] = f(mi, ni, 1, sig2 * exp_neg(d1));
        }
    }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------------------------------------
// Scratch behind the 128x128 diagonal tile in shared memory (the tile routines of gpr_diag_tile.cuh live in it; sized for the
// block-major exchange buffer of the round-1 routines, 136 blocks of 8x8 at 66 doubles, kept as the budget).
// ---------------------------------------------------------------------------------------------
constexpr int BM_DOUBLES = 136 * 66;

// ---------------------------------------------------------------------------------------------
// Split-K with a last-arriver fix-up, for launches that have fewer tiles than SMs (few experts, late block columns).
// Each of the S CTAs of a tile multiplies a slice of the k-range, parks its 128x128 partial in a slot, and bumps the
// tile's counter; the CTA that observes S-1 sums all S slots IN PART ORDER (bit-reproducible whichever CTA it is) and
// carries on with the tile's epilogue.  No CTA ever waits for another one.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool splitk_combine(double (&acc)[4][8][2], double* slots, int part, int S, int* counter, int* s_flag) {
    double* mine = slots + (long long)part * TILE * TILE;
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 8; ++ni)
            __stcg(reinterpret_cast<double2*>(mine + acc_row(mi) * TILE + acc_col(ni, 0)), make_double2(acc[mi][ni][0], acc[mi][ni][1]));
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int old = atomicAdd(counter, 1);
        *s_flag = (old == S - 1);
        if (old == S - 1) *counter = 0;   // ready for the next launch
    }
    __syncthreads();
    if (!*s_flag) return false;
    __threadfence();
    zero_acc(acc);
    for (int p = 0; p < S; ++p) {
        const double* sp = slots + (long long)p * TILE * TILE;
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < 8; ++ni) {
                const double2 v = __ldcg(reinterpret_cast<const double2*>(sp + acc_row(mi) * TILE + acc_col(ni, 0)));
                acc[mi][ni][0] += v.x;
                acc[mi][ni][1] += v.y;
            }
    }
    return true;
}

}  // namespace gpr
