// pygpr_b200 — FP64-accurate GEMM  C = A * B^T  on the INT8 tcgen05 tensor cores ("emulated FP64").
//
// The FP64 products of the path (the long `K* L^-T` product of the prediction, gpr.py:100, and the trailing updates of the
// Cholesky, gpr.py:69) run on DMMA at <= 37 TFLOP/s on B200 — FP64 has no tcgen05 kind.  This file is the way past that
// ceiling: an Ozaki-style splitting of both operands into signed 8-bit digit planes, all digit-plane products on
// `tcgen05.mma.cta_group::1.kind::i8` (exact int32 accumulation in TMEM), and an FP64 recombination in the epilogue.
//
//   a_ik = 2^ea_i * sum_t qa_t[i,k] 256^-(t+1),  qa_t in [-128,127]   (row-wise exponent, balanced radix-256 digits)
//   C_ij = 2^(ea_i+eb_j-16) * sum_g 256^-g * D_g[i,j],   D_g = sum_{t+u=g} qa_t qb_u^T   (int32, exact)
//
// Only the diagonals g < slices are formed (slices*(slices+1)/2 plane products); the dropped ones are below the last kept
// digit.  `slices` is the accuracy knob: 6 -> ~1e-14 of the row-norm product (FP64 grade), 3-4 -> FP32 grade at 6-10 plane
// products instead of 21 — that is the opt-in reduced-precision mode of north_star.
//
// Kernel structure (one 128 x 64 output tile per CTA, warp-specialised, B200-native):
//   warp 0   TMA producer: one 3-D `cp.async.bulk.tensor` per operand and k-block brings ALL digit planes of the tile
//            ([slices][rows][64 B], SWIZZLE_64B) into a multi-stage shared-memory ring (mbarrier full/empty);
//   warp 1   allocates TMEM (slices x 64 columns of int32 accumulators, one per diagonal) and issues the UMMA stream;
//            `tcgen05.commit` releases ring slots / signals the epilogue;
//   warps 2-9 epilogue (two per TMEM lane quadrant, 32 columns each): `tcgen05.ld` 32x32b, Horner recombination of the
//            diagonals in FP64, power-of-two row/column scales, store.
#include <cuda.h>   // CUtensorMap + enums only: the encoder is fetched with cudaGetDriverEntryPoint, libcuda is not linked
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/pygpr_b200.h"

extern "C" int gpr_internal_fail(int code, const char* fmt, const char* a, long long v);   // gpr_kernels.cu
int g_ozaki_bk = 0;   // developer A/B knob (gpr_set_option "ozaki_bk"): 0 = automatic, 32 / 64 = force the k-block

namespace oz {

constexpr int BM = 128;           // output rows per CTA  = TMEM lanes = UMMA M
constexpr int BN = 64;            // output columns per CTA = UMMA N; slices * BN <= 512 TMEM columns
constexpr int BK = 64;            // int8 per k-block = one SWIZZLE_64B row; two UMMA k-steps of 32 (planes are padded to this)
constexpr int BK_SMALL = 32;      // alternative k-block (A/B knob "ozaki_bk"): one SWIZZLE_32B row, one UMMA k-step — half the stage,
                                  // twice the ring depth; measured slower, see gpr_ozaki_gemm_nt
constexpr int MAX_SLICES = 8;
constexpr int NTHREADS = 320;      // warp 0 TMA, warp 1 UMMA, warps 2..9 epilogue (two per TMEM lane quadrant)
constexpr int SMEM_LIMIT = 227 * 1024;

__host__ __device__ inline int stage_bytes(int slices, int bk) { return slices * (BM + BN) * bk; }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done, spins = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        // a wedged pipeline (a TMA that never lands) must fail loudly instead of hanging the device: try_wait sleeps in
        // hardware between polls, so this bound is seconds, far beyond any legitimate wait of this kernel
        if (!done && ++spins > (1u << 24)) __trap();
    } while (!done);
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        :
        : "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, int8 x int8 -> int32, issued by ONE thread for the CTA.  Descriptors are passed as their
// low words (the high words are compile-time constants) so that the per-product address arithmetic is one 32-bit add.
template <int ACC>
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], da, db, %4, p;\n\t}"
        :
        : "r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "n"(ACC)
        : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred)::"memory");
    return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {   // arrives on `bar` when every UMMA issued so far has finished
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 16 consecutive int32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t addr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(addr)
        : "memory");
}
// the wait carries the registers so that no use of them can be scheduled above it
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&v)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]),
                   "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
                 :
                 : "memory");
}

// exact int32 -> double without the conversion pipe: the biased integer is dropped into the mantissa of 2^52 and the bias
// (2^52 + 2^31) subtracted — one logic op and one DADD instead of I2F.F64
__device__ __forceinline__ double i32_to_f64(uint32_t v) {
    return __hiloint2double(0x43300000, (int)(v ^ 0x80000000u)) - 4503601774854144.0;
}

// K-major operand tile [rows][64 B] written by TMA with SWIZZLE_64B: 8-row groups 512 B apart; LBO unused for swizzled
// K-major layouts (set to 1 like CUTLASS); descriptor version 1 (Blackwell); layout type 4 = SWIZZLE_64B.
// (SWIZZLE_32B rows are 32 B: groups 256 B apart, layout type 6.)
template <int BKT>
struct Desc {
    static constexpr uint64_t HI = (1ull << 16) | (((uint64_t)(8 * BKT) >> 4) << 32) | (1ull << 46) | ((BKT == 64 ? 4ull : 6ull) << 61);
    static constexpr uint32_t HI32 = (uint32_t)(HI >> 32);
};
constexpr uint32_t DESC_LO_FIXED = 1u << 16;                                  // LBO field
__device__ __forceinline__ uint32_t desc_lo(uint32_t addr) { return DESC_LO_FIXED | ((addr & 0x3FFFF) >> 4); }

// instruction descriptor: D = S32 (2 @bit 4), A = B = signed 8 bit (1 @bit 7, 1 @bit 10), both K-major, N>>3 @bit 17, M>>4 @bit 24
__host__ __device__ constexpr uint32_t idesc_for(int n) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

// All plane products of one k-block: D_g += A_t B_u^T for t + u = g < S, two k-steps of 32.
// The accumulators of the diagonals sit side by side in TMEM (diagonal g at columns [g BN, (g+1) BN)) and the digit planes of
// the B tile sit one after the other in shared memory, so for a fixed A plane t the products with B planes u0, u0+1, ... land in
// CONSECUTIVE accumulators: they are issued as ONE UMMA with N = planes x BN <= 256 whose B operand is the stack of those
// planes.  A 128 x N x 32 i8 UMMA fed from shared memory costs max(N/2, (128+N)/4) clocks (tools/microbench/umma_i8_rate.cu:
// N=64 48, N=128 64, N=256 128): fused, the A plane is read once per 256 columns instead of once per 64 and the stream runs at
// ~95 % of the tensor pipe instead of 67 %.  Fully unrolled: every offset and descriptor field is an immediate.
template <int S, bool FIRST, int BKT>
__device__ __forceinline__ void issue_k_block(uint32_t tmem, uint32_t a_lo, uint32_t b_lo) {
    constexpr int PL = 256 / BN;                       // B planes per UMMA
    constexpr int A_PLANE = BM * BKT, B_PLANE = BN * BKT;
    constexpr uint32_t DESC_HI32 = Desc<BKT>::HI32;
#pragma unroll
    for (int kk = 0; kk < BKT / 32; ++kk)
#pragma unroll
        for (int t = 0; t < S; ++t)
#pragma unroll
            for (int u0 = 0; u0 < S - t; u0 += PL) {
                const int planes = (S - t - u0) < PL ? (S - t - u0) : PL;
                const uint32_t al = a_lo + (uint32_t)((t * A_PLANE + kk * 32) >> 4);
                const uint32_t bl = b_lo + (uint32_t)((u0 * B_PLANE + kk * 32) >> 4);
                const uint32_t d = tmem + (uint32_t)((t + u0) * BN);
                // the t = 0 products of the very first k-step cover every accumulator column once: they overwrite
                if (FIRST && kk == 0 && t == 0) umma_i8<0>(d, al, bl, DESC_HI32, idesc_for(planes * BN));
                else umma_i8<1>(d, al, bl, DESC_HI32, idesc_for(planes * BN));
            }
}

// ---------------------------------------------------------------------------------------------
// digit planes:  X [rows, k] (FP64, leading dimension ld)  ->  planes [slices][rpad][kpad] int8, scale [rpad]
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) split_kernel(const double* __restrict__ X, long long ld, int rows, int k, int rpad, int kpad,
                                                    int slices, double bound, int8_t* __restrict__ planes, double* __restrict__ scale,
                                                    long long x_batch_stride) {
    // batch entry blockIdx.y: its matrix starts x_batch_stride doubles further, its planes / scales rpad rows further
    const int r = blockIdx.x;
    X += (long long)blockIdx.y * x_batch_stride;
    scale += (long long)blockIdx.y * rpad;
    const long long plane_stride = (long long)gridDim.y * rpad * kpad;
    int8_t* out = planes + ((long long)blockIdx.y * rpad + r) * kpad;
    __shared__ double red[8];
    __shared__ int e_sh;
    // row scale from the row maximum, or — when the caller knows a bound on |x| (bound > 0, e.g. sig^2 for an SE kernel matrix) —
    // from that bound for every row: one pass over the data instead of two
    double amax = bound > 0.0 ? bound : 0.0;
    if (r < rows && !(bound > 0.0)) {
        const double* x = X + (long long)r * ld;
        for (int c = threadIdx.x; c < k; c += blockDim.x) amax = fmax(amax, fabs(x[c]));
    }
    for (int o = 16; o; o >>= 1) amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = amax;
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = red[0];
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) a = fmax(a, red[w]);
        // |x| 2^-e < 1/4: the leading digit stays within [-64, 64] and the carries of the balanced digits cannot overflow it
        const int e = (a > 0.0 && isfinite(a)) ? ilogb(a) + 3 : 0;
        e_sh = e;
        scale[r] = (r < rows && a > 0.0) ? ldexp(1.0, e - 8) : 0.0;
    }
    __syncthreads();
    const int e = e_sh;
    const double mult = ldexp(1.0, 8 * slices - e);     // x * mult = the row's fixed-point integer with 8*slices fraction bits
    for (int c4 = threadIdx.x * 4; c4 < kpad; c4 += blockDim.x * 4) {
        uint32_t packed[MAX_SLICES];
#pragma unroll
        for (int t = 0; t < MAX_SLICES; ++t) packed[t] = 0u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = c4 + j;
            const double x = (r < rows && c < k) ? X[(long long)r * ld + c] : 0.0;
            long long I = __double2ll_rn(x * mult);
            // balanced radix-256 digits, least significant first: exact integer arithmetic, digits in [-128, 127]
#pragma unroll
            for (int t = MAX_SLICES - 1; t >= 0; --t) {
                if (t < slices) {
                    const int8_t dgt = (int8_t)(I & 0xFF);
                    I = (I - (long long)dgt) >> 8;
                    packed[t] |= ((uint32_t)(uint8_t)dgt) << (8 * j);
                }
            }
        }
#pragma unroll
        for (int t = 0; t < MAX_SLICES; ++t)
            if (t < slices) *reinterpret_cast<uint32_t*>(out + (long long)t * plane_stride + c4) = packed[t];
    }
}

// ---------------------------------------------------------------------------------------------
// the GEMM.  One CTA per output tile on purpose: a persistent variant (one CTA per SM looping over a static tile list, TMEM / barriers
// set up once, the next tile's first TMA loads issued under the epilogue) was built and measured on the same box — +3-4 % on long
// dense products, but every product of the path is triangular (b_lower / a_upper / c_lower: k-ranges differ from tile to tile) and
// there the hardware's block scheduler balances better than a static list, while the restructured kernel cost the short-k panel
// products of the factorisation 12 % (80 -> 91 ms per 32 experts).  Not kept.
// ---------------------------------------------------------------------------------------------
template <int S, int BKT>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
            const double* __restrict__ scaleA, const double* __restrict__ scaleB, double* __restrict__ C, long long ldc,
            int M, int N, int kpad, int stages, int b_lower, double beta, int group_m, int tmem_cols,
            int a_batch_rows, int b_batch_rows, long long c_batch_stride, int a_upper, int c_lower, double alpha, int a_off, int b_off) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;             // swizzled tiles want (at least) 512 B alignment
    constexpr int A_PLANE = BM * BKT;
    constexpr int sbytes = S * (BM + BN) * BKT;
    const uint32_t bar_full = base + (uint32_t)(stages * sbytes);
    const uint32_t bar_empty = bar_full + 8u * stages;
    const uint32_t bar_tmem = bar_empty + 8u * stages;
    const uint32_t slot_addr = bar_tmem + 8u;
    volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (slot_addr - raw));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // tile of this CTA: groups of `group_m` row tiles are swept column by column so that a wave of CTAs shares operand tiles in L2;
    // with a lower-triangular B the column tiles with the longest k-range go first
    const int tiles_m = (M + BM - 1) / BM, tiles_n = (N + BN - 1) / BN;
    const int per_group = group_m * tiles_n;
    const int gid = blockIdx.x / per_group, in_group = blockIdx.x % per_group;
    const int first_m = gid * group_m;
    const int gsz = min(tiles_m - first_m, group_m);
    const int tm = first_m + in_group % gsz;
    int tn = in_group / gsz;
    if (b_lower >= 0) tn = tiles_n - 1 - tn;
    if (c_lower && tn * BN > tm * BM + BM - 1) return;          // only tiles on / below the diagonal of C are wanted (whole CTA leaves)
    int nkb = kpad / BKT;
    if (b_lower >= 0) nkb = min(nkb, (b_lower + (tn + 1) * BN + BKT - 1) / BKT);   // B[j, k] = 0 for k > j + b_lower
    const int kb0 = a_upper >= 0 ? min(nkb - 1, (a_upper + tm * BM) / BKT) : 0;   // A[i, k] = 0 for k < i + a_upper (rows of an upper factor)
    // batched product: batch entry blockIdx.y owns a_batch_rows / b_batch_rows consecutive (padded) rows of the plane arrays
    // (a_off / b_off: the product uses a window of rows of every batch entry's operand, starting that many rows in)
    const int a_row0 = blockIdx.y * a_batch_rows + a_off + tm * BM, b_row0 = blockIdx.y * b_batch_rows + b_off + tn * BN;
    scaleA += (long long)blockIdx.y * a_batch_rows + a_off;
    scaleB += (long long)blockIdx.y * b_batch_rows + b_off;
    C += (long long)blockIdx.y * c_batch_stride;

    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(bar_full + 8u * s, 1);
            mbar_init(bar_empty + 8u * s, 1);
        }
        mbar_init(bar_tmem, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_addr), "r"((uint32_t)tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot_ptr;

    if (warp == 0) {
        // ---- TMA producer: the whole warp walks the ring (warp-uniform control flow), one elected lane issues ----------------
        const bool leader = elect_one();
        if (leader) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapA)) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapB)) : "memory");
        }
        int stage = 0;
        uint32_t phase = 0;
        for (int kb = kb0; kb < nkb; ++kb) {
            mbar_wait(bar_empty + 8u * stage, phase ^ 1u);
            if (leader) {
                const uint32_t full = bar_full + 8u * stage;
                const uint32_t dst = base + (uint32_t)(stage * sbytes);
                mbar_expect_tx(full, (uint32_t)sbytes);
                tma_load_3d(dst, &mapA, full, kb * BKT, a_row0, 0);
                tma_load_3d(dst + (uint32_t)(S * A_PLANE), &mapB, full, kb * BKT, b_row0, 0);
            }
            __syncwarp();
            if (++stage == stages) { stage = 0; phase ^= 1u; }
        }
    } else if (warp == 1) {
        // ---- UMMA issuer: same pattern; everything but the elected lane's instructions is uniform-datapath arithmetic ---------
        const bool leader = elect_one();
        int stage = 0;
        uint32_t phase = 0;
        for (int kb = kb0; kb < nkb; ++kb) {
            mbar_wait(bar_full + 8u * stage, phase);
            tc_fence_after();
            const uint32_t a_lo = desc_lo(base + (uint32_t)(stage * sbytes));
            const uint32_t b_lo = a_lo + (uint32_t)((S * A_PLANE) >> 4);
            if (leader) {
                if (kb == kb0) issue_k_block<S, true, BKT>(tmem, a_lo, b_lo);
                else issue_k_block<S, false, BKT>(tmem, a_lo, b_lo);
                umma_commit(bar_empty + 8u * stage);       // ring slot free once these products have read it
            }
            __syncwarp();
            if (++stage == stages) { stage = 0; phase ^= 1u; }
        }
        if (leader) umma_commit(bar_tmem);                  // accumulators complete
        __syncwarp();
    } else {
        // ---- epilogue warps 2..9: warp w may touch TMEM lanes [32 (w % 4), +32); warps w and w + 4 split the columns --------------
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;                   // which 32 of the tile's 64 columns this warp drains
        const int row = tm * BM + q * 32 + lane;
        const double sa = alpha * scaleA[row];              // scale arrays are padded to whole tiles
        mbar_wait(bar_tmem, 0);
        tc_fence_after();
        const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
        for (int c0 = half * (BN / 2); c0 < (half + 1) * (BN / 2); c0 += 16) {
            // all diagonals of these 16 columns in flight at once, ONE wait: S TMEM round trips overlap instead of serialising
            uint32_t v[S][16];
#pragma unroll
            for (int g = 0; g < S; ++g) tmem_ld16(lane_base + (uint32_t)(g * BN + c0), v[g]);
#pragma unroll
            for (int g = 0; g < S; ++g) tmem_ld_wait(v[g]);
            double acc[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = 0.0;
#pragma unroll
            for (int g = S - 1; g >= 0; --g)               // Horner in 1/256, smallest diagonal first
#pragma unroll
                for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], 0.00390625, i32_to_f64(v[g][i]));
            const int col0 = tn * BN + c0;
            if (row < M) {
                double* crow = C + (long long)row * ldc + col0;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    if (col0 + i < N) {
                        const double val = acc[i] * sa * scaleB[col0 + i];
                        crow[i] = (beta != 0.0) ? fma(beta, crow[i], val) : val;
                    }
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)tmem_cols) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*encode_fn_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static encode_fn_t encoder() {
    static encode_fn_t fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<encode_fn_t>(p);
    }
    return fn;
}

// planes [slices][rpad][kpad] int8 as a 3-D tensor (k fastest); box = all planes of a [box_rows x 64] tile
static int make_map(CUtensorMap* map, const void* planes, int rpad, int kpad, int slices, int box_rows, int bk) {
    encode_fn_t enc = encoder();
    if (!enc) return gpr_internal_fail(GPR_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver%s", "", 0);
    cuuint64_t dims[3] = {(cuuint64_t)kpad, (cuuint64_t)rpad, (cuuint64_t)slices};
    cuuint64_t strides[2] = {(cuuint64_t)kpad, (cuuint64_t)kpad * (cuuint64_t)rpad};
    cuuint32_t box[3] = {(cuuint32_t)bk, (cuuint32_t)box_rows, (cuuint32_t)slices};
    cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(planes), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, bk == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return gpr_internal_fail(GPR_ERR_CUDA, "cuTensorMapEncodeTiled failed%s (CUresult %lld)", "", (long long)r);
    return GPR_OK;
}

static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

}  // namespace oz

extern "C" long long gpr_ozaki_rows_padded(int rows) { return oz::round_up(rows, oz::BM); }
extern "C" long long gpr_ozaki_k_padded(int k) { return oz::round_up(k, oz::BK); }
extern "C" long long gpr_ozaki_planes_bytes(int rows, int k, int slices) {
    return (long long)slices * oz::round_up(rows, oz::BM) * oz::round_up(k, oz::BK);
}

extern "C" int gpr_ozaki_split(const double* X, long long ld, int rows, int k, int slices, double bound, void* planes, double* scale,
                               int batch, long long x_batch_stride, void* stream) {
    if (!(X && planes && scale && rows > 0 && k > 0 && ld >= k && slices >= 1 && slices <= oz::MAX_SLICES && batch >= 1 && batch <= 65535))
        return gpr_internal_fail(GPR_ERR_BAD_ARG, "bad argument: gpr_ozaki_split%s (slices=%lld)", "", slices);
    const int rpad = oz::round_up(rows, oz::BM), kpad = oz::round_up(k, oz::BK);
    oz::split_kernel<<<dim3(rpad, batch), 256, 0, (cudaStream_t)stream>>>(X, ld, rows, k, rpad, kpad, slices, bound,
                                                                          static_cast<int8_t*>(planes), scale, x_batch_stride);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return gpr_internal_fail(GPR_ERR_CUDA, "%s (cuda error %lld)", cudaGetErrorString(e), (long long)e);
    return GPR_OK;
}

extern "C" int gpr_ozaki_gemm_nt(const void* planesA, const double* scaleA, int M, const void* planesB, const double* scaleB, int N,
                                 int K, int slices, int b_lower, double beta, double* C, long long ldc, int batch,
                                 long long c_batch_stride, int a_upper, int c_lower, double alpha, int b_shared, void* stream) {
    return gpr_ozaki_gemm_nt_window(planesA, scaleA, M, 0, 0, planesB, scaleB, N, 0, 0, K, slices, b_lower, beta, C, ldc, batch, c_batch_stride,
                                    a_upper, c_lower, alpha, b_shared, stream);
}

extern "C" int gpr_ozaki_gemm_nt_window(const void* planesA, const double* scaleA, int M, int a_rows_total, int a_row_off, const void* planesB,
                                        const double* scaleB, int N, int b_rows_total, int b_row_off, int K, int slices, int b_lower,
                                        double beta, double* C, long long ldc, int batch, long long c_batch_stride, int a_upper,
                                        int c_lower, double alpha, int b_shared, void* stream) {
    if (!(planesA && planesB && scaleA && scaleB && C && M > 0 && N > 0 && K > 0 && ldc >= N && slices >= 1 && slices <= oz::MAX_SLICES &&
          batch >= 1 && batch <= 65535 && b_lower >= -1 && a_upper >= -1))
        return gpr_internal_fail(GPR_ERR_BAD_ARG, "bad argument: gpr_ozaki_gemm_nt%s (slices=%lld)", "", slices);
    if (a_rows_total <= 0) { a_rows_total = oz::round_up(M, oz::BM); a_row_off = 0; }
    if (b_rows_total <= 0) { b_rows_total = oz::round_up(N, oz::BM); b_row_off = 0; }
    if (!(a_rows_total % oz::BM == 0 && b_rows_total % oz::BM == 0 && a_row_off >= 0 && b_row_off >= 0 && a_row_off % oz::BN == 0 &&
          b_row_off % oz::BN == 0 && a_row_off + M <= a_rows_total && b_row_off + N <= b_rows_total))
        return gpr_internal_fail(GPR_ERR_BAD_ARG, "bad argument: gpr_ozaki_gemm_nt_window%s (row window, offset %lld)", "", a_row_off);
    // int32 headroom of a diagonal: (g+1) products of |q| <= 128 over K terms
    if ((long long)K * slices * 16384LL >= (1LL << 31))
        return gpr_internal_fail(GPR_ERR_UNSUPPORTED, "K too long for exact int32 accumulation%s (K=%lld): split the product", "", K);
    const int kpad = oz::round_up(K, oz::BK);
    const int rpadA = a_rows_total, rpadB = b_rows_total;     // rows of one batch entry in the plane arrays
    // k-block: 64 B rows.  The 32 B variant (twice the ring depth where 7-8 digit planes leave room for only two 64-wide stages)
    // is kept as an A/B knob: measured slower at every plane count (4096^3, 7 planes: 92 vs 106 TFLOP/s-equivalent; 6 planes:
    // 110 vs 142) — the per-stage handshake and the 32 B TMA rows cost more than the deeper ring hides.
    int bk = oz::BK;
    if (g_ozaki_bk == 32 || g_ozaki_bk == 64) bk = g_ozaki_bk;
    CUtensorMap mapA, mapB;
    int rc = oz::make_map(&mapA, planesA, rpadA * batch, kpad, slices, oz::BM, bk);
    if (rc != GPR_OK) return rc;
    rc = oz::make_map(&mapB, planesB, b_shared ? rpadB : rpadB * batch, kpad, slices, oz::BN, bk);
    if (rc != GPR_OK) return rc;
    const int sbytes = oz::stage_bytes(slices, bk);
    int stages = (oz::SMEM_LIMIT - 2048) / sbytes;
    if (stages > 8) stages = 8;
    if (stages < 2) return gpr_internal_fail(GPR_ERR_UNSUPPORTED, "too many digit planes for a two-stage ring%s (%lld)", "", slices);
    const int smem = stages * sbytes + 1024 + 8 * (2 * stages + 1) + 16;
    int tmem_cols = 32;
    while (tmem_cols < slices * oz::BN) tmem_cols *= 2;
    const int tiles_m = (M + oz::BM - 1) / oz::BM, tiles_n = (N + oz::BN - 1) / oz::BN;
    const int group_m = tiles_m < 16 ? tiles_m : 16;
    cudaError_t e = cudaSuccess;
#define OZ_LAUNCH_BK(S, BKT)                                                                                                   \
    e = cudaFuncSetAttribute(oz::gemm_kernel<S, BKT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);                      \
    if (e == cudaSuccess)                                                                                                      \
        oz::gemm_kernel<S, BKT><<<dim3(tiles_m * tiles_n, batch), oz::NTHREADS, smem, (cudaStream_t)stream>>>(                  \
            mapA, mapB, scaleA, scaleB, C, ldc, M, N, kpad, stages, b_lower, beta, group_m, tmem_cols, rpadA,                  \
            b_shared ? 0 : rpadB, c_batch_stride, a_upper, c_lower, alpha, a_row_off, b_row_off);
#define OZ_LAUNCH(S)                                                                                                           \
    case S:                                                                                                                    \
        if (bk == 64) { OZ_LAUNCH_BK(S, 64) } else { OZ_LAUNCH_BK(S, 32) }                                                     \
        break;
    switch (slices) {
        OZ_LAUNCH(1) OZ_LAUNCH(2) OZ_LAUNCH(3) OZ_LAUNCH(4) OZ_LAUNCH(5) OZ_LAUNCH(6) OZ_LAUNCH(7) OZ_LAUNCH(8)
    }
#undef OZ_LAUNCH_BK
#undef OZ_LAUNCH
    if (e != cudaSuccess) return gpr_internal_fail(GPR_ERR_CUDA, "%s (cuda error %lld)", cudaGetErrorString(e), (long long)e);
    e = cudaGetLastError();
    if (e != cudaSuccess) return gpr_internal_fail(GPR_ERR_CUDA, "%s (cuda error %lld)", cudaGetErrorString(e), (long long)e);
    return GPR_OK;
}
