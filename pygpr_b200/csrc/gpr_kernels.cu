// pygpr_b200 — sm_100a kernels + C ABI for the PyGPR hot path (see include/pygpr_b200.h).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -shared -Xcompiler -fPIC (pygpr_b200/csrc/build.py)
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "../../include/pygpr_b200.h"
#include "gpr_core.cuh"
#include "gpr_diag_tile.cuh"

using namespace gpr;

// Cross-chunk fragment prefetch of the mainloop, chosen per kernel family from same-box A/B runs (tools/_ab.sh):
// it helps the long single-product loops of the prediction (+1.5 %) and costs the factorisation ~1 %.
#ifndef GPR_XPF_MASK
#define GPR_XPF_MASK 2
#endif
constexpr bool XPF_POTRF = (GPR_XPF_MASK & 1) != 0;
constexpr bool XPF_PREDICT = (GPR_XPF_MASK & 2) != 0;
constexpr bool XPF_TRTRI = (GPR_XPF_MASK & 4) != 0;
constexpr bool XPF_LAUUM = (GPR_XPF_MASK & 8) != 0;

// ---------------------------------------------------------------------------------------------
// error plumbing (no exceptions cross the C boundary)
// ---------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, const char* a = "", long long v = 0) {
    snprintf(g_err, sizeof(g_err), fmt, a, v);
    return code;
}
#define CUDA_TRY(expr)                                                                                  \
    do {                                                                                                \
        cudaError_t e__ = (expr);                                                                       \
        if (e__ != cudaSuccess) return fail(GPR_ERR_CUDA, "%s (cuda error %lld)", cudaGetErrorString(e__), (long long)e__); \
    } while (0)
#define ARG_CHECK(cond)                                                                 \
    do {                                                                                \
        if (!(cond)) return fail(GPR_ERR_BAD_ARG, "bad argument: %s (line %lld)", #cond, __LINE__); \
    } while (0)

// Schedule overrides (developer A/B knobs).  Read from the environment ONCE, when the library is loaded
// (PYGPR_POTRF_MODE=auto|columns|columns1|dataflow, PYGPR_TRSV_MODE=auto|single); gpr_set_option changes them at run time.
enum { POTRF_AUTO = 0, POTRF_COLUMNS = 1, POTRF_COLUMNS1 = 2, POTRF_DATAFLOW = 3 };
enum { TRSV_AUTO = 0, TRSV_SINGLE = 1 };
static int parse_potrf_mode(const char* v) {
    if (!v) return POTRF_AUTO;
    if (!strcmp(v, "columns1")) return POTRF_COLUMNS1;
    if (!strncmp(v, "columns", 7)) return POTRF_COLUMNS;
    if (!strcmp(v, "dataflow")) return POTRF_DATAFLOW;
    return POTRF_AUTO;
}
static int parse_trsv_mode(const char* v) { return (v && !strcmp(v, "single")) ? TRSV_SINGLE : TRSV_AUTO; }
extern int g_ozaki_bk;   // ozaki_gemm.cu
static int g_potrf_mode = parse_potrf_mode(getenv("PYGPR_POTRF_MODE"));
static int g_trsv_mode = parse_trsv_mode(getenv("PYGPR_TRSV_MODE"));

extern "C" int gpr_set_option(const char* name, const char* value) {
    if (name && !strcmp(name, "potrf_mode")) { g_potrf_mode = parse_potrf_mode(value); return GPR_OK; }
    if (name && !strcmp(name, "trsv_mode")) { g_trsv_mode = parse_trsv_mode(value); return GPR_OK; }
    if (name && !strcmp(name, "ozaki_bk")) { g_ozaki_bk = value ? atoi(value) : 0; return GPR_OK; }
    return fail(GPR_ERR_BAD_ARG, "unknown option %s", name ? name : "(null)");
}

// shared with the other translation units of the library (ozaki_gemm.cu)
extern "C" int gpr_internal_fail(int code, const char* fmt, const char* a, long long v) { return fail(code, fmt, a, v); }

extern "C" int gpr_abi_version(void) { return GPR_ABI_VERSION; }
extern "C" const char* gpr_last_error(void) { return g_err; }
extern "C" int gpr_device_sm_count(void) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    return n;
}

template <typename K>
static cudaError_t opt_in_smem(K kernel, int bytes) {
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

// ---------------------------------------------------------------------------------------------
// K0: ARD scaling  a = x * ls (zero padded), nrm = |a|^2            covar.py:137-140, 106
// HBM-bound elementwise; one thread per row.
// ---------------------------------------------------------------------------------------------
__global__ void se_scale_kernel(const double* __restrict__ x, long long x_bstride, const double* __restrict__ hp,
                                int hp_stride, int n, int d, int npad, int dpad, double* __restrict__ a,
                                double* __restrict__ nrm) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npad) return;
    const double* ls = hp + (long long)b * hp_stride + 1;
    double* ar = a + ((long long)b * npad + i) * dpad;
    double s = 0.0;
    if (i < n) {
        const double* xr = x + (long long)b * x_bstride + (long long)i * d;
        for (int k = 0; k < d; ++k) {
            const double v = xr[k] * ls[k];
            ar[k] = v;
            s += v * v;  // same left-to-right order as tc.sum(x.square(), 2) on short rows
        }
        for (int k = d; k < dpad; ++k) ar[k] = 0.0;
    } else {
        for (int k = 0; k < dpad; ++k) ar[k] = 0.0;
    }
    nrm[(long long)b * npad + i] = s;
}

extern "C" int gpr_se_scale(const double* x, long long x_bstride, const double* hp, int hp_stride, int B, int n, int d,
                            int npad, int dpad, double* a, double* nrm, void* stream) {
    ARG_CHECK(x && hp && a && nrm);
    ARG_CHECK(B > 0 && n > 0 && d > 0 && npad >= n && dpad >= d && dpad % 4 == 0);
    dim3 grid((npad + 127) / 128, B);
    se_scale_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(x, x_bstride, hp, hp_stride, n, d, npad, dpad, a, nrm);
    CUDA_TRY(cudaGetLastError());
    return GPR_OK;
}

// ---------------------------------------------------------------------------------------------
// K1/K2: stand-alone kernel-matrix builders (API parity of Covar.kernel; the fit/predict path
// generates these tiles inside the factorisation / solve kernels instead).
// ---------------------------------------------------------------------------------------------
struct FeatSmem {
    double* sR;
    double* sC;
    double* nR;
    double* nC;
    int sld;
};
__device__ __forceinline__ FeatSmem carve_feat(double* base, int dpad) {
    FeatSmem f;
    f.sld = feat_sld(dpad);
    f.sR = base;
    f.sC = base + TILE * f.sld;
    f.nR = f.sC + TILE * f.sld;
    f.nC = f.nR + TILE;
    return f;
}
__host__ __device__ static inline int feat_smem_bytes(int dpad) { return (2 * TILE * feat_sld(dpad) + 2 * TILE) * 8; }
// dynamic shared memory of a kernel that factorises diagonal tiles: the pipeline stages, overlaid by the 128x128 tile, and
// behind the tile the larger of the feature tiles (GEN) and the scratch of potf2_tile / trtri_tile (gpr_diag_tile.cuh)
static inline int diag_smem_bytes(int gen, int dpad) {
    const int behind = (gen && feat_smem_bytes(dpad) > BM_DOUBLES * 8) ? feat_smem_bytes(dpad) : BM_DOUBLES * 8;
    const int need = TILE * XLD * 8 + behind;
    return need > MAIN_SMEM_BYTES ? need : MAIN_SMEM_BYTES;
}

// rows from (aR, nrmR) tile tr, cols from (aC, nrmC) tile tc; out[b][row][col] with leading dim ld.
//   mode 0 (symmetric build): diag_add on row==col, identity padding up to `ldlim` when pad_identity
//   mode 1 (cross): plain
//   mode 2 (squared distance): |aR|^2 + |aC|^2 - 2 aR.aC without the exponential (sampler.py:94-100)
__global__ void __launch_bounds__(NTHREADS) se_tile_kernel(const double* __restrict__ aR, const double* __restrict__ nrmR,
                                                           int rpad, const double* __restrict__ aC,
                                                           const double* __restrict__ nrmC, int cpad,
                                                           const double* __restrict__ hp, int hp_stride,
                                                           const double* __restrict__ diag_add, int nrows, int ncols,
                                                           int dpad, double* __restrict__ out, long long ld,
                                                           long long out_bstride, int mode, int pad_identity,
                                                           int lower_only) {
    extern __shared__ double smem[];
    const int tc = blockIdx.x, tr = blockIdx.y, b = blockIdx.z;
    if (lower_only && tc > tr) return;
    FeatSmem f = carve_feat(smem, dpad);
    load_feat_tile(f.sR, f.sld, aR + ((long long)b * rpad + tr * TILE) * dpad, dpad);
    load_feat_tile(f.sC, f.sld, aC + ((long long)b * cpad + tc * TILE) * dpad, dpad);
    cp_async_commit();
    if (threadIdx.x < TILE) {
        f.nR[threadIdx.x] = nrmR[(long long)b * rpad + tr * TILE + threadIdx.x];
        f.nC[threadIdx.x] = nrmC[(long long)b * cpad + tc * TILE + threadIdx.x];
    }
    cp_async_wait<0>();
    __syncthreads();
    const double sig = hp[(long long)b * hp_stride];
    const double sig2 = sig * sig;
    const double dadd = (mode == 0 && diag_add) ? diag_add[b] : 0.0;
    double* ob = out + (long long)b * out_bstride;
    const long long lim_r = (mode == 0 && pad_identity) ? ld : nrows;
    const long long lim_c = (mode == 0 && pad_identity) ? ld : ncols;
    auto store = [&](int mi, int ni, int e, double kv) {
        const long long r = (long long)tr * TILE + acc_row(mi);
        const long long c = (long long)tc * TILE + acc_col(ni, e);
        if (r < nrows && c < ncols) {
            ob[r * ld + c] = (mode == 0 && r == c) ? kv + dadd : kv;
        } else if (r < lim_r && c < lim_c) {
            ob[r * ld + c] = (r == c) ? 1.0 : 0.0;
        }
    };
    if (mode == 2) gen_se_tile<true>(f.sR, f.sC, f.sld, f.nR, f.nC, dpad, sig2, store);
    else gen_se_tile<false>(f.sR, f.sC, f.sld, f.nR, f.nC, dpad, sig2, store);
}

// The same tile for WIDE inputs (dpad beyond what two [128 x dpad] feature tiles can occupy in shared memory, d >= 109): the
// cross term is accumulated over slices of SE_SLICE features — load a slice of both feature tiles, 4 x 8 DMMA blocks per k4 step
// into the thread's accumulators, next slice — and the epilogue is the one of se_tile_kernel.  The reference has no limit on d
// (covar.py:102-127 is a plain matmul); neither has the kernel-level API now.
constexpr int SE_SLICE = 64;
__global__ void __launch_bounds__(NTHREADS) se_tile_wide_kernel(const double* __restrict__ aR, const double* __restrict__ nrmR,
                                                                int rpad, const double* __restrict__ aC,
                                                                const double* __restrict__ nrmC, int cpad,
                                                                const double* __restrict__ hp, int hp_stride,
                                                                const double* __restrict__ diag_add, int nrows, int ncols,
                                                                int dpad, double* __restrict__ out, long long ld,
                                                                long long out_bstride, int mode, int pad_identity,
                                                                int lower_only) {
    extern __shared__ double smem[];
    const int tc = blockIdx.x, tr = blockIdx.y, b = blockIdx.z;
    if (lower_only && tc > tr) return;
    FeatSmem f = carve_feat(smem, SE_SLICE);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp & 3, wn = warp >> 2;
    if (threadIdx.x < TILE) {
        f.nR[threadIdx.x] = nrmR[(long long)b * rpad + tr * TILE + threadIdx.x];
        f.nC[threadIdx.x] = nrmC[(long long)b * cpad + tc * TILE + threadIdx.x];
    }
    double acc[4][8][2];
    zero_acc(acc);
    const double* gR = aR + ((long long)b * rpad + tr * TILE) * dpad;
    const double* gC = aC + ((long long)b * cpad + tc * TILE) * dpad;
    for (int k0 = 0; k0 < dpad; k0 += SE_SLICE) {
        const int ds = (dpad - k0 < SE_SLICE) ? dpad - k0 : SE_SLICE;   // a multiple of 4 (dpad is)
        __syncthreads();                                                 // the previous slice has been consumed
        const int per_row = ds / 2;
        for (int id = threadIdx.x; id < TILE * per_row; id += NTHREADS) {
            const int row = id / per_row, cc = (id % per_row) * 2;
            cp_async16(f.sR + row * f.sld + cc, gR + (long long)row * dpad + k0 + cc);
            cp_async16(f.sC + row * f.sld + cc, gC + (long long)row * dpad + k0 + cc);
        }
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
        const double* pa = f.sR + (wm * 32 + g) * f.sld + t;
        const double* pb = f.sC + (wn * 64 + g) * f.sld + t;
        for (int k = 0; k < ds; k += 4) {
            double fa[4], fb[8];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) fa[mi] = pa[mi * 8 * f.sld + k];
#pragma unroll
            for (int ni = 0; ni < 8; ++ni) fb[ni] = pb[ni * 8 * f.sld + k];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 8; ++ni) dmma(acc[mi][ni], fa[mi], fb[ni]);
        }
    }
    const double sig = hp[(long long)b * hp_stride];
    const double sig2 = sig * sig;
    const double dadd = (mode == 0 && diag_add) ? diag_add[b] : 0.0;
    double* ob = out + (long long)b * out_bstride;
    const long long lim_r = (mode == 0 && pad_identity) ? ld : nrows;
    const long long lim_c = (mode == 0 && pad_identity) ? ld : ncols;
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
        const double nr = f.nR[wm * 32 + mi * 8 + g];
#pragma unroll
        for (int ni = 0; ni < 8; ++ni)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const double d2 = -2.0 * acc[mi][ni][e] + (nr + f.nC[wn * 64 + ni * 8 + 2 * t + e]);
                const double kv = (mode == 2) ? d2 : sig2 * exp_neg(d2);
                const long long r = (long long)tr * TILE + acc_row(mi);
                const long long c = (long long)tc * TILE + acc_col(ni, e);
                if (r < nrows && c < ncols) ob[r * ld + c] = (mode == 0 && r == c) ? kv + dadd : kv;
                else if (r < lim_r && c < lim_c) ob[r * ld + c] = (r == c) ? 1.0 : 0.0;
            }
    }
}

// launches the tile builder that fits: the whole feature tiles in shared memory, or slices of them for wide inputs
static int launch_se_tile(dim3 grid, cudaStream_t st, const double* aR, const double* nrmR, int rpad, const double* aC,
                          const double* nrmC, int cpad, const double* hp, int hp_stride, const double* diag_add, int nrows, int ncols,
                          int dpad, double* out, long long ld, long long out_bstride, int mode, int pad_identity, int lower_only) {
    const int smem = feat_smem_bytes(dpad);
    if (smem <= 227 * 1024) {
        CUDA_TRY(opt_in_smem(se_tile_kernel, smem));
        se_tile_kernel<<<grid, NTHREADS, smem, st>>>(aR, nrmR, rpad, aC, nrmC, cpad, hp, hp_stride, diag_add, nrows, ncols, dpad, out, ld,
                                                     out_bstride, mode, pad_identity, lower_only);
    } else {
        const int smem_w = feat_smem_bytes(SE_SLICE);
        CUDA_TRY(opt_in_smem(se_tile_wide_kernel, smem_w));
        se_tile_wide_kernel<<<grid, NTHREADS, smem_w, st>>>(aR, nrmR, rpad, aC, nrmC, cpad, hp, hp_stride, diag_add, nrows, ncols, dpad, out,
                                                            ld, out_bstride, mode, pad_identity, lower_only);
    }
    CUDA_TRY(cudaGetLastError());
    return GPR_OK;
}

extern "C" int gpr_se_build(const double* a, const double* nrm, const double* hp, int hp_stride, const double* diag_add,
                            int B, int n, int npad, int dpad, double* K, int ldk, int pad_identity, int lower_only,
                            void* stream) {
    ARG_CHECK(a && nrm && hp && K);
    ARG_CHECK(B > 0 && n > 0 && npad % TILE == 0 && npad >= n && dpad % 4 == 0 && ldk >= n);
    ARG_CHECK(!pad_identity || ldk <= npad);
    const int nt = (pad_identity ? (ldk + TILE - 1) / TILE : (n + TILE - 1) / TILE);
    return launch_se_tile(dim3(nt, nt, B), (cudaStream_t)stream, a, nrm, npad, a, nrm, npad, hp, hp_stride, diag_add, n, n, dpad, K, ldk,
                          (long long)ldk * ldk, 0, pad_identity, lower_only);
}

extern "C" int gpr_se_cross(const double* as, const double* nrms, int mpad, const double* a, const double* nrm, int npad,
                            const double* hp, int hp_stride, int B, int m, int n, int dpad, double* Ks, void* stream) {
    ARG_CHECK(as && nrms && a && nrm && hp && Ks);
    ARG_CHECK(B > 0 && m > 0 && n > 0 && mpad % TILE == 0 && npad % TILE == 0 && mpad >= m && npad >= n && dpad % 4 == 0);
    return launch_se_tile(dim3((n + TILE - 1) / TILE, (m + TILE - 1) / TILE, B), (cudaStream_t)stream, as, nrms, mpad, a, nrm, npad, hp,
                          hp_stride, nullptr, m, n, dpad, Ks, n, (long long)m * n, 1, 0, 0);
}

extern "C" int gpr_se_cross_ld(const double* as, const double* nrms, int mpad, const double* a, const double* nrm, int npad,
                               const double* hp, int hp_stride, int B, int m, int n, int dpad, double* Ks, long long ld, long long bstride,
                               void* stream) {
    ARG_CHECK(as && nrms && a && nrm && hp && Ks);
    ARG_CHECK(B > 0 && m > 0 && n > 0 && mpad % TILE == 0 && npad % TILE == 0 && mpad >= m && npad >= n && dpad % 4 == 0);
    ARG_CHECK(ld >= n && bstride >= (long long)m * ld);
    return launch_se_tile(dim3((n + TILE - 1) / TILE, (m + TILE - 1) / TILE, B), (cudaStream_t)stream, as, nrms, mpad, a, nrm, npad, hp,
                          hp_stride, nullptr, m, n, dpad, Ks, ld, bstride, 1, 0, 0);
}

// ---------------------------------------------------------------------------------------------
// K* straight into the digit planes of the emulated-FP64 product (ozaki_gemm.cu): the tile of gpr_se_cross is never written in
// FP64 and read back by gpr_ozaki_split — every entry is bounded by `bound` (sig^2 for an SE kernel matrix), so the row scale is
// known before the tile exists and the balanced radix-256 digits (the arithmetic of oz::split_kernel, bit for bit) come out of the
// epilogue and leave as whole 32-byte sectors.  Columns n .. npad of the planes are written as zeros (padded k-range).
//   planes [slices][B * mpad][npad] int8, scale [B * mpad]
// ---------------------------------------------------------------------------------------------
constexpr int SEP_LD = 48;   // row stride (bytes) of the per-warp transposition buffer: 32 B of digits + pad (rows 12 banks apart)
__global__ void __launch_bounds__(NTHREADS) se_cross_planes_kernel(const double* __restrict__ aR, const double* __restrict__ nrmR, int mpad,
                                                                   const double* __restrict__ aC, const double* __restrict__ nrmC, int npad,
                                                                   const double* __restrict__ hp, int hp_stride, int n, int dpad,
                                                                   int slices, double mult, double row_scale,
                                                                   signed char* __restrict__ planes, double* __restrict__ scale) {
    extern __shared__ double smem[];
    const int tc = blockIdx.x, tr = blockIdx.y, b = blockIdx.z;
    FeatSmem f = carve_feat(smem, dpad);
    load_feat_tile(f.sR, f.sld, aR + ((long long)b * mpad + tr * TILE) * dpad, dpad);
    load_feat_tile(f.sC, f.sld, aC + ((long long)b * npad + tc * TILE) * dpad, dpad);
    cp_async_commit();
    if (threadIdx.x < TILE) {
        f.nR[threadIdx.x] = nrmR[(long long)b * mpad + tr * TILE + threadIdx.x];
        f.nC[threadIdx.x] = nrmC[(long long)b * npad + tc * TILE + threadIdx.x];
        if (tc == 0) scale[(long long)b * mpad + tr * TILE + threadIdx.x] = row_scale;
    }
    cp_async_wait<0>();
    __syncthreads();
    const double sig = hp[(long long)b * hp_stride];
    const long long plane_stride = (long long)gridDim.z * mpad * npad;
    signed char* out = planes + ((long long)b * mpad + tr * TILE) * npad + tc * TILE;
    // Balanced radix-256 digits without the digit-by-digit carry loop: with J = I + sum_t 128 * 256^t the bytes of J are d_t + 128,
    // so the digits are the bytes of J ^ 0x80..80 (the representation with d_t in [-128, 127] is unique: the same bytes as
    // oz::split_kernel's loop).  Byte k (least significant first) belongs to plane slices-1-k.
    unsigned long long bias = 0ull;
    for (int t = 0; t < slices; ++t) bias |= 0x80ull << (8 * t);
    // A warp's fragment covers 32 rows x 64 columns, 8 columns (one ni) at a time.  The digit words of four ni (32 columns = one
    // 32-byte sector per row and plane) are collected in registers, then plane by plane transposed through a per-warp buffer
    // [32 rows][48 B] and written as whole sectors (16-bit stores straight from the fragments ran the L2 at four partial-sector
    // writes per sector: 2.1 ms per 8 experts instead of 0.75 + 1.13 ms for the FP64 tile + gpr_ozaki_split).
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t4 = lane & 3;
    signed char* wbuf = reinterpret_cast<signed char*>(smem) + ((feat_smem_bytes(dpad) + 15) & ~15) + warp * (32 * SEP_LD);
    unsigned long long Wb[4][4][2];
    gen_se_tile<false>(f.sR, f.sC, f.sld, f.nR, f.nC, dpad, sig * sig, [&](int mi, int ni, int e, double kv) {
        const int c = acc_col(ni, e);
        const long long I = (tc * TILE + c < n) ? __double2ll_rn(kv * mult) : 0LL;
        Wb[ni & 3][mi][e] = ((unsigned long long)I + bias) ^ bias;
        if ((ni & 3) == 3 && mi == 3 && e == 1) {
            const int row0 = (warp & 3) * 32, col0 = (warp >> 2) * 64 + (ni >> 2) * 32;
#pragma unroll
            for (int k = 0; k < 8; ++k) {                   // byte k of every collected word -> plane slices-1-k
                if (k < slices) {
#pragma unroll
                    for (int q = 0; q < 4; ++q)
#pragma unroll
                        for (int m2 = 0; m2 < 4; ++m2) {
                            const unsigned long long w0 = Wb[q][m2][0], w1 = Wb[q][m2][1];
                            const unsigned pair = k < 4 ? __byte_perm((unsigned)w0, (unsigned)w1, 0x0040 + 0x11 * k)
                                                        : __byte_perm((unsigned)(w0 >> 32), (unsigned)(w1 >> 32), 0x0040 + 0x11 * (k - 4));
                            *reinterpret_cast<unsigned short*>(wbuf + (m2 * 8 + g) * SEP_LD + q * 8 + 2 * t4) = (unsigned short)pair;
                        }
                    __syncwarp();
                    signed char* dst = out + (long long)(slices - 1 - k) * plane_stride + col0;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int idx = lane + 32 * h, rr = idx >> 1, half = idx & 1;
                        *reinterpret_cast<int4*>(dst + (long long)(row0 + rr) * npad + 16 * half) =
                            *reinterpret_cast<const int4*>(wbuf + rr * SEP_LD + 16 * half);
                    }
                    __syncwarp();
                }
            }
        }
    });
}

static inline int se_cross_planes_smem(int dpad, int slices) { (void)slices; return ((feat_smem_bytes(dpad) + 15) & ~15) + (NTHREADS / 32) * 32 * SEP_LD; }

// 1 when the fused builder can run for this input width / plane count (shared memory), else the caller takes gpr_se_cross_ld + gpr_ozaki_split
extern "C" int gpr_se_cross_planes_ok(int dpad, int slices) {
    return dpad % 4 == 0 && dpad <= 108 && slices >= 1 && slices <= 8 && se_cross_planes_smem(dpad, slices) <= 226 * 1024;
}

extern "C" int gpr_se_cross_planes(const double* as, const double* nrms, int mpad, const double* a, const double* nrm, int npad,
                                   const double* hp, int hp_stride, int B, int n, int dpad, int slices, double bound, void* planes,
                                   double* scale, void* stream) {
    ARG_CHECK(as && nrms && a && nrm && hp && planes && scale);
    ARG_CHECK(B > 0 && n > 0 && mpad % TILE == 0 && npad % TILE == 0 && npad >= n && bound > 0.0 && isfinite(bound));
    if (!gpr_se_cross_planes_ok(dpad, slices)) return fail(GPR_ERR_UNSUPPORTED, "fused K* digit planes: input width / plane count too large%s (dpad=%lld)", "", dpad);
    // the row scale of oz::split_kernel for a known bound: |x| 2^-e < 1/4, scale = 2^(e-8), digits of rint(x 2^(8 slices - e))
    const int e = ilogb(bound) + 3;
    const int smem_bytes = se_cross_planes_smem(dpad, slices);
    CUDA_TRY(opt_in_smem(se_cross_planes_kernel, smem_bytes));
    se_cross_planes_kernel<<<dim3(npad / TILE, mpad / TILE, B), NTHREADS, smem_bytes, (cudaStream_t)stream>>>(
        as, nrms, mpad, a, nrm, npad, hp, hp_stride, n, dpad, slices, ldexp(1.0, 8 * slices - e), ldexp(1.0, e - 8),
        static_cast<signed char*>(planes), scale);
    CUDA_TRY(cudaGetLastError());
    return GPR_OK;
}

// Squared Euclidean distances in GEMM form, D[i,j] = |x_i|^2 + |y_j|^2 - 2 x_i.y_j   (sampler.py:94-100, N3)
extern "C" int gpr_sq_dist(const double* ax, const double* nx, int npadx, const double* ay, const double* ny, int npady, int nx_rows,
                           int ny_rows, int dpad, double* D, void* stream) {
    ARG_CHECK(ax && nx && ay && ny && D && nx_rows > 0 && ny_rows > 0 && npadx % TILE == 0 && npady % TILE == 0 && dpad % 4 == 0);
    // hp is only read for sig (unused in raw mode); pass nx as a valid device pointer
    return launch_se_tile(dim3((ny_rows + TILE - 1) / TILE, (nx_rows + TILE - 1) / TILE, 1), (cudaStream_t)stream, ax, nx, npadx, ay, ny,
                          npady, nx, 0, nullptr, nx_rows, ny_rows, dpad, D, ny_rows, 0, 2, 0, 0);
}

// dK materialisation for kernel_and_grad (API parity; small n only)                 covar.py:184-199
__global__ void se_grad_kernel(const double* __restrict__ Kse, const double* __restrict__ x, const double* __restrict__ hp,
                               int hp_stride, int n, int d, double* __restrict__ dK) {
    const int b = blockIdx.z, p = blockIdx.y;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)n * n) return;
    const int i = idx / n, j = idx % n;
    const double k = Kse[(long long)b * n * n + idx];
    const double* h = hp + (long long)b * hp_stride;
    double v;
    if (p == 0) {
        v = k * (2.0 / h[0]);  // covar.py:189
    } else {
        const double* xb = x + (long long)b * n * d;
        const double df = xb[(long long)i * d + (p - 1)] - xb[(long long)j * d + (p - 1)];
        v = df * df * h[p] * k * (-2.0);  // covar.py:194-197
    }
    dK[(((long long)b * (d + 1) + p) * n) * n + idx] = v;
}

extern "C" int gpr_se_kernel_grad(const double* Kse, const double* x, const double* hp, int hp_stride, int B, int n, int d,
                                  double* dK, void* stream) {
    ARG_CHECK(Kse && x && hp && dK && B > 0 && n > 0 && d > 0);
    dim3 grid((unsigned)(((long long)n * n + 255) / 256), d + 1, B);
    se_grad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(Kse, x, hp, hp_stride, n, d, dK);
    CUDA_TRY(cudaGetLastError());
    return GPR_OK;
}

// ---------------------------------------------------------------------------------------------
// K3: batched blocked left-looking Cholesky.  Two launches per block column j:
//   potrf_diag_kernel : T = K_jj - L[j,0:j] L[j,0:j]^T -> potf2 + inverse in registers -> L_jj, inv(L_jj);
//                       forward solve z_j = inv(L_jj)(y_j - L[j,0:j] z[0:j]) fused (GEMV on the chunks
//                       already in smem), plus sum log L_ii and z.z  -> NLML pieces for free.
//   potrf_panel_kernel: L_ij = (K_ij - L[i,0:j] L[j,0:j]^T) inv(L_jj)^T for every row tile i>j (DMMA twice).
// Look-ahead: the diagonal kernel sits on the critical path with only B CTAs, so all of its product except the last
// 128-deep block is computed one launch earlier by an extra CTA per expert inside the (wide) panel launch of column
// j-1 and parked in the unused upper block (j-1, j) of L; the diagonal kernel then costs O(1) tiles, not O(j).
// ---------------------------------------------------------------------------------------------
// ---- inter-CTA progress flags (release / acquire at GPU scope; bounded polling as a hang guard) ----------------------------
constexpr unsigned long long DF_WAIT_LIMIT_NS = 30ull * 1000ull * 1000ull * 1000ull;   // 30 s of ELAPSED time (globaltimer): far beyond
                                                                                         // any legitimate wait, short of a wedged box
#define GPR_INFO_STALLED (-2147483647 - 1)   // info[] sentinel: the launch was aborted by the hang guard (not a numerical verdict)

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// All threads call.  Thread 0 polls, the CTA barrier hands its acquire to everyone.  false = aborted: the abort flag is raised
// and every CTA drains; report_stall_kernel, enqueued behind the launch, then marks EVERY info entry GPR_INFO_STALLED so the host
// can tell a scheduling stall from a non-positive-definite matrix (no tile task can overwrite it any more at that point).
__device__ __forceinline__ bool df_wait(const int* flag, int need, int* abort_flag, int* info, int* s_ok) {
    (void)info;
    if (threadIdx.x == 0) {
        int ok = 1;
        if (ld_acquire_gpu(flag) < need) {
            const unsigned long long t0 = globaltimer_ns();
            for (;;) {
                __nanosleep(100);
                if (ld_acquire_gpu(flag) >= need) break;
                if (*reinterpret_cast<volatile int*>(abort_flag) || globaltimer_ns() - t0 > DF_WAIT_LIMIT_NS) {
                    atomicExch(abort_flag, 1);
                    ok = 0;
                    break;
                }
            }
        }
        *s_ok = ok;
    }
    __syncthreads();
    const bool ok = *s_ok != 0;
    __syncthreads();  // *s_ok is rewritten by the next wait
    return ok;
}

__global__ void report_stall_kernel(const int* __restrict__ abort_flag, int* __restrict__ info, int B, int* __restrict__ info2, int B2) {
    if (*abort_flag == 0) return;
    for (int i = threadIdx.x; i < B; i += blockDim.x) info[i] = GPR_INFO_STALLED;
    if (info2)
        for (int i = threadIdx.x; i < B2; i += blockDim.x) info2[i] = GPR_INFO_STALLED;
}

struct GemvHook {  // s[r] += sum_k A[r][k] z[k] on every landed chunk; half-warp per row, 8 rows per half-warp
    const double* z;
    double zk;
    double part[8];
    __device__ __forceinline__ void init(const double* zp) {
        z = zp;
#pragma unroll
        for (int r = 0; r < 8; ++r) part[r] = 0.0;
    }
    __device__ __forceinline__ void pre(int c) { zk = z ? __ldcg(z + c * KC + (threadIdx.x & 15)) : 0.0; }  // L2: other CTAs write z
    __device__ __forceinline__ void post(const double* sA, int) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const double* p = sA + (warp * 16 + (lane >> 4) * 8) * SLD + (lane & 15);
#pragma unroll
        for (int r = 0; r < 8; ++r) part[r] += p[r * SLD] * zk;
    }
    // out[row] = (add ? add[row] : 0) + reduced partial; rows owned by lanes 0 and 16 of each warp
    __device__ __forceinline__ void reduce_to(double* out, const double* add) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            double v = part[r];
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            const int row = warp * 16 + (lane >> 4) * 8 + r;
            if ((lane & 15) == 0) out[row] = v + (add ? add[row] : 0.0);
        }
    }
};

// Tail of a diagonal-tile task, shared by the per-column kernel and the dataflow kernel: T = K_jj - acc (- dpart) ->
// potf2_tile + trtri_tile (gpr_diag_tile.cuh) -> L_jj, inv(L_jj), log-determinant / info, forward-solve block z_j.
// Scalars that several CTAs update in turn (logdet_half, zz, info) and z go through L2 (__ldcg/__stcg): in the dataflow
// kernel consecutive block columns of one expert run on different SMs and L1 is not coherent.
// developer build (-DGPR_DIAG_STAMPS, tools/diag_phases.py): clock stamps of the phases of a diagonal-tile task, expert 0
#ifdef GPR_DIAG_STAMPS
__device__ long long g_diag_stamps[16];
#define DIAG_STAMP(i) do { if (threadIdx.x == 0 && b == 0) g_diag_stamps[i] = clock64(); } while (0)
extern "C" int gpr_debug_diag_stamps(long long* out) { return cudaMemcpyFromSymbol(out, g_diag_stamps, sizeof(g_diag_stamps)) == cudaSuccess ? 0 : 1; }
#else
#define DIAG_STAMP(i) do { } while (0)
#endif

template <bool GEN>
__device__ __forceinline__ void diag_finish(double (&acc)[4][8][2], double* smem, double* pub, double* svec, int* s_fail, double* Lb,
                                            int b, int j, int n, int npad, const double* __restrict__ a,
                                            const double* __restrict__ nrm, const double* __restrict__ hp, int hp_stride,
                                            const double* __restrict__ diag_add, int dpad, double* invdiag, const double* y,
                                            double* zb, double* logdet_half, double* zz, int* info, const double* dpart,
                                            bool first = false, bool use_first = false, const double* base_ld = nullptr,
                                            const double* base_zz = nullptr, const int* base_info = nullptr, int base_n = 0,
                                            double* ld_blk = nullptr, double* zz_blk = nullptr) {
    // base_ld / base_zz: per-block-column partial sums of ANOTHER factorisation (the communication expert's, same launch);
    // at `start` the first base_n of them are the value the sums continue from.  ld_blk / zz_blk: where THIS expert records
    // its own per-block-column partials (row b*nt + j), for exactly that use.
    // first: this is the expert's first diagonal task (logdet / z.z / info start here).  Default (use_first == false): j == 0.
    // base_*: values the sums start from (the communication expert's, when it is factorised in the same launch).
    const int tid = threadIdx.x;
    const int nt = npad / TILE;
    const bool start = use_first ? first : (j == 0);
    double* Ljj = Lb + ((long long)j * TILE) * npad + j * TILE;
    if (tid == 0) *s_fail = 1 << 30;
    double* T = smem;  // full tile overlays the (drained) pipeline stages
    // the look-ahead partial product parked by the previous column joins the accumulators first: eight independent 16-byte loads
    // per batch (inside the element-wise epilogues below every load sat behind the previous store: ~400 clocks each, serial)
    if (dpart != nullptr) {
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) {
            double2 p2[8];
#pragma unroll
            for (int ni = 0; ni < 8; ++ni) p2[ni] = *reinterpret_cast<const double2*>(dpart + (long long)acc_row(mi) * npad + acc_col(ni, 0));
#pragma unroll
            for (int ni = 0; ni < 8; ++ni) {
                acc[mi][ni][0] += p2[ni].x;
                acc[mi][ni][1] += p2[ni].y;
            }
        }
    }
    if (GEN) {
        double* fbase = smem + TILE * XLD;
        FeatSmem f = carve_feat(fbase, dpad);
        load_feat_tile(f.sR, f.sld, a + ((long long)b * npad + j * TILE) * dpad, dpad);
        cp_async_commit();
        if (tid < TILE) f.nR[tid] = nrm[(long long)b * npad + j * TILE + tid];
        cp_async_wait<0>();
        __syncthreads();
        const double sig = hp[(long long)b * hp_stride];
        const double dadd = diag_add ? diag_add[b] : 0.0;
        gen_se_tile(f.sR, f.sR, f.sld, f.nR, f.nR, dpad, sig * sig, [&](int mi, int ni, int e, double kv) {
            const int r = acc_row(mi), c = acc_col(ni, e);
            const int gr = j * TILE + r, gc = j * TILE + c;
            double v;
            if (gr < n && gc < n) v = (r == c) ? kv + dadd : kv;
            else v = (r == c) ? 1.0 : 0.0;
            T[r * XLD + c] = v - acc[mi][ni][e];
        });
    } else {
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) {
            const int r = acc_row(mi);
            double2 k2[8];
#pragma unroll
            for (int ni = 0; ni < 8; ++ni) k2[ni] = *reinterpret_cast<const double2*>(Ljj + (long long)r * npad + acc_col(ni, 0));
#pragma unroll
            for (int ni = 0; ni < 8; ++ni)
                *reinterpret_cast<double2*>(T + r * XLD + acc_col(ni, 0)) = make_double2(k2[ni].x - acc[mi][ni][0], k2[ni].y - acc[mi][ni][1]);
        }
    }
    __syncthreads();
    DIAG_STAMP(3);
    double* Bm = smem + TILE * XLD;  // scratch of the tile routines, behind the tile (the GEN feature tiles there are spent)
    potf2_tile(T, Bm, s_fail);
    DIAG_STAMP(4);

    // write L_jj (strict upper zeroed) and accumulate sum log L_ii
    for (int id = tid; id < TILE * TILE; id += NTHREADS) {
        const int r = id >> 7, c = id & 127;
        Ljj[(long long)r * npad + c] = (c <= r) ? T[r * XLD + c] : 0.0;
    }
    if (tid < 32) {
        double s = 0.0;
        for (int k = tid; k < TILE; k += 32) s += log(T[k * XLD + k]);
        s = warp_sum(s);
        if (tid == 0) {
            double prev_ld = start ? 0.0 : __ldcg(logdet_half + b);
            if (start && base_ld)
                for (int q = 0; q < base_n; ++q) prev_ld += __ldcg(base_ld + q);
            if (logdet_half) __stcg(logdet_half + b, prev_ld + s);
            if (ld_blk) __stcg(ld_blk + (long long)b * nt + j, s);
            if (info) {
                int prev = start ? (base_info ? __ldcg(base_info) : 0) : __ldcg(info + b);
                if (start && base_info && prev > base_n * TILE) prev = 0;  // the other factorisation failed outside the shared block
                if (start || (prev == 0 && *s_fail <= TILE)) __stcg(info + b, prev != 0 ? prev : (*s_fail <= TILE ? j * TILE + *s_fail : 0));
            }
        }
    }
    DIAG_STAMP(5);
    trtri_tile(T, Bm);
    DIAG_STAMP(6);
    double* inv = invdiag + ((long long)b * nt + j) * TILE * TILE;
    for (int id = tid; id < TILE * TILE; id += NTHREADS) inv[id] = T[(id >> 7) * XLD + (id & 127)];
    DIAG_STAMP(7);
    if (y) {
        // z_j = inv(L_jj) (y_j - L[j,0:j] z[0:j]); warp per row, lanes along k, fixed-order reduction
        if (tid < TILE) pub[tid] = y[(long long)b * npad + j * TILE + tid] - svec[tid];
        __syncthreads();
        const int lane = tid & 31, warp = tid >> 5;
        // four rows of a warp in flight: their shuffle reductions overlap (one at a time the 16 rows of a warp cost 16 x 10 dependent
        // shuffles, ~12 k clocks of the chain)
        const double pk0 = pub[lane], pk1 = pub[lane + 32], pk2 = pub[lane + 64], pk3 = pub[lane + 96];
#pragma unroll
        for (int rr = 0; rr < TILE / (NTHREADS / 32); rr += 4) {
            double s[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int r = warp + (rr + q) * (NTHREADS / 32);
                const double* Tr = T + r * XLD + lane;     // X is lower triangular with zeros above: whole rows can be summed
                s[q] = Tr[0] * pk0;
                if (r >= 32) s[q] += Tr[32] * pk1;
                if (r >= 64) s[q] += Tr[64] * pk2;
                if (r >= 96) s[q] += Tr[96] * pk3;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                for (int q = 0; q < 4; ++q) s[q] += __shfl_xor_sync(0xffffffffu, s[q], o);
            if (lane == 0) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int r = warp + (rr + q) * (NTHREADS / 32);
                    zb[j * TILE + r] = s[q];
                    svec[r] = s[q] * s[q];
                }
            }
        }
        __syncthreads();
        if (tid < 32 && zz) {
            double s = svec[tid] + svec[tid + 32] + svec[tid + 64] + svec[tid + 96];
            s = warp_sum(s);
            if (tid == 0) {
                double prev_zz = start ? 0.0 : __ldcg(zz + b);
                if (start && base_zz)
                    for (int q = 0; q < base_n; ++q) prev_zz += __ldcg(base_zz + q);
                __stcg(zz + b, prev_zz + s);
                if (zz_blk) __stcg(zz_blk + (long long)b * nt + j, s);
            }
        }
    }
    DIAG_STAMP(8);
}

template <bool GEN>
__global__ void __launch_bounds__(NTHREADS, 1)
potrf_diag_kernel(double* __restrict__ L, int npad, int n, int j, const double* __restrict__ a, const double* __restrict__ nrm,
                  const double* __restrict__ hp, int hp_stride, const double* __restrict__ diag_add, int dpad,
                  double* __restrict__ invdiag, const double* __restrict__ y, double* __restrict__ z,
                  double* __restrict__ logdet_half, double* __restrict__ zz, int* __restrict__ info, int* __restrict__ done_flag) {
    extern __shared__ double smem[];
    __shared__ unsigned long long pipe_bars[2 * STAGES];
    Pipe pipe;
    pipe_init(pipe, pipe_bars);
    __shared__ double pub[3 * TILE];
    __shared__ double svec[TILE];
    __shared__ int s_fail;
    const int b = blockIdx.x;
    DIAG_STAMP(0);
    double* Lb = L + (long long)b * npad * npad;
    double* zb = y ? z + (long long)b * npad : nullptr;

    // product over the LAST block column only (k-block j-1); blocks 0..j-2 were summed by the look-ahead CTA
    double acc[4][8][2];
    zero_acc(acc);
    GemvHook hook;
    hook.init(zb && j > 0 ? zb + (long long)(j - 1) * TILE : nullptr);
    const double* Arow = Lb + ((long long)j * TILE) * npad + (long long)(j > 0 ? j - 1 : 0) * TILE;
    Seg s0{Arow, npad, Arow, npad, j > 0 ? TILE / KC : 0};
    Seg s1{nullptr, 0, nullptr, 0, 0};
    mainloop<true, XPF_POTRF>(acc, s0, s1, smem, hook, pipe);
    DIAG_STAMP(1);
    const bool have_part = j >= 2;
    const double* dpart = Lb + ((long long)(j - 1) * TILE) * npad + (long long)j * TILE;  // upper block (j-1, j)
    hook.reduce_to(svec, (have_part && zb) ? zb + (long long)j * TILE : nullptr);

    diag_finish<GEN>(acc, smem, pub, svec, &s_fail, Lb, b, j, n, npad, a, nrm, hp, hp_stride, diag_add, dpad, invdiag, y, zb,
                     logdet_half, zz, info, have_part ? dpart : nullptr);
    if (done_flag) {  // the first panel tile of this column runs concurrently and waits for inv(L_jj) only before its solve
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) st_release_gpu(done_flag + b, j + 1);
    }
    DIAG_STAMP(9);
}

// grid.x = (nt-1-j) row tiles, plus one look-ahead CTA per expert when lookahead != 0 (j >= 1, j+1 < nt):
// that CTA forms  P = L[j+1,0:j] L[j+1,0:j]^T  (k-blocks 0..j-1, complete since the previous launches) for the NEXT
// diagonal tile into the upper block (j, j+1) and the matching forward-solve partial into z[(j+1)*128 ...].
template <bool GEN>
__global__ void __launch_bounds__(NTHREADS, 1)
potrf_panel_kernel(double* __restrict__ L, int npad, int n, int j, const double* __restrict__ a, const double* __restrict__ nrm,
                   const double* __restrict__ hp, int hp_stride, int dpad, const double* __restrict__ invdiag, int lookahead,
                   double* __restrict__ z, int S, double* __restrict__ sk_slots, double* __restrict__ sk_gemv,
                   int* __restrict__ sk_counters, int row0, int* __restrict__ diag_flag) {
    extern __shared__ double smem[];
    __shared__ unsigned long long pipe_bars[2 * STAGES];
    __shared__ int s_flag;
    Pipe pipe;
    pipe_init(pipe, pipe_bars);
    const int b = blockIdx.y;
    const int nt = npad / TILE;
    const int ntile_x = gridDim.x / S;                 // row tiles (+ the look-ahead tile) in this launch; first row tile j+1+row0
    const int tx = blockIdx.x / S, part = blockIdx.x % S;
    // this CTA's slice of the j k-blocks
    const int kb0 = (int)((long long)part * j / S), kb1 = (int)((long long)(part + 1) * j / S);
    double* slots = sk_slots + ((long long)b * ntile_x + tx) * S * TILE * TILE;
    int* counter = sk_counters + (long long)b * nt + tx;
    double* Lb = L + (long long)b * npad * npad;
    double acc[4][8][2];
    zero_acc(acc);
    Seg none{nullptr, 0, nullptr, 0, 0};

    if (lookahead && tx == ntile_x - 1) {
        const double* Arow = Lb + ((long long)(j + 1) * TILE) * npad + (long long)kb0 * TILE;
        GemvHook hook;
        hook.init(z ? z + (long long)b * npad + (long long)kb0 * TILE : nullptr);
        Seg s0{Arow, npad, Arow, npad, (kb1 - kb0) * (TILE / KC)};
        mainloop<true, XPF_POTRF>(acc, s0, none, smem, hook, pipe);
        double* zdst = z ? z + (long long)b * npad + (long long)(j + 1) * TILE : nullptr;
        if (S > 1) {
            double* gv = sk_gemv + (long long)b * S * TILE;
            if (z) hook.reduce_to(gv + part * TILE, nullptr);
            if (!splitk_combine(acc, slots, part, S, counter, &s_flag)) return;
            if (z && threadIdx.x < TILE) {
                double v = 0.0;
                for (int p = 0; p < S; ++p) v += __ldcg(gv + p * TILE + threadIdx.x);
                zdst[threadIdx.x] = v;
            }
        } else if (z) {
            hook.reduce_to(zdst, nullptr);
        }
        double* dpart = Lb + ((long long)j * TILE) * npad + (long long)(j + 1) * TILE;
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < 8; ++ni)
                *reinterpret_cast<double2*>(dpart + (long long)acc_row(mi) * npad + acc_col(ni, 0)) =
                    make_double2(acc[mi][ni][0], acc[mi][ni][1]);
        return;
    }

    const int i = j + 1 + row0 + tx;
    double* Lij = Lb + ((long long)i * TILE) * npad + j * TILE;
    NoHook nh;
    Seg s0{Lb + ((long long)i * TILE) * npad + (long long)kb0 * TILE, npad, Lb + ((long long)j * TILE) * npad + (long long)kb0 * TILE, npad,
           (kb1 - kb0) * (TILE / KC)};
    mainloop<false, XPF_POTRF>(acc, s0, none, smem, nh, pipe);
    if (S > 1 && !splitk_combine(acc, slots, part, S, counter, &s_flag)) return;

    if (GEN) {
        FeatSmem f = carve_feat(smem, dpad);
        load_feat_tile(f.sR, f.sld, a + ((long long)b * npad + i * TILE) * dpad, dpad);
        load_feat_tile(f.sC, f.sld, a + ((long long)b * npad + j * TILE) * dpad, dpad);
        cp_async_commit();
        if (threadIdx.x < TILE) {
            f.nR[threadIdx.x] = nrm[(long long)b * npad + i * TILE + threadIdx.x];
            f.nC[threadIdx.x] = nrm[(long long)b * npad + j * TILE + threadIdx.x];
        }
        cp_async_wait<0>();
        __syncthreads();
        const double sig = hp[(long long)b * hp_stride];
        gen_se_tile(f.sR, f.sC, f.sld, f.nR, f.nC, dpad, sig * sig, [&](int mi, int ni, int e, double kv) {
            const int gr = i * TILE + acc_row(mi), gc = j * TILE + acc_col(ni, e);
            acc[mi][ni][e] = ((gr < n && gc < n) ? kv : 0.0) - acc[mi][ni][e];
        });
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < 8; ++ni)
                *reinterpret_cast<double2*>(Lij + (long long)acc_row(mi) * npad + acc_col(ni, 0)) =
                    make_double2(acc[mi][ni][0], acc[mi][ni][1]);
    } else {
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < 8; ++ni) {
                double2* p = reinterpret_cast<double2*>(Lij + (long long)acc_row(mi) * npad + acc_col(ni, 0));
                double2 k2 = *p;
                k2.x -= acc[mi][ni][0];
                k2.y -= acc[mi][ni][1];
                *p = k2;
            }
    }
    __threadfence_block();
    __syncthreads();
    if (diag_flag) {  // launched alongside the diagonal-tile kernel of this column: inv(L_jj) is the only thing it owes us
        if (!df_wait(diag_flag + b, j + 1, diag_flag - 1, nullptr, &s_flag)) return;
    }

    // triangular solve as a product with the inverted diagonal block:  L_ij = Xhat * inv(L_jj)^T
    zero_acc(acc);
    Seg s1{Lij, npad, invdiag + ((long long)b * nt + j) * TILE * TILE, TILE, TILE / KC};
    mainloop<false, XPF_POTRF>(acc, s1, none, smem, nh, pipe);
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 8; ++ni)
            *reinterpret_cast<double2*>(Lij + (long long)acc_row(mi) * npad + acc_col(ni, 0)) =
                make_double2(acc[mi][ni][0], acc[mi][ni][1]);
}

// ---------------------------------------------------------------------------------------------
// K3 (dataflow form): the whole batched factorisation in ONE persistent launch.  Tile tasks are handed out by an atomic
// counter in an order in which every dependency has a smaller index (block column by block column; inside a column
// the B diagonal tasks first, then the panel tiles expert by expert), and per-(expert, row tile) progress counters
// replace the launch boundaries:   progress[b][i] = number of finished tiles in block row i  (j+1 after tile (i, j),
// i+1 after the diagonal task of row i).  A task only ever waits for tasks that a RUNNING CTA already holds, so the
// scheme cannot deadlock whatever the number of resident CTAs.  What it buys over one diag + one panel launch per
// column: no ragged last wave per column, the potf2/trti2 of the diagonal tile (tensor-idle, ~6 tile products long)
// overlaps other experts' products instead of serialising the column, and no look-ahead round trip through HBM.
//   D(b, j):  T = K_jj - L[j,0:j] L[j,0:j]^T  -> potf2/trti2 -> L_jj, inv(L_jj), z_j      waits (j, j-1) for the last block
//   P(b,i,j): X = K_ij - L[i,0:j] L[j,0:j]^T ; waits for D(b, j) only before  L_ij = X inv(L_jj)^T
// Waits are bounded (DF_WAIT_LIMIT_NS of elapsed time): on a time-out the kernel raises the abort flag, every CTA drains, and
// info[*] = GPR_INFO_STALLED tells the host (report_stall_kernel); the bound is a guard against a wedged GPU, not part of the protocol.
// ---------------------------------------------------------------------------------------------
struct PotrfArgs {
    double* L;
    int B, npad, n;
    const double *a, *nrm, *hp;
    int hp_stride;
    const double* diag_add;
    int dpad;
    double* invdiag;
    const double* y;
    double* z;
    double* logdet_half;
    double* zz;
    int* info;
    int* ctl;       // [0] next task, [1] abort flag
    int* progress;  // [B, nt]
    // Shared leading block (GRBCM with one hyper-parameter vector, gr_bcm.py:19-29): the first gt block rows/columns of every
    // expert are the communication set, so L[0:gt, 0:gt] and inv(L_jj), j < gt, equal the communication expert's factor.
    // They are read from (Lg, invdiag_g) and never computed; tasks exist only for block rows i >= gt.  gt = 0: off.
    int gt;
    const double* Lg;
    int npad_g;
    const double* invdiag_g;
    // Communication expert factorised IN THE SAME LAUNCH (gpr_potrf_grbcm): its tasks come first in every shared block column,
    // local tiles of those columns wait on its progress counters (gprog), its z blocks are broadcast into every expert's z
    // (zbcast: [zbcast_B, zbcast_stride]) and the experts' logdet / z.z / info start from its totals (base_*).
    const int* gprog;
    double* zbcast;
    int zbcast_B, zbcast_stride, zbcast_gt;   // only the first zbcast_gt blocks of z are shared
    const double *base_ld, *base_zz;          // the communication expert's per-block-column partials (first gt are the base)
    const int* base_info;
    double *ld_blk, *zz_blk;                  // where this factorisation records its own per-block-column partials
};

// The two task bodies are separate NON-inlined functions so that each gets the register allocation of a stand-alone
// kernel (inlined into the persistent loop the four mainloop copies spilled and grew by ~100 instructions per chunk).
// Both return the pipeline's chunk counter modulo 2*STAGES (all that stage/parity selection needs), or -1 on abort.
template <bool GEN>
__device__ __forceinline__ int df_diag_task(const PotrfArgs& A, int b, int j, unsigned g_in, unsigned bars) {
    extern __shared__ double smem[];
    __shared__ double pub[3 * TILE];
    __shared__ double svec[TILE];
    __shared__ int s_fail, s_ok;
    Pipe pipe{bars, g_in};
    const int npad = A.npad, nt = npad / TILE;
    double* Lb = A.L + (long long)b * npad * npad;
    double* zb = A.y ? A.z + (long long)b * npad : nullptr;
    int* prog = A.progress + (long long)b * nt;
    int* abort_flag = A.ctl + 1;
    const double* Arow = Lb + ((long long)j * TILE) * npad;
    Seg none{nullptr, 0, nullptr, 0, 0};
    double acc[4][8][2];
    zero_acc(acc);
    GemvHook hook;
    hook.init(zb);
    if (j >= 2) {  // k-blocks 0..j-2 are older than the previous column
        if (!df_wait(prog + j, j - 1, abort_flag, A.info, &s_ok)) return -1;
        Seg s0{Arow, npad, Arow, npad, (j - 1) * (TILE / KC)};
        mainloop<true, XPF_POTRF>(acc, s0, none, smem, hook, pipe);
    }
    if (j >= 1) {  // the last k-block is tile (j, j-1) of the previous column
        if (!df_wait(prog + j, j, abort_flag, A.info, &s_ok)) return -1;
        hook.z = zb ? zb + (long long)(j - 1) * TILE : nullptr;
        const double* Alast = Arow + (long long)(j - 1) * TILE;
        Seg s0{Alast, npad, Alast, npad, TILE / KC};
        mainloop<true, XPF_POTRF>(acc, s0, none, smem, hook, pipe);
    }
    hook.reduce_to(svec, nullptr);
    // with a shared leading block the host pre-loads logdet_half / zz with the communication expert's sums: never "start"
    // start of the expert's sums: block column 0, or — with the communication expert in the same launch — block column gt
    // continuing from its totals; with a precomputed shared block the host pre-loaded them and there is no start
    diag_finish<GEN>(acc, smem, pub, svec, &s_fail, Lb, b, j, A.n, npad, A.a, A.nrm, A.hp, A.hp_stride, A.diag_add, A.dpad, A.invdiag,
                     A.y, zb, A.logdet_half, A.zz, A.info, nullptr, (A.gt == 0 && j == 0) || (A.base_ld != nullptr && j == A.gt), true,
                     A.base_ld, A.base_zz, A.base_info, A.gt, A.ld_blk, A.zz_blk);
    if (A.zbcast && zb && j < A.zbcast_gt) {  // communication expert: hand z_j to every local expert before publishing the task
        __syncthreads();
        if (threadIdx.x < TILE) {
            const double v = zb[j * TILE + threadIdx.x];
            for (int e = 0; e < A.zbcast_B; ++e) __stcg(A.zbcast + (long long)e * A.zbcast_stride + j * TILE + threadIdx.x, v);
        }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) st_release_gpu(prog + j, j + 1);
    return (int)(pipe.g % (2 * STAGES));
}

template <bool GEN>
__device__ __forceinline__ int df_panel_task(const PotrfArgs& A, int b, int i, int j, unsigned g_in, unsigned bars) {
    extern __shared__ double smem[];
    __shared__ int s_ok;
    Pipe pipe{bars, g_in};
    const int tid = threadIdx.x;
    const int npad = A.npad, nt = npad / TILE, n = A.n, dpad = A.dpad;
    double* Lb = A.L + (long long)b * npad * npad;
    int* prog = A.progress + (long long)b * nt;
    int* abort_flag = A.ctl + 1;
    double* Lij = Lb + ((long long)i * TILE) * npad + j * TILE;
    NoHook nh;
    Seg none{nullptr, 0, nullptr, 0, 0};
    double acc[4][8][2];
    zero_acc(acc);
    const bool shared_col = j < A.gt;  // block column of the communication set: L[j, 0:j] and inv(L_jj) come from (Lg, invdiag_g)
    if (j > 0) {
        if (!df_wait(prog + i, j, abort_flag, A.info, &s_ok)) return -1;  // L[i, 0:j]
        if (!shared_col && !df_wait(prog + j, j, abort_flag, A.info, &s_ok)) return -1;  // L[j, 0:j]
        if (shared_col && A.gprog && !df_wait(A.gprog + j, j, abort_flag, A.info, &s_ok)) return -1;  // L_g[j, 0:j], same launch
        const double* Brow = shared_col ? A.Lg + ((long long)j * TILE) * A.npad_g : Lb + ((long long)j * TILE) * npad;
        Seg s0{Lb + ((long long)i * TILE) * npad, npad, Brow, shared_col ? (long long)A.npad_g : (long long)npad, j * (TILE / KC)};
        mainloop<false, XPF_POTRF>(acc, s0, none, smem, nh, pipe);
    }
    if (GEN) {
        FeatSmem f = carve_feat(smem, dpad);
        load_feat_tile(f.sR, f.sld, A.a + ((long long)b * npad + i * TILE) * dpad, dpad);
        load_feat_tile(f.sC, f.sld, A.a + ((long long)b * npad + j * TILE) * dpad, dpad);
        cp_async_commit();
        if (tid < TILE) {
            f.nR[tid] = A.nrm[(long long)b * npad + i * TILE + tid];
            f.nC[tid] = A.nrm[(long long)b * npad + j * TILE + tid];
        }
        cp_async_wait<0>();
        __syncthreads();
        const double sig = A.hp[(long long)b * A.hp_stride];
        gen_se_tile(f.sR, f.sC, f.sld, f.nR, f.nC, dpad, sig * sig, [&](int mi, int ni, int e, double kv) {
            const int gr = i * TILE + acc_row(mi), gc = j * TILE + acc_col(ni, e);
            acc[mi][ni][e] = ((gr < n && gc < n) ? kv : 0.0) - acc[mi][ni][e];
        });
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < 8; ++ni)
                *reinterpret_cast<double2*>(Lij + (long long)acc_row(mi) * npad + acc_col(ni, 0)) =
                    make_double2(acc[mi][ni][0], acc[mi][ni][1]);
    } else {
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < 8; ++ni) {
                double2* p = reinterpret_cast<double2*>(Lij + (long long)acc_row(mi) * npad + acc_col(ni, 0));
                double2 k2 = *p;
                k2.x -= acc[mi][ni][0];
                k2.y -= acc[mi][ni][1];
                *p = k2;
            }
    }
    __threadfence_block();
    if (shared_col && !A.gprog) __syncthreads();  // publishes X to the CTA (df_wait's barrier does it otherwise)
    else if (!df_wait(shared_col ? A.gprog + j : prog + j, j + 1, abort_flag, A.info, &s_ok)) return -1;  // inv(L_jj)
    zero_acc(acc);
    const double* invj = shared_col ? A.invdiag_g + (long long)j * TILE * TILE : A.invdiag + ((long long)b * nt + j) * TILE * TILE;
    Seg s1{Lij, npad, invj, TILE, TILE / KC};
    mainloop<false, XPF_POTRF>(acc, s1, none, smem, nh, pipe);
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 8; ++ni)
            *reinterpret_cast<double2*>(Lij + (long long)acc_row(mi) * npad + acc_col(ni, 0)) =
                make_double2(acc[mi][ni][0], acc[mi][ni][1]);
    __threadfence();
    __syncthreads();
    if (tid == 0) st_release_gpu(prog + i, j + 1);
    return (int)(pipe.g % (2 * STAGES));
}

// One CTA per task, NOT a persistent loop: the task index is a ticket drawn at CTA start (so "waits only on tickets held by
// running CTAs" holds whatever order the hardware dispatches blocks in), and the body is straight-line code like the
// per-column kernels (inside a persistent loop ptxas spilled into all four mainloop copies, +100 instructions per chunk).
template <bool GEN>
__global__ void __launch_bounds__(NTHREADS, 1)
potrf_dataflow_kernel(const __grid_constant__ PotrfArgs A, const __grid_constant__ PotrfArgs G) {
    __shared__ unsigned long long pipe_bars[2 * STAGES];
    __shared__ int s_col, s_local;
    Pipe pipe;
    // communication expert in this launch (G): its own ntg - col tile tasks lead block column col (ntg >= gt block columns:
    // one more than the shared gt when its size is not a multiple of 128)
    const int ntg = A.gprog ? G.npad / TILE : 0;
    if (threadIdx.x == 0) {
        const int B = A.B, nt = A.npad / TILE, gt = A.gt;
        const int task = atomicAdd(A.ctl, 1);
        // tasks of block column `col`: B diagonal tasks (col >= gt only) then B * (number of rows below max(col, gt-1)) panel tiles
        int col = 0, col_start = 0;
        for (;;) {
            const int ntask = (col < ntg ? ntg - col : 0) + (col < gt ? B * (nt - gt) : B * (nt - col));
            if (task < col_start + ntask) break;
            col_start += ntask;
            ++col;
        }
        s_col = col;
        s_local = task - col_start;
    }
    pipe_init(pipe, pipe_bars);  // ends with a CTA barrier
    const int col = s_col, nt = A.npad / TILE;
    int local = s_local;
    if (col < ntg) {  // the communication expert's own tasks of this column: its diagonal tile, then its panel tiles
        if (local == 0) {
            df_diag_task<GEN>(G, 0, col, 0u, pipe.bars);
            return;
        }
        if (local < ntg - col) {
            df_panel_task<GEN>(G, 0, col + local, col, 0u, pipe.bars);
            return;
        }
        local -= ntg - col;
    }
    if (col < A.gt) {  // shared block column: panel tiles of the rows below the shared block only
        const int per = nt - A.gt;
        df_panel_task<GEN>(A, local / per, A.gt + local % per, col, 0u, pipe.bars);
    } else if (local < A.B) {
        df_diag_task<GEN>(A, local, col, 0u, pipe.bars);
    } else {
        const int per = nt - 1 - col;
        df_panel_task<GEN>(A, (local - A.B) / per, col + 1 + (local - A.B) % per, col, 0u, pipe.bars);
    }
}

// Split factor for a launch of `tiles` equal-depth tiles of `depth` k-blocks.  Two effects are traded: (i) a launch with
// fewer tiles than SMs leaves SMs idle, (ii) a launch of a few waves wastes up to one wave in the ragged tail.  Cost
// model in tile units (one 128^3 product): waves * (depth/S + epilogue) + fix-up; S = 1..8, slices keep >= 1 block.
static int choose_split(long long tiles, int depth, int sms, long long slot_cap) {
    if (depth < 2 || tiles >= 8LL * sms) return 1;
    int best = 1;
    double best_cost = 1e300;
    for (int S = 1; S <= 8 && S <= depth; ++S) {
        if (S > 1 && tiles * S > slot_cap) break;
        const double waves = (double)((tiles * S + sms - 1) / sms);
        const double cost = waves * ((double)depth / S + 1.1) + (S > 1 ? 0.4 + 0.08 * S : 0.0);
        if (cost < best_cost - 1e-9) { best_cost = cost; best = S; }
    }
    if (tiles < sms) {  // far below one wave: allow deeper splits, ~2 CTAs per SM
        long long s = (2LL * sms) / tiles;
        if (s > depth) s = depth;
        if (s > 32) s = 32;
        if (s > best && tiles * s <= slot_cap) best = (int)s;
    }
    return best;
}
static inline long long align256(long long v) { return (v + 255) / 256 * 256; }
// workspace layout: [counters: B*nt ints][gemv slots: B*32*128 doubles][tile slots: 12*sms + B*nt tiles of 128x128 doubles]
static long long potrf_slot_tiles(int B, int nt, int sms) { return 12LL * sms + (long long)B * nt; }
static long long potrf_chain_slot_tiles(int sms) { return 6LL * sms; }

extern "C" long long gpr_potrf_work_bytes(int B, int npad) {
    int sms = gpr_device_sm_count();
    if (sms <= 0) sms = 148;
    const int nt = npad / TILE;
    return 256 + align256((long long)B * nt * 4) + align256((long long)B * 32 * TILE * 8) + potrf_slot_tiles(B, nt, sms) * TILE * TILE * 8 +
           align256((long long)B * nt * 4) + potrf_chain_slot_tiles(sms) * TILE * TILE * 8 +   // + the chain launches' own counters / slots
           align256((long long)(B + 1) * 4);                                                  // + abort flag and per-expert "diagonal tile done"
}

// Internal streams / events of the two-stream per-column schedule (created once per device; no device memory)
struct PotrfAux {
    cudaStream_t chain = nullptr, chain2 = nullptr, bulk = nullptr, bulk_hi = nullptr;
    cudaEvent_t fork = nullptr, join_chain = nullptr, join_chain2 = nullptr, join_bulk = nullptr, diag[4] = {}, bulk_done[4] = {},
                first_done[4] = {};
    bool ready = false;
};
static PotrfAux g_potrf_aux[64];
static std::mutex g_potrf_aux_mutex, g_potrf_enqueue_mutex;
static PotrfAux* potrf_aux() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lock(g_potrf_aux_mutex);
    PotrfAux& x = g_potrf_aux[dev];
    if (!x.ready) {
        int lo = 0, hi = 0;
        if (cudaDeviceGetStreamPriorityRange(&lo, &hi) != cudaSuccess) return nullptr;
        if (cudaStreamCreateWithPriority(&x.chain, cudaStreamNonBlocking, hi) != cudaSuccess) return nullptr;   // hi = highest priority
        if (cudaStreamCreateWithPriority(&x.chain2, cudaStreamNonBlocking, hi) != cudaSuccess) return nullptr;
        if (cudaStreamCreateWithPriority(&x.bulk, cudaStreamNonBlocking, lo) != cudaSuccess) return nullptr;
        // a second bulk stream one step above the lowest priority, used when the caller's own stream is a high-priority one: a
        // factorisation enqueued under someone else's tensor-core-bound launch (engine.fit_emulated's look-ahead) must not queue its
        // panels behind that launch's thousands of default-priority CTAs
        if (cudaStreamCreateWithPriority(&x.bulk_hi, cudaStreamNonBlocking, (hi < lo - 1) ? lo - 1 : lo) != cudaSuccess) return nullptr;
        cudaEvent_t* evs[] = {&x.fork, &x.join_chain, &x.join_chain2, &x.join_bulk, &x.diag[0], &x.diag[1], &x.diag[2], &x.diag[3],
                              &x.bulk_done[0], &x.bulk_done[1], &x.bulk_done[2], &x.bulk_done[3],
                              &x.first_done[0], &x.first_done[1], &x.first_done[2], &x.first_done[3]};
        for (cudaEvent_t* e : evs)
            if (cudaEventCreateWithFlags(e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        x.ready = true;
    }
    return &x;
}

static int potrf_impl(double* L, int gen, const double* a, const double* nrm, const double* hp, int hp_stride,
                      const double* diag_add, int B, int n, int npad, int dpad, double* invdiag, const double* y, double* z,
                      double* logdet_half, double* zz, int* info, void* work, void* stream, int gt, const double* Lg, int npad_g,
                      const double* invdiag_g) {
    ARG_CHECK(L && invdiag);
    ARG_CHECK(B > 0 && n > 0 && npad % TILE == 0 && npad >= n);
    ARG_CHECK(!gen || (a && nrm && hp && dpad > 0 && dpad % 4 == 0));
    ARG_CHECK(!y || z);
    ARG_CHECK(!zz || y);
    cudaStream_t st = (cudaStream_t)stream;
    const int nt = npad / TILE;
    int sms = gpr_device_sm_count();
    if (sms <= 0) return fail(GPR_ERR_CUDA, "no CUDA device%s", "");
    int* counters = nullptr;
    double* gemv = nullptr;
    double* slots = nullptr;
    if (work) {  // without a workspace the factorisation runs unsplit
        counters = reinterpret_cast<int*>(work);
        gemv = reinterpret_cast<double*>(reinterpret_cast<char*>(work) + align256((long long)B * nt * 4));
        slots = reinterpret_cast<double*>(reinterpret_cast<char*>(gemv) + align256((long long)B * 32 * TILE * 8));
        CUDA_TRY(cudaMemsetAsync(counters, 0, (size_t)B * nt * 4, st));
    }
    // single-launch dataflow form when there are enough tile tasks per block column to keep the SMs busy across the
    // dependency chain of any one expert; the per-column form with split-K for few experts (PYGPR_POTRF_MODE overrides)
    {
        const int mode = g_potrf_mode;
        const long long total = (long long)B * ((long long)gt * (nt - gt) + (long long)(nt - gt) * (nt - gt + 1) / 2);
        // measured on B200 (tools/potrf_modes.py): dataflow wins from ~8 experts up (+8 % at B=32..63), ties at 4, loses below
        // (three-stream per-column form: B=1 n=16384 59 ms vs 145 ms dataflow, B=3 n=2048 2.6 vs 3.1 ms, B=1 n=1024 a tie)
        bool dataflow = work && nt >= 2 && B >= 6 && total < (1LL << 30);
        if (mode == POTRF_COLUMNS || mode == POTRF_COLUMNS1) dataflow = false;
        if (mode == POTRF_DATAFLOW && work && total < (1LL << 30)) dataflow = true;
        if (gt > 0) {  // the shared leading block exists in the dataflow form only
            if (!work || total >= (1LL << 30)) return fail(GPR_ERR_BAD_ARG, "shared leading block needs a workspace%s", "");
            dataflow = true;
        }
        if (dataflow) {
            const int smem_df = diag_smem_bytes(gen, dpad);
            if (smem_df > 226 * 1024)
                return fail(GPR_ERR_UNSUPPORTED, "input dimension too large for the fused factorisation%s (dpad=%lld)", "", dpad);
            int* ctl = reinterpret_cast<int*>(work);
            PotrfArgs A{L, B, npad, n, a, nrm, hp, hp_stride, diag_add, dpad, invdiag, y, z, logdet_half, zz, info, ctl, ctl + 64,
                        gt, Lg, npad_g, invdiag_g, nullptr, nullptr, 0, 0, 0, nullptr, nullptr, nullptr, nullptr, nullptr};
            CUDA_TRY(cudaMemsetAsync(work, 0, 256 + (size_t)B * nt * 4, st));
            const int grid = (int)total;  // one CTA per tile task
            if (gen) {
                CUDA_TRY(opt_in_smem(potrf_dataflow_kernel<true>, smem_df));
                potrf_dataflow_kernel<true><<<grid, NTHREADS, smem_df, st>>>(A, A);
            } else {
                CUDA_TRY(opt_in_smem(potrf_dataflow_kernel<false>, smem_df));
                potrf_dataflow_kernel<false><<<grid, NTHREADS, smem_df, st>>>(A, A);
            }
            if (info) report_stall_kernel<<<1, 128, 0, st>>>(ctl + 1, info, B, nullptr, 0);
            CUDA_TRY(cudaGetLastError());
            return GPR_OK;
        }
    }
    int smem_diag = diag_smem_bytes(gen, dpad), smem_panel = MAIN_SMEM_BYTES;
    if (smem_diag > 226 * 1024)
        return fail(GPR_ERR_UNSUPPORTED, "input dimension too large for the fused factorisation%s (dpad=%lld)", "", dpad);
    if (gen) {
        const int fb = feat_smem_bytes(dpad);
        if (fb > smem_panel) smem_panel = fb;
        CUDA_TRY(opt_in_smem(potrf_diag_kernel<true>, smem_diag));
        CUDA_TRY(opt_in_smem(potrf_panel_kernel<true>, smem_panel));
    } else {
        CUDA_TRY(opt_in_smem(potrf_diag_kernel<false>, smem_diag));
        CUDA_TRY(opt_in_smem(potrf_panel_kernel<false>, smem_panel));
    }
    // Per-column schedule.  With a workspace it runs on three internal streams: the diagonal tiles (latency bound: the
    // factorisation + inversion of the tile was ~70 us per column in round 1, ~31 us since gpr_diag_tile.cuh) and the FIRST panel tile + look-ahead product of each column on two
    // high-priority streams — the first tile's partial products run alongside the diagonal tile of its own column and only its
    // solve waits, on a flag, for inv(L_jj) — and the BULK of each panel (row tiles j+2 ...) on a low-priority one that only
    // needs diag(j).  A column then costs max(bulk, diag + solve) instead of diag + first tile + bulk
    // (one matrix: n = 4096 8.4 -> 5.3 ms, 8192 22 -> 14.5, 16384 79 -> 59; 4 x 8192: 40 -> 30).
    // PYGPR_POTRF_MODE=columns1 keeps everything on the caller's stream.
    PotrfAux* aux = (work && nt >= 3 && g_potrf_mode != POTRF_COLUMNS1) ? potrf_aux() : nullptr;
    int* counters_c = nullptr;
    double* slots_c = nullptr;
    int* dflags = nullptr;  // [0] abort, [1 + b] = number of diagonal tiles of expert b finished
    // the internal streams and events are per device: one factorisation at a time may be ENQUEUED through them
    std::unique_lock<std::mutex> enqueue_lock(g_potrf_enqueue_mutex, std::defer_lock);
    if (aux) enqueue_lock.lock();
    cudaStream_t bulk_stream = nullptr;
    if (aux) {
        int lo_p = 0, hi_p = 0, st_p = 0;
        bulk_stream = aux->bulk;
        if (cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p) == cudaSuccess && cudaStreamGetPriority(st, &st_p) == cudaSuccess && st_p < lo_p)
            bulk_stream = aux->bulk_hi;   // the caller runs this factorisation on a high-priority stream (under another launch)
    }
    if (aux) {
        char* p = reinterpret_cast<char*>(slots) + potrf_slot_tiles(B, nt, sms) * TILE * TILE * 8;
        counters_c = reinterpret_cast<int*>(p);
        slots_c = reinterpret_cast<double*>(p + align256((long long)B * nt * 4));
        dflags = reinterpret_cast<int*>(reinterpret_cast<char*>(slots_c) + potrf_chain_slot_tiles(sms) * TILE * TILE * 8);
        CUDA_TRY(cudaMemsetAsync(counters_c, 0, (size_t)B * nt * 4, st));
        CUDA_TRY(cudaMemsetAsync(dflags, 0, (size_t)(B + 1) * 4, st));
        CUDA_TRY(cudaEventRecord(aux->fork, st));
        CUDA_TRY(cudaStreamWaitEvent(aux->chain, aux->fork, 0));
        CUDA_TRY(cudaStreamWaitEvent(aux->chain2, aux->fork, 0));
        CUDA_TRY(cudaStreamWaitEvent(bulk_stream, aux->fork, 0));
    }
    // sd: diagonal tiles; sf: first panel tile + look-ahead product (runs ALONGSIDE the diagonal tile of its column: only its
    // solve waits, on a flag, for inv(L_jj)); sb: the bulk of the panel
    cudaStream_t sd = aux ? aux->chain : st, sf = aux ? aux->chain2 : st, sb = aux ? bulk_stream : st;
    auto launch_panel = [&](cudaStream_t s_, int j, int row0, int ntile_rows, int lookahead, int S, double* sl, int* cnt, int* dflag) {
        dim3 grid((ntile_rows + lookahead) * S, B);
        if (gen)
            potrf_panel_kernel<true><<<grid, NTHREADS, smem_panel, s_>>>(L, npad, n, j, a, nrm, hp, hp_stride, dpad, invdiag, lookahead,
                                                                         y ? z : nullptr, S, sl, gemv, cnt, row0, dflag);
        else
            potrf_panel_kernel<false><<<grid, NTHREADS, smem_panel, s_>>>(L, npad, n, j, a, nrm, hp, hp_stride, dpad, invdiag, lookahead,
                                                                          y ? z : nullptr, S, sl, gemv, cnt, row0, dflag);
    };
    bool bulk_prev = false;  // column j-1 had a bulk launch
    for (int j = 0; j < nt; ++j) {
        if (aux && j >= 1) CUDA_TRY(cudaStreamWaitEvent(sd, aux->first_done[(j - 1) & 3], 0));   // tile (j, j-1) and the look-ahead sums
        int* dflag = aux ? dflags + 1 : nullptr;
        if (gen)
            potrf_diag_kernel<true><<<B, NTHREADS, smem_diag, sd>>>(L, npad, n, j, a, nrm, hp, hp_stride, diag_add, dpad, invdiag,
                                                                     y, z, logdet_half, zz, info, dflag);
        else
            potrf_diag_kernel<false><<<B, NTHREADS, smem_diag, sd>>>(L, npad, n, j, a, nrm, hp, hp_stride, diag_add, dpad,
                                                                      invdiag, y, z, logdet_half, zz, info, dflag);
        if (j + 1 >= nt) break;
        const int lookahead = (j >= 1) ? 1 : 0;
        if (!aux) {
            const int ntile_x = nt - 1 - j + lookahead;
            const int S = work ? choose_split((long long)B * ntile_x, j, sms, potrf_slot_tiles(B, nt, sms)) : 1;
            launch_panel(st, j, 0, nt - 1 - j, lookahead, S, slots, counters, nullptr);
            continue;
        }
        const int rows_bulk = nt - 2 - j;
        CUDA_TRY(cudaEventRecord(aux->diag[j & 3], sd));
        if (rows_bulk > 0) {
            CUDA_TRY(cudaStreamWaitEvent(sb, aux->diag[j & 3], 0));
            const int S = choose_split((long long)B * rows_bulk, j, sms, potrf_slot_tiles(B, nt, sms));
            launch_panel(sb, j, 1, rows_bulk, 0, S, slots, counters, nullptr);
            CUDA_TRY(cudaEventRecord(aux->bulk_done[j & 3], sb));
        }
        if (bulk_prev) CUDA_TRY(cudaStreamWaitEvent(sf, aux->bulk_done[(j - 1) & 3], 0));   // row j+1 of column j-1
        {
            const int S = choose_split((long long)B * (1 + lookahead), j, sms, potrf_chain_slot_tiles(sms));
            launch_panel(sf, j, 0, 1, lookahead, S, slots_c, counters_c, dflag);
            CUDA_TRY(cudaEventRecord(aux->first_done[j & 3], sf));
        }
        bulk_prev = rows_bulk > 0;
    }
    if (aux) {
        CUDA_TRY(cudaEventRecord(aux->join_chain, sd));
        CUDA_TRY(cudaEventRecord(aux->join_chain2, sf));
        CUDA_TRY(cudaEventRecord(aux->join_bulk, sb));
        CUDA_TRY(cudaStreamWaitEvent(st, aux->join_chain, 0));
        CUDA_TRY(cudaStreamWaitEvent(st, aux->join_chain2, 0));
        CUDA_TRY(cudaStreamWaitEvent(st, aux->join_bulk, 0));
    }
    CUDA_TRY(cudaGetLastError());
    return GPR_OK;
}

extern "C" int gpr_potrf_batched(double* L, int gen, const double* a, const double* nrm, const double* hp, int hp_stride,
                                 const double* diag_add, int B, int n, int npad, int dpad, double* invdiag, const double* y,
                                 double* z, double* logdet_half, double* zz, int* info, void* work, void* stream) {
    return potrf_impl(L, gen, a, nrm, hp, hp_stride, diag_add, B, n, npad, dpad, invdiag, y, z, logdet_half, zz, info, work, stream, 0,
                      nullptr, 0, nullptr);
}

extern "C" int gpr_potrf_batched_shared(double* L, const double* a, const double* nrm, const double* hp, int hp_stride,
                                        const double* diag_add, int B, int n, int npad, int dpad, double* invdiag, const double* y,
                                        double* z, double* logdet_half, double* zz, int* info, void* work, void* stream, int ng,
                                        const double* Lg, int npad_g, const double* invdiag_g) {
    ARG_CHECK(Lg && invdiag_g && work && ng > 0 && ng % TILE == 0 && ng < n && npad_g >= ng && npad_g % TILE == 0);
    ARG_CHECK(y && z && logdet_half && zz && info);
    return potrf_impl(L, 1, a, nrm, hp, hp_stride, diag_add, B, n, npad, dpad, invdiag, y, z, logdet_half, zz, info, work, stream,
                      ng / TILE, Lg, npad_g, invdiag_g);
}

// Whole GRBCM fit in one launch: the communication expert (xg, yg) and B local experts that start with its ng rows under its
// hyper-parameters (see gpr_potrf_batched_shared; here the shared factor is produced by the same launch, so its serial chain
// of ng/128 block columns overlaps the local experts' tiles instead of preceding them).
extern "C" long long gpr_potrf_grbcm_work_bytes(int B, int npad, int npad_g) {
    return 256 + align256((long long)B * (npad / TILE) * 4) + align256((long long)(npad_g / TILE) * 4) +
           2 * align256((long long)(npad_g / TILE) * 8);
}

extern "C" int gpr_potrf_grbcm(double* L, const double* a, const double* nrm, const double* hp, int hp_stride, const double* diag_add,
                               int B, int n, int npad, int dpad, double* invdiag, const double* y, double* z, double* logdet_half,
                               double* zz, int* info, double* Lg, const double* a_g, const double* nrm_g, const double* hp_g,
                               const double* diag_add_g, int ng, int npad_g, double* invdiag_g, const double* y_g, double* z_g,
                               double* logdet_half_g, double* zz_g, int* info_g, void* work, void* stream) {
    ARG_CHECK(L && a && nrm && hp && invdiag && y && z && logdet_half && zz && info && work);
    ARG_CHECK(Lg && a_g && nrm_g && hp_g && invdiag_g && y_g && z_g && logdet_half_g && zz_g && info_g);
    ARG_CHECK(B > 0 && n > 0 && npad % TILE == 0 && npad >= n && dpad > 0 && dpad % 4 == 0);
    ARG_CHECK(ng >= TILE && ng < n && npad_g % TILE == 0 && npad_g >= ng && npad_g - ng < TILE);
    cudaStream_t st = (cudaStream_t)stream;
    // gt whole 128-row blocks of the communication set are shared; a ragged last block of it is factorised by the communication
    // expert for itself and, as part of their own rows, by every local expert
    const int nt = npad / TILE, gt = ng / TILE, ntg = npad_g / TILE;
    const long long total = (long long)ntg * (ntg + 1) / 2 + (long long)B * ((long long)gt * (nt - gt) + (long long)(nt - gt) * (nt - gt + 1) / 2);
    if (total >= (1LL << 30)) return fail(GPR_ERR_UNSUPPORTED, "too many tile tasks for one launch%s (%lld)", "", total);
    const int smem_df = diag_smem_bytes(1, dpad);
    if (smem_df > 226 * 1024)
        return fail(GPR_ERR_UNSUPPORTED, "input dimension too large for the fused factorisation%s (dpad=%lld)", "", dpad);
    int* ctl = reinterpret_cast<int*>(work);
    int* progress = ctl + 64;
    char* wp = reinterpret_cast<char*>(work) + 256 + align256((long long)B * nt * 4);
    int* gprog = reinterpret_cast<int*>(wp);
    double* ld_blk_g = reinterpret_cast<double*>(wp + align256((long long)ntg * 4));
    double* zz_blk_g = ld_blk_g + align256((long long)ntg * 8) / 8;
    CUDA_TRY(cudaMemsetAsync(work, 0, (size_t)gpr_potrf_grbcm_work_bytes(B, npad, npad_g), st));
    PotrfArgs A{L, B, npad, n, a, nrm, hp, hp_stride, diag_add, dpad, invdiag, y, z, logdet_half, zz, info, ctl, progress,
                gt, Lg, npad_g, invdiag_g, gprog, nullptr, 0, 0, 0, ld_blk_g, zz_blk_g, info_g, nullptr, nullptr};
    PotrfArgs G{Lg, 1, npad_g, ng, a_g, nrm_g, hp_g, 0, diag_add_g, dpad, invdiag_g, y_g, z_g, logdet_half_g, zz_g, info_g, ctl, gprog,
                0, nullptr, 0, nullptr, nullptr, z, B, npad, gt, nullptr, nullptr, nullptr, ld_blk_g, zz_blk_g};
    CUDA_TRY(opt_in_smem(potrf_dataflow_kernel<true>, smem_df));
    potrf_dataflow_kernel<true><<<(int)total, NTHREADS, smem_df, st>>>(A, G);
    if (info) report_stall_kernel<<<1, 128, 0, st>>>(ctl + 1, info, B, info_g, 1);
    CUDA_TRY(cudaGetLastError());
    return GPR_OK;
}

// sum_{i < rows} log L_ii and sum_{i < rows} z_i^2 of a finished factorisation: what gpr_potrf_batched_shared continues from when
// only the first `rows` rows of the communication expert are shared (out[0], out[1]).
__global__ void factor_partials_kernel(const double* __restrict__ L, int npad, const double* __restrict__ z, int rows, double* __restrict__ out) {
    __shared__ double sl[NTHREADS / 32], sz[NTHREADS / 32];
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < rows; i += NTHREADS) {
        a += log(L[(long long)i * npad + i]);
        const double v = z[i];
        b += v * v;
    }
    a = warp_sum(a);
    b = warp_sum(b);
    if ((threadIdx.x & 31) == 0) { sl[threadIdx.x >> 5] = a; sz[threadIdx.x >> 5] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double x = 0.0, y = 0.0;
        for (int w = 0; w < NTHREADS / 32; ++w) { x += sl[w]; y += sz[w]; }
        out[0] = x;
        out[1] = y;
    }
}

extern "C" int gpr_factor_partials(const double* L, int npad, const double* z, int rows, double* out, void* stream) {
    ARG_CHECK(L && z && out && rows > 0 && rows <= npad);
    factor_partials_kernel<<<1, NTHREADS, 0, (cudaStream_t)stream>>>(L, npad, z, rows, out);
    CUDA_TRY(cudaGetLastError());
    return GPR_OK;
}

// Replicate the communication expert's leading factor block into every expert's buffer (128 x 128 tiles; lower_only skips the
// tiles above the block diagonal, which are only needed when they carry L_g^-T for gpr_lml_grad_shared).
__global__ void __launch_bounds__(NTHREADS) copy_blocks_kernel(const double* __restrict__ src, long long src_ld, double* __restrict__ dst,
                                                               long long dst_ld, long long dst_bstride, int lower_only) {
    const int tc = blockIdx.x, tr = blockIdx.y;
    if (lower_only == 1 && tc > tr) return;   // 1: tiles on / below the block diagonal
    if (lower_only == 2 && tc <= tr) return;  // 2: only the tiles above it (L_g^-T, fetched later for the shared-block gradient)
    const double* s = src + (long long)tr * TILE * src_ld + (long long)tc * TILE;
    double* d = dst + (long long)blockIdx.z * dst_bstride + (long long)tr * TILE * dst_ld + (long long)tc * TILE;
    for (int id = threadIdx.x; id < TILE * (TILE / 2); id += NTHREADS) {
        const int r = id / (TILE / 2), c2 = (id % (TILE / 2)) * 2;
        __stcs(reinterpret_cast<double2*>(d + (long long)r * dst_ld + c2), __ldg(reinterpret_cast<const double2*>(s + (long long)r * src_ld + c2)));
    }
}

extern "C" int gpr_copy_leading_block(double* L, int B, int npad, const double* Lg, int npad_g, int ng, int lower_only, void* stream) {
    ARG_CHECK(L && Lg && B > 0 && ng > 0 && ng % TILE == 0 && ng <= npad && ng <= npad_g && npad % 2 == 0 && npad_g % 2 == 0);
    copy_blocks_kernel<<<dim3(ng / TILE, ng / TILE, B), NTHREADS, 0, (cudaStream_t)stream>>>(Lg, npad_g, L, npad, (long long)npad * npad,
                                                                                            lower_only);
    CUDA_TRY(cudaGetLastError());
    return GPR_OK;
}

// ---------------------------------------------------------------------------------------------
// K4: alpha = L^-T z.  HBM-bound (reads the lower triangle of L once); one CTA per expert walks the block
// columns from last to first:  alpha_j = inv(L_jj)^T (z_j - sum_{i>j} L_ij^T alpha_i).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHREADS) trsv_backward_kernel(const double* __restrict__ L, const double* __restrict__ invdiag,
                                                                 int npad, const double* __restrict__ z,
                                                                 double* __restrict__ alpha) {
    extern __shared__ double sal[];  // alpha for this expert [npad]
    __shared__ double part[NTHREADS / 32][TILE];
    __shared__ double rhs[TILE];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nt = npad / TILE;
    const double* Lb = L + (long long)b * npad * npad;
    for (int j = nt - 1; j >= 0; --j) {
        // s[c] = sum_{r >= (j+1)*128} L[r, j*128 + c] * alpha[r]; lane owns columns 4*lane..4*lane+3
        double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
        const double* col = Lb + (long long)j * TILE + 4 * lane;
        const int r0 = (j + 1) * TILE;
        int r = r0 + warp;
        for (; r + 3 * 8 < npad; r += 4 * 8) {
            double2 v[4][2];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const double* p = col + (long long)(r + u * 8) * npad;
                v[u][0] = *reinterpret_cast<const double2*>(p);
                v[u][1] = *reinterpret_cast<const double2*>(p + 2);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const double al = sal[r + u * 8];
                s0 += v[u][0].x * al; s1 += v[u][0].y * al; s2 += v[u][1].x * al; s3 += v[u][1].y * al;
            }
        }
        for (; r < npad; r += 8) {
            const double* p = col + (long long)r * npad;
            const double2 v0 = *reinterpret_cast<const double2*>(p), v1 = *reinterpret_cast<const double2*>(p + 2);
            const double al = sal[r];
            s0 += v0.x * al; s1 += v0.y * al; s2 += v1.x * al; s3 += v1.y * al;
        }
        part[warp][4 * lane + 0] = s0; part[warp][4 * lane + 1] = s1;
        part[warp][4 * lane + 2] = s2; part[warp][4 * lane + 3] = s3;
        __syncthreads();
        if (tid < TILE) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < NTHREADS / 32; ++w) s += part[w][tid];
            rhs[tid] = z[(long long)b * npad + j * TILE + tid] - s;
        }
        __syncthreads();
        // alpha_j[c] = sum_{k >= c} inv[k][c] * rhs[k]     (inv lower triangular, read coalesced by rows)
        const double* inv = invdiag + ((long long)b * nt + j) * TILE * TILE;
        s0 = s1 = s2 = s3 = 0;
        for (int k = warp; k < TILE; k += 8) {
            const double* p = inv + k * TILE + 4 * lane;
            const double2 v0 = *reinterpret_cast<const double2*>(p), v1 = *reinterpret_cast<const double2*>(p + 2);
            const double rk = rhs[k];
            s0 += v0.x * rk; s1 += v0.y * rk; s2 += v1.x * rk; s3 += v1.y * rk;
        }
        part[warp][4 * lane + 0] = s0; part[warp][4 * lane + 1] = s1;
        part[warp][4 * lane + 2] = s2; part[warp][4 * lane + 3] = s3;
        __syncthreads();
        if (tid < TILE) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < NTHREADS / 32; ++w) s += part[w][tid];
            sal[j * TILE + tid] = s;
            alpha[(long long)b * npad + j * TILE + tid] = s;
        }
        __syncthreads();
    }
}

// Few experts: one CTA per expert cannot pull more than a few % of HBM bandwidth (B=1, n=16384: 35 ms for 1 GB).
// Column-owner form with G CTAs per expert: CTA c owns the block columns j = c (mod G) and keeps their running
// right-hand sides  r_j = z_j - sum_{i>j, alpha_i known} L_ij^T alpha_i  in shared memory.  As soon as alpha_i is
// published (per-expert counter, release/acquire) every owner subtracts its tiles of block row i; the owner of block
// i-1 does that tile first, solves  alpha_{i-1} = inv(L_{i-1,i-1})^T r_{i-1}  and publishes before touching the rest.
// The CTAs of one expert wait for each other, so the launch is cooperative (all CTAs co-resident or the launch fails).
__global__ void __launch_bounds__(NTHREADS)
trsv_backward_owner_kernel(const double* __restrict__ L, const double* __restrict__ invdiag, int npad, const double* __restrict__ z,
                           double* __restrict__ alpha, int* __restrict__ done, int* __restrict__ abort_flag) {
    extern __shared__ double srhs[];  // [owned blocks][128] running right-hand sides
    __shared__ double part[NTHREADS / 32][TILE];
    __shared__ double sal[TILE];
    __shared__ int s_ok;
    const int c = blockIdx.x, G = gridDim.x, b = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nt = npad / TILE;
    const double* Lb = L + (long long)b * npad * npad;
    const double* zb = z + (long long)b * npad;
    double* ab = alpha + (long long)b * npad;
    int* flag = done + b;  // number of alpha blocks published so far (blocks nt-1, nt-2, ... in that order)
    for (int j = c; j < nt; j += G)
        if (tid < TILE) srhs[(j / G) * TILE + tid] = zb[j * TILE + tid];
    __syncthreads();

    // alpha_j = inv(L_jj)^T r_j : alpha_j[col] = sum_{k >= col} inv[k][col] r_j[k]; publish
    auto solve_publish = [&](int j) {
        const double* inv = invdiag + ((long long)b * nt + j) * TILE * TILE;
        const double* r = srhs + (j / G) * TILE;
        double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
        for (int k = warp; k < TILE; k += NTHREADS / 32) {
            const double* p = inv + k * TILE + 4 * lane;
            const double2 v0 = *reinterpret_cast<const double2*>(p), v1 = *reinterpret_cast<const double2*>(p + 2);
            const double rk = r[k];
            s0 += v0.x * rk; s1 += v0.y * rk; s2 += v1.x * rk; s3 += v1.y * rk;
        }
        part[warp][4 * lane + 0] = s0; part[warp][4 * lane + 1] = s1;
        part[warp][4 * lane + 2] = s2; part[warp][4 * lane + 3] = s3;
        __syncthreads();
        if (tid < TILE) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < NTHREADS / 32; ++w) s += part[w][tid];
            __stcg(ab + j * TILE + tid, s);
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) st_release_gpu(flag, nt - j);
    };

    if ((nt - 1) % G == c) solve_publish(nt - 1);
    for (int i = nt - 1; i >= 1; --i) {
        if (c >= i) break;  // no owned block column left of block row i
        if (!df_wait(flag, nt - i, abort_flag, nullptr, &s_ok)) return;
        if (tid < TILE) sal[tid] = __ldcg(ab + i * TILE + tid);
        __syncthreads();
        // owned j < i, descending: j = i-1 first when it is ours (it is on the critical path)
        int j = i - 1 - ((i - 1 - c) % G + G) % G;  // largest j <= i-1 with j = c (mod G)
        for (; j >= 0; j -= G) {
            // s[col] = sum_r L[i*128 + r, j*128 + col] alpha_i[r]; lane owns 4 columns, warps stride the rows
            double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
            const double* tile = Lb + ((long long)i * TILE) * npad + (long long)j * TILE + 4 * lane;
#pragma unroll 4
            for (int r = warp; r < TILE; r += NTHREADS / 32) {
                const double* p = tile + (long long)r * npad;
                const double2 v0 = __ldcg(reinterpret_cast<const double2*>(p)), v1 = __ldcg(reinterpret_cast<const double2*>(p + 2));
                const double al = sal[r];
                s0 += v0.x * al; s1 += v0.y * al; s2 += v1.x * al; s3 += v1.y * al;
            }
            part[warp][4 * lane + 0] = s0; part[warp][4 * lane + 1] = s1;
            part[warp][4 * lane + 2] = s2; part[warp][4 * lane + 3] = s3;
            __syncthreads();
            if (tid < TILE) {
                double s = 0.0;
#pragma unroll
                for (int w = 0; w < NTHREADS / 32; ++w) s += part[w][tid];
                srhs[(j / G) * TILE + tid] -= s;
            }
            __syncthreads();
            if (j == i - 1) solve_publish(j);
        }
    }
}

extern "C" long long gpr_trsv_work_bytes(int B) { return 256 + align256((long long)B * 4); }

extern "C" int gpr_trsv_backward(const double* L, const double* invdiag, int B, int npad, const double* z, double* alpha,
                                 void* work, void* stream) {
    ARG_CHECK(L && invdiag && z && alpha && B > 0 && npad > 0 && npad % TILE == 0);
    cudaStream_t st = (cudaStream_t)stream;
    const int nt = npad / TILE;
    int sms = gpr_device_sm_count();
    if (sms <= 0) return fail(GPR_ERR_CUDA, "no CUDA device%s", "");
    bool owner = work && B < sms && nt >= 4 && z != alpha;
    if (g_trsv_mode == TRSV_SINGLE) owner = false;
    if (owner) {
        int G = (2 * sms) / B;
        if (G > nt) G = nt;
        const int own = (nt + G - 1) / G;
        const size_t smem = (size_t)own * TILE * 8;
        int dev = 0, per_sm = 0, coop = 0;
        CUDA_TRY(cudaGetDevice(&dev));
        CUDA_TRY(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
        if (smem > 48 * 1024) CUDA_TRY(opt_in_smem(trsv_backward_owner_kernel, (int)smem));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, trsv_backward_owner_kernel, NTHREADS, smem));
        if (coop && G >= 2 && (long long)per_sm * sms >= (long long)G * B) {
            int* abort_flag = reinterpret_cast<int*>(work);
            int* done = abort_flag + 64;
            CUDA_TRY(cudaMemsetAsync(work, 0, 256 + (size_t)B * 4, st));
            void* args[] = {(void*)&L, (void*)&invdiag, (void*)&npad, (void*)&z, (void*)&alpha, (void*)&done, (void*)&abort_flag};
            CUDA_TRY(cudaLaunchCooperativeKernel((const void*)trsv_backward_owner_kernel, dim3(G, B), dim3(NTHREADS), args, smem, st));
            return GPR_OK;
        }
    }
    const int smem = npad * 8;
    if (smem > 200 * 1024) return fail(GPR_ERR_UNSUPPORTED, "matrix order too large for the single-CTA backward solve%s (npad=%lld)", "", npad);
    CUDA_TRY(opt_in_smem(trsv_backward_kernel, smem));
    trsv_backward_kernel<<<B, NTHREADS, smem, st>>>(L, invdiag, npad, z, alpha);
    CUDA_TRY(cudaGetLastError());
    return GPR_OK;
}

// ---------------------------------------------------------------------------------------------
// K5: batched prediction.  Persistent CTAs pull (expert, 128-test-point tile) tasks; each task is an
// independent blocked forward substitution  V^T[:, i] = (K*[:, i] - V^T[:, 0:i] L[i, 0:i]^T) inv(L_ii)^T
// — i.e. the Cholesky factorisation extended by 128 extra rows — with K* tiles generated in smem
// (never in HBM), the mean  K* alpha  and the variance reduction  sum V^2  fused into the epilogues.
// V^T lives in a per-CTA scratch slab [128, npad] that only this CTA touches.
// ---------------------------------------------------------------------------------------------
// SHARED (GRBCM with one hyper-parameter vector): the first gt block rows of every expert are the communication set, so
// the first gt substitution steps of a test tile are the same for all experts — they ARE the communication expert's
// V_g^T (kept by its own prediction, [m tiles][128][npad_g]).  A task then starts at step gt, multiplies against V_g^T for the
// k-blocks < gt, and reports the expert as an update of the communication expert:
//     mean = mean_g + V_l^T z_l      (K* alpha = V^T z, and V_g^T z_g is the communication expert's mean)
//     var  = var_g  - |V_l|^2        (var_g = prior - |V_g|^2)
// which removes 1/4 of the flops at ng = n/2 (and all K* tiles of the shared block).  `alpha` carries z in this mode.
struct PredShared {
    int gt;
    const double* Vg;
    int npad_g;
    const double* mean_g;
    const double* var_g;
};

template <bool SHARED>
__global__ void __launch_bounds__(NTHREADS, 1)
predict_kernel(const double* __restrict__ L, const double* __restrict__ invdiag, const double* __restrict__ a,
               const double* __restrict__ nrm, const double* __restrict__ alpha, const double* __restrict__ hp, int hp_stride,
               const double* __restrict__ prior, const double* __restrict__ xs, int B, int n, int npad, int d, int dpad, int m,
               double* __restrict__ mean, double* __restrict__ var, int want_var, double* __restrict__ scratch,
               int* __restrict__ counter, double* __restrict__ vt_out, int feat_own, const PredShared sh) {
    extern __shared__ double smem[];
    __shared__ unsigned long long pipe_bars[2 * STAGES];
    Pipe pipe;
    pipe_init(pipe, pipe_bars);
    __shared__ int s_task;
    __shared__ double red[2][2][TILE];  // [mean|var][wn][row]
    const int tid = threadIdx.x;
    const int nt = npad / TILE;
    const int mt = (m + TILE - 1) / TILE;
    const int ntasks = B * mt;
    const int sld = feat_sld(dpad);
    // persistent region (after the pipeline stages): scaled test tile + its norms + alpha block + col norms
    double* sT = smem + MAIN_SMEM_BYTES / 8;  // [128 x sld]
    double* nT = sT + TILE * sld;             // [128]
    double* sAl = nT + TILE;                  // [128] alpha block of the current step
    // feat_own: the training-block features of step i get their own region and are fetched BEFORE the step's long product
    // (they land while it runs); otherwise they are staged in the drained pipeline buffers after it (large d)
    double* sCown = sAl + TILE;               // [128 x sld] + [128] norms + [128] second alpha buffer (feat_own only)
    double* Vt = scratch + (long long)blockIdx.x * TILE * npad;   // replaced per task when V^T is kept (vt_out)
    NoHook nh;
    Seg none{nullptr, 0, nullptr, 0, 0};

    for (;;) {
        __syncthreads();
        if (tid == 0) s_task = atomicAdd(counter, 1);
        __syncthreads();
        const int task = s_task;
        if (task >= ntasks) break;
        const int b = task / mt, ct = task % mt;
        const double* hb = hp + (long long)b * hp_stride;
        const double sig2 = hb[0] * hb[0];
        const double* Lb = L + (long long)b * npad * npad;
        if (vt_out) Vt = vt_out + ((long long)b * mt + ct) * TILE * npad;

        // scaled test tile  as = xs * ls_b  (zero rows beyond m, zero cols beyond d) and its norms
        for (int id = tid; id < TILE * dpad; id += NTHREADS) {
            const int r = id / dpad, k = id % dpad;
            const int gr = ct * TILE + r;
            sT[r * sld + k] = (gr < m && k < d) ? xs[(long long)gr * d + k] * hb[1 + k] : 0.0;
        }
        __syncthreads();
        if (tid < TILE) {
            double s = 0.0;
            for (int k = 0; k < d; ++k) s += sT[tid * sld + k] * sT[tid * sld + k];
            nT[tid] = s;
        }
        double macc[4] = {0, 0, 0, 0}, vacc[4] = {0, 0, 0, 0};
        if (SHARED) {
            // the communication expert's V_g^T tile becomes columns 0 .. gt*128 of this CTA's slab (a 128 x ng copy, ~0.1 % of
            // the task) so that every step is ONE contiguous product like the unshared kernel; reading it through a second
            // segment cost ~3 % (two sets of addresses in the copy-issue path of the mainloop)
            const double* src = sh.Vg + (long long)ct * TILE * sh.npad_g;
            const int w2 = sh.gt * (TILE / 2);  // 16-byte words per row
            for (int id = tid; id < TILE * w2; id += NTHREADS) {
                const int r = id / w2, c2 = id % w2;
                *reinterpret_cast<double2*>(Vt + (long long)r * npad + 2 * c2) =
                    __ldcg(reinterpret_cast<const double2*>(src + (long long)r * sh.npad_g + 2 * c2));
            }
            __threadfence_block();
        }

        const int i0 = SHARED ? sh.gt : 0;
        // feat_own: features / norms / alpha block of step i are fetched during step i-1 into their own region
        // (alpha double buffered: the shared form still reads block i after the next fetch has been issued)
        double* sAl2 = sCown + TILE * sld + TILE;
        auto fetch_feat = [&](int i) {
            load_feat_tile(sCown, sld, a + ((long long)b * npad + i * TILE) * dpad, dpad);
            cp_async_commit();
            if (tid < TILE) {
                sCown[TILE * sld + tid] = nrm[(long long)b * npad + i * TILE + tid];
                ((i & 1) ? sAl2 : sAl)[tid] = alpha[(long long)b * npad + i * TILE + tid];
            }
        };
        if (feat_own) {
            __syncthreads();   // the previous task's last step is done with the region
            fetch_feat(i0);
        }

        for (int i = i0; i < nt; ++i) {
            double acc[4][8][2];
            zero_acc(acc);
            double* Vi = Vt + i * TILE;
            const double* sAli = (feat_own && (i & 1)) ? sAl2 : sAl;
            if (feat_own) {
                // K* tile FIRST, straight into the accumulators as -K*, then  acc += V^T[:, 0:i] L[i, 0:i]^T  on top: X = -acc
                cp_async_wait<0>();
                __syncthreads();  // features of step i landed; V^T columns written by the previous step are visible
                se_seed_acc(acc, sT, sCown, sld, nT, sCown + TILE * sld, dpad, sig2, [&](int mi, int ni, int e, double kv) {
                    const int c = acc_col(ni, e);
                    if (i * TILE + c >= n) kv = 0.0;  // padding columns carry no covariance
                    if (!SHARED) macc[mi] += kv * sAli[c];
                    return -kv;
                });
                __syncthreads();  // every warp is done with the feature region
                if (i + 1 < nt) fetch_feat(i + 1);
                if (!want_var) continue;
                if (SHARED || i > 0) {
                    Seg s0{Vt, npad, Lb + ((long long)i * TILE) * npad, npad, i * (TILE / KC)};
                    mainloop<false, XPF_PREDICT>(acc, s0, none, smem, nh, pipe);
                }
#pragma unroll
                for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                    for (int ni = 0; ni < 8; ++ni)
                        *reinterpret_cast<double2*>(Vi + (long long)acc_row(mi) * npad + acc_col(ni, 0)) =
                            make_double2(-acc[mi][ni][0], -acc[mi][ni][1]);
            } else {
                __syncthreads();  // V^T columns written by the previous step are visible; staged features are free
                if (SHARED || (want_var && i > 0)) {
                    Seg s0{Vt, npad, Lb + ((long long)i * TILE) * npad, npad, i * (TILE / KC)};
                    mainloop<false, XPF_PREDICT>(acc, s0, none, smem, nh, pipe);
                }
                // K*(test tile, train block i) from features staged in the drained pipeline buffers
                double* sC = smem;
                double* nC = sC + TILE * sld;
                load_feat_tile(sC, sld, a + ((long long)b * npad + i * TILE) * dpad, dpad);
                cp_async_commit();
                if (tid < TILE) {
                    nC[tid] = nrm[(long long)b * npad + i * TILE + tid];
                    sAl[tid] = alpha[(long long)b * npad + i * TILE + tid];
                }
                cp_async_wait<0>();
                __syncthreads();
                gen_se_tile(sT, sC, sld, nT, nC, dpad, sig2, [&](int mi, int ni, int e, double kv) {
                    const int c = acc_col(ni, e);
                    if (i * TILE + c >= n) kv = 0.0;  // padding columns carry no covariance
                    if (!SHARED) macc[mi] += kv * sAl[c];
                    acc[mi][ni][e] = kv - acc[mi][ni][e];
                });
                if (!want_var) continue;
#pragma unroll
                for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                    for (int ni = 0; ni < 8; ++ni)
                        *reinterpret_cast<double2*>(Vi + (long long)acc_row(mi) * npad + acc_col(ni, 0)) =
                            make_double2(acc[mi][ni][0], acc[mi][ni][1]);
            }
            __threadfence_block();
            __syncthreads();
            zero_acc(acc);
            Seg s1{Vi, npad, invdiag + ((long long)b * nt + i) * TILE * TILE, TILE, TILE / KC};
            mainloop<false, XPF_PREDICT>(acc, s1, none, smem, nh, pipe);
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 8; ++ni) {
                    vacc[mi] += acc[mi][ni][0] * acc[mi][ni][0] + acc[mi][ni][1] * acc[mi][ni][1];
                    if (SHARED) macc[mi] += acc[mi][ni][0] * sAli[acc_col(ni, 0)] + acc[mi][ni][1] * sAli[acc_col(ni, 1)];  // V^T z
                    *reinterpret_cast<double2*>(Vi + (long long)acc_row(mi) * npad + acc_col(ni, 0)) =
                        make_double2(acc[mi][ni][0], acc[mi][ni][1]);
                }
            __threadfence_block();
        }
        // reduce the per-thread row partials: over the 4 lanes of a quad, then over the two column warps
        const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) {
            double mv = macc[mi], vv = vacc[mi];
            mv += __shfl_xor_sync(0xffffffffu, mv, 1);
            mv += __shfl_xor_sync(0xffffffffu, mv, 2);
            vv += __shfl_xor_sync(0xffffffffu, vv, 1);
            vv += __shfl_xor_sync(0xffffffffu, vv, 2);
            if ((lane & 3) == 0) {
                red[0][warp >> 2][acc_row(mi)] = mv;
                red[1][warp >> 2][acc_row(mi)] = vv;
            }
        }
        __syncthreads();
        if (tid < TILE) {
            const int gr = ct * TILE + tid;
            if (gr < m) {
                if (SHARED) {
                    mean[(long long)b * m + gr] = sh.mean_g[gr] + (red[0][0][tid] + red[0][1][tid]);
                    var[(long long)b * m + gr] = sh.var_g[gr] - (red[1][0][tid] + red[1][1][tid]);
                } else {
                    mean[(long long)b * m + gr] = red[0][0][tid] + red[0][1][tid];
                    if (want_var) var[(long long)b * m + gr] = prior[b] - (red[1][0][tid] + red[1][1][tid]);
                }
            }
        }
    }
}

static inline long long predict_smem_bytes(int dpad) { return MAIN_SMEM_BYTES + (long long)(TILE * feat_sld(dpad) + 2 * TILE) * 8; }
// with a private region for the training-block features (+ norms) when it still fits in 227 KB
static inline long long predict_smem_bytes_own(int dpad) { return predict_smem_bytes(dpad) + (long long)(TILE * feat_sld(dpad) + 2 * TILE) * 8; }

extern "C" long long gpr_predict_scratch_bytes(int npad) {
    int sms = gpr_device_sm_count();
    if (sms <= 0) sms = 148;
    return (long long)sms * TILE * npad * 8 + 256;
}

static int predict_impl(const double* L, const double* invdiag, const double* a, const double* nrm, const double* alpha,
                        const double* hp, int hp_stride, const double* prior, const double* xs, int B, int n, int npad, int d,
                        int dpad, int m, double* mean, double* var, int want_var, void* scratch, void* stream, double* vt_out,
                        const PredShared* shp) {
    ARG_CHECK(L && invdiag && a && nrm && alpha && hp && xs && mean && scratch);
    ARG_CHECK(!want_var || (var && (prior || shp)));
    ARG_CHECK(B > 0 && n > 0 && m > 0 && d > 0 && npad % TILE == 0 && npad >= n && dpad % 4 == 0 && dpad >= d);
    long long smem = predict_smem_bytes(dpad);
    if (smem > 227 * 1024 || (TILE * feat_sld(dpad) + TILE) * 8 > MAIN_SMEM_BYTES)
        return fail(GPR_ERR_UNSUPPORTED, "input dimension too large for the fused prediction%s (dpad=%lld)", "", dpad);
    const int feat_own = predict_smem_bytes_own(dpad) <= 227 * 1024 ? 1 : 0;
    if (feat_own) smem = predict_smem_bytes_own(dpad);
    cudaStream_t st = (cudaStream_t)stream;
    int sms = gpr_device_sm_count();
    if (sms <= 0) return fail(GPR_ERR_CUDA, "no CUDA device%s", "");
    const int mt = (m + TILE - 1) / TILE;
    const long long ntasks = (long long)B * mt;
    const int grid = (int)(ntasks < sms ? ntasks : sms);
    // layout of scratch: [int counter (256 B)] [grid slabs of 128 x npad doubles]
    int* counter = reinterpret_cast<int*>(scratch);
    double* slabs = reinterpret_cast<double*>(reinterpret_cast<char*>(scratch) + 256);
    CUDA_TRY(cudaMemsetAsync(counter, 0, 256, st));
    if (shp) {
        CUDA_TRY(opt_in_smem(predict_kernel<true>, (int)smem));
        predict_kernel<true><<<grid, NTHREADS, (int)smem, st>>>(L, invdiag, a, nrm, alpha, hp, hp_stride, prior, xs, B, n, npad, d, dpad,
                                                                m, mean, var, 1, slabs, counter, nullptr, feat_own, *shp);
    } else {
        CUDA_TRY(opt_in_smem(predict_kernel<false>, (int)smem));
        predict_kernel<false><<<grid, NTHREADS, (int)smem, st>>>(L, invdiag, a, nrm, alpha, hp, hp_stride, prior, xs, B, n, npad, d, dpad,
                                                                 m, mean, var, want_var, vt_out ? vt_out : slabs, counter, vt_out, feat_own,
                                                                 PredShared{0, nullptr, 0, nullptr, nullptr});
    }
    CUDA_TRY(cudaGetLastError());
    return GPR_OK;
}

extern "C" int gpr_predict(const double* L, const double* invdiag, const double* a, const double* nrm, const double* alpha,
                           const double* hp, int hp_stride, const double* prior, const double* xs, int B, int n, int npad,
                           int d, int dpad, int m, double* mean, double* var, int want_var, void* scratch, void* stream) {
    return predict_impl(L, invdiag, a, nrm, alpha, hp, hp_stride, prior, xs, B, n, npad, d, dpad, m, mean, var, want_var, scratch,
                        stream, nullptr, nullptr);
}

// gpr_predict that also keeps V^T = K* L^-T  [B, ceil(m/128), 128, npad]  (the shared prefix of gpr_predict_shared)
extern "C" int gpr_predict_keep(const double* L, const double* invdiag, const double* a, const double* nrm, const double* alpha,
                                const double* hp, int hp_stride, const double* prior, const double* xs, int B, int n, int npad,
                                int d, int dpad, int m, double* mean, double* var, double* vt, void* scratch, void* stream) {
    ARG_CHECK(vt);
    return predict_impl(L, invdiag, a, nrm, alpha, hp, hp_stride, prior, xs, B, n, npad, d, dpad, m, mean, var, 1, scratch, stream, vt,
                        nullptr);
}

extern "C" int gpr_predict_shared(const double* L, const double* invdiag, const double* a, const double* nrm, const double* z,
                                  const double* hp, int hp_stride, const double* xs, int B, int n, int npad, int d, int dpad, int m,
                                  double* mean, double* var, void* scratch, void* stream, int ng, const double* vt_g, int npad_g,
                                  const double* mean_g, const double* var_g) {
    ARG_CHECK(vt_g && mean_g && var_g && z && var && ng > 0 && ng % TILE == 0 && ng < n && npad_g >= ng && npad_g % TILE == 0);
    PredShared sh{ng / TILE, vt_g, npad_g, mean_g, var_g};
    return predict_impl(L, invdiag, a, nrm, z, hp, hp_stride, nullptr, xs, B, n, npad, d, dpad, m, mean, var, 1, scratch, stream, nullptr,
                        &sh);
}

// ---------------------------------------------------------------------------------------------
// N1: full predictive covariance  K** - V^T V  (gpr.py:108-120).  The forward substitution of predict_kernel is run
// with V^T kept in a caller buffer [B, mpad, npad]; a tile kernel then forms  cov = K**(xs,xs) + noise I - V^T V
// with K** generated on chip from the scaled test points.  API-parity path (O(B m n) memory), not the bench path.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHREADS, 1)
covar_full_kernel(const double* __restrict__ vt, const double* __restrict__ as, const double* __restrict__ nrms,
                  const double* __restrict__ hp, int hp_stride, const double* __restrict__ noise_var, int m, int mpad, int npad,
                  int dpad, double* __restrict__ cov) {
    extern __shared__ double smem[];
    __shared__ unsigned long long pipe_bars[2 * STAGES];
    Pipe pipe;
    pipe_init(pipe, pipe_bars);
    const int tj = blockIdx.x, ti = blockIdx.y, b = blockIdx.z;
    const double* Vb = vt + (long long)b * mpad * npad;
    double acc[4][8][2];
    zero_acc(acc);
    NoHook nh;
    Seg none{nullptr, 0, nullptr, 0, 0};
    Seg s0{Vb + (long long)ti * TILE * npad, npad, Vb + (long long)tj * TILE * npad, npad, npad / KC};
    mainloop<false, XPF_PREDICT>(acc, s0, none, smem, nh, pipe);
    FeatSmem f = carve_feat(smem, dpad);
    load_feat_tile(f.sR, f.sld, as + ((long long)b * mpad + ti * TILE) * dpad, dpad);
    load_feat_tile(f.sC, f.sld, as + ((long long)b * mpad + tj * TILE) * dpad, dpad);
    cp_async_commit();
    if (threadIdx.x < TILE) {
        f.nR[threadIdx.x] = nrms[(long long)b * mpad + ti * TILE + threadIdx.x];
        f.nC[threadIdx.x] = nrms[(long long)b * mpad + tj * TILE + threadIdx.x];
    }
    cp_async_wait<0>();
    __syncthreads();
    const double sig = hp[(long long)b * hp_stride];
    const double nv = noise_var ? noise_var[b] : 0.0;
    double* cb = cov + (long long)b * m * m;
    gen_se_tile(f.sR, f.sC, f.sld, f.nR, f.nC, dpad, sig * sig, [&](int mi, int ni, int e, double kv) {
        const long long r = (long long)ti * TILE + acc_row(mi), c = (long long)tj * TILE + acc_col(ni, e);
        if (r < m && c < m) cb[r * m + c] = ((r == c) ? kv + nv : kv) - acc[mi][ni][e];
    });
}

extern "C" long long gpr_predict_full_scratch_bytes(int B, int npad, int m) {
    const long long mpad = (m + TILE - 1) / TILE * TILE;
    return 256 + (long long)B * mpad * npad * 8;
}

extern "C" int gpr_predict_full(const double* L, const double* invdiag, const double* a, const double* nrm, const double* alpha,
                                const double* hp, int hp_stride, const double* prior, const double* noise_var, const double* xs,
                                const double* as, const double* nrms, int B, int n, int npad, int d, int dpad, int m, int mpad,
                                double* mean, double* var, double* cov, void* scratch, void* stream) {
    ARG_CHECK(L && invdiag && a && nrm && alpha && hp && prior && xs && as && nrms && mean && var && cov && scratch);
    ARG_CHECK(B > 0 && n > 0 && m > 0 && d > 0 && npad % TILE == 0 && npad >= n && dpad % 4 == 0 && dpad >= d);
    ARG_CHECK(mpad % TILE == 0 && mpad >= m && mpad - m < TILE);
    long long smem = predict_smem_bytes(dpad);
    if (smem > 227 * 1024 || feat_smem_bytes(dpad) > MAIN_SMEM_BYTES)
        return fail(GPR_ERR_UNSUPPORTED, "input dimension too large for the fused prediction%s (dpad=%lld)", "", dpad);
    const int feat_own = predict_smem_bytes_own(dpad) <= 227 * 1024 ? 1 : 0;
    if (feat_own) smem = predict_smem_bytes_own(dpad);
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(opt_in_smem(predict_kernel<false>, (int)smem));
    CUDA_TRY(opt_in_smem(covar_full_kernel, MAIN_SMEM_BYTES));
    int sms = gpr_device_sm_count();
    if (sms <= 0) return fail(GPR_ERR_CUDA, "no CUDA device%s", "");
    const int mt = mpad / TILE;
    const long long ntasks = (long long)B * mt;
    const int grid = (int)(ntasks < sms ? ntasks : sms);
    int* counter = reinterpret_cast<int*>(scratch);
    double* vt = reinterpret_cast<double*>(reinterpret_cast<char*>(scratch) + 256);
    CUDA_TRY(cudaMemsetAsync(counter, 0, 256, st));
    predict_kernel<false><<<grid, NTHREADS, (int)smem, st>>>(L, invdiag, a, nrm, alpha, hp, hp_stride, prior, xs, B, n, npad, d, dpad, m,
                                                             mean, var, 1, vt, counter, vt, feat_own,
                                                             PredShared{0, nullptr, 0, nullptr, nullptr});
    covar_full_kernel<<<dim3(mt, mt, B), NTHREADS, MAIN_SMEM_BYTES, st>>>(vt, as, nrms, hp, hp_stride, noise_var, m, mpad, npad, dpad, cov);
    CUDA_TRY(cudaGetLastError());
    return GPR_OK;
}

// ---------------------------------------------------------------------------------------------
// K8: GRBCM aggregation (HBM-trivial elementwise + sum over the expert axis)        gr_bcm.py:116-149
// ---------------------------------------------------------------------------------------------
// 32 test points x 8 expert groups per block: a warp reads 32 consecutive values of one expert row (256 B, coalesced), the 8
// warps take the experts i = w, w+8, ... and their partial sums are added in the fixed order w = 0..7 (deterministic; the
// one-thread-per-test-point version walked 255 dependent log() chains per thread on 32 blocks: 0.9 % of the HBM rate)
constexpr int AGG_GROUPS = 8;
__global__ void __launch_bounds__(32 * AGG_GROUPS)
grbcm_partial_kernel(const double* __restrict__ mean_l, const double* __restrict__ var_l, const double* __restrict__ var_g,
                     int nc_local, int first_global_index, int m, double* __restrict__ S, double* __restrict__ beta_rows,
                     double* __restrict__ prec_rows) {
    __shared__ double part[3][AGG_GROUPS][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + lane;
    double sb = 0.0, sp = 0.0, sm = 0.0;
    if (c < m) {
        const double pg = 1.0 / var_g[c];              // gr_bcm.py:128
        const double lpg = log(pg);
        for (int i = w; i < nc_local; i += AGG_GROUPS) {
            const double p = 1.0 / var_l[(long long)i * m + c];             // gr_bcm.py:129
            double beta = 0.5 * (log(p) - lpg);                              // gr_bcm.py:131
            if (first_global_index + i == 0) beta = 1.0;                     // gr_bcm.py:132
            const double pb = p * beta;                                      // gr_bcm.py:140
            sb += beta;
            sp += pb;
            sm += mean_l[(long long)i * m + c] * pb;
            if (beta_rows) beta_rows[(long long)i * m + c] = beta;
            if (prec_rows) prec_rows[(long long)i * m + c] = p;
        }
    }
    part[0][w][lane] = sb;
    part[1][w][lane] = sp;
    part[2][w][lane] = sm;
    __syncthreads();
    if (w < 3 && c < m) {
        double t = 0.0;
#pragma unroll
        for (int g = 0; g < AGG_GROUPS; ++g) t += part[w][g][lane];
        S[(long long)w * m + c] = t;
    }
}

__global__ void grbcm_finalize_kernel(const double* __restrict__ S, const double* __restrict__ mean_g,
                                      const double* __restrict__ var_g, int m, double* __restrict__ mu,
                                      double* __restrict__ var, double* __restrict__ beta0, double* __restrict__ prec0) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= m) return;
    const double pg = 1.0 / var_g[c];
    const double b0 = -(S[c] - 1.0);                                     // gr_bcm.py:133
    const double pb0 = pg * b0;
    const double v = 1.0 / (pb0 + S[m + c]);                             // gr_bcm.py:143
    mu[c] = (mean_g[c] * pb0 + S[2 * (long long)m + c]) * v;             // gr_bcm.py:144
    var[c] = v;
    if (beta0) beta0[c] = b0;
    if (prec0) prec0[c] = pg;
}

extern "C" int grbcm_partial(const double* mean_l, const double* var_l, const double* var_g, int nc_local,
                             int first_global_index, int m, double* S, double* beta_rows, double* prec_rows, void* stream) {
    ARG_CHECK(var_g && S && m > 0 && nc_local >= 0 && (nc_local == 0 || (mean_l && var_l)));
    grbcm_partial_kernel<<<(m + 31) / 32, 32 * AGG_GROUPS, 0, (cudaStream_t)stream>>>(mean_l, var_l, var_g, nc_local,
                                                                                      first_global_index, m, S, beta_rows, prec_rows);
    CUDA_TRY(cudaGetLastError());
    return GPR_OK;
}

extern "C" int grbcm_finalize(const double* S, const double* mean_g, const double* var_g, int m, double* mu, double* var,
                              double* beta0, double* prec0, void* stream) {
    ARG_CHECK(S && mean_g && var_g && mu && var && m > 0);
    grbcm_finalize_kernel<<<(m + 127) / 128, 128, 0, (cudaStream_t)stream>>>(S, mean_g, var_g, m, mu, var, beta0, prec0);
    CUDA_TRY(cudaGetLastError());
    return GPR_OK;
}

// ---------------------------------------------------------------------------------------------
// K6/K7: gradient of the negative log marginal likelihood without forming dK      loss.py:59-128
//   (1) transpose_invdiag_kernel : U_jj = inv(L_jj)^T
//   (2) trtri_kernel   : U = L^-T, one persistent task per (expert, block row j); forward substitution on the
//                        identity, i.e.  U[j,i] = -(sum_{k=j}^{i-1} U[j,k] L[i,k]^T) inv(L_ii)^T,  written into the
//                        strict block-upper part of L's buffer (row tile j is private to its task)
//   (3) lauum_grad_kernel : K^-1[i,j] = sum_{k>=i} U[i,k] U[j,k]^T (lower tiles only), never stored: the epilogue
//                        forms M = (K^-1 - alpha alpha^T) o K_se with K_se generated on chip and reduces
//                        sum M, sum M (a_ik - a_jk)^2 per input dimension, and tr(W) into per-tile partials
//   (4) grad_finalize_kernel : fixed-order (deterministic) sum of the partials and the 1/sig, -1/ls_k, sig_n factors
// ---------------------------------------------------------------------------------------------
__global__ void transpose_invdiag_kernel(const double* __restrict__ invdiag, double* __restrict__ invdiagT) {
    __shared__ double t[32][33];
    const long long base = (long long)blockIdx.z * TILE * TILE;
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) t[r][threadIdx.x] = invdiag[base + (long long)(by + r) * TILE + bx + threadIdx.x];
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) invdiagT[base + (long long)(bx + r) * TILE + by + threadIdx.x] = t[threadIdx.x][r];
}

__global__ void __launch_bounds__(NTHREADS, 1)
trtri_kernel(double* __restrict__ L, const double* __restrict__ invdiag, const double* __restrict__ invdiagT, int B, int npad,
             int* __restrict__ counter, int gt, int seg) {
    extern __shared__ double smem[];
    __shared__ unsigned long long pipe_bars[2 * STAGES];
    Pipe pipe;
    pipe_init(pipe, pipe_bars);
    __shared__ int s_task;
    const int nt = npad / TILE;
    const int ntasks = B * (nt - 1);
    NoHook nh;
    Seg none{nullptr, 0, nullptr, 0, 0};
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_task = atomicAdd(counter, 1);
        __syncthreads();
        const int task = s_task;
        if (task >= ntasks) break;
        const int b = task / (nt - 1), j = task % (nt - 1);  // j ascending = longest rows first within an expert
        double* Lb = L + (long long)b * npad * npad;
        const double* Ujj = invdiagT + ((long long)b * nt + j) * TILE * TILE;
        double* Urow = Lb + ((long long)j * TILE) * npad;      // row tile j of U lives in the block-upper part
        // gt > 0 (shared leading block): U[j, i] for i < gt is the communication expert's L_g^-T, already in place.
        // seg > 0: only the diagonal segments of `seg` block rows are inverted here (U[j, i] for i inside j's segment); the blocks
        // between segments are products of the segment inverses and come from the emulated-FP64 GEMM (engine.inverse_factors).
        const int iend = seg > 0 ? min(nt, (j / seg + 1) * seg) : nt;
        for (int i = (j + 1 > gt ? j + 1 : gt); i < iend; ++i) {
            double acc[4][8][2];
            zero_acc(acc);
            __syncthreads();
            const double* Li = Lb + ((long long)i * TILE) * npad;
            Seg s0{Ujj, TILE, Li + (long long)j * TILE, npad, TILE / KC};
            Seg s1{Urow + (long long)(j + 1) * TILE, npad, Li + (long long)(j + 1) * TILE, npad, (i - j - 1) * (TILE / KC)};
            mainloop<false, XPF_TRTRI>(acc, s0, s1, smem, nh, pipe);
            double* Uji = Urow + (long long)i * TILE;
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 8; ++ni)
                    *reinterpret_cast<double2*>(Uji + (long long)acc_row(mi) * npad + acc_col(ni, 0)) =
                        make_double2(-acc[mi][ni][0], -acc[mi][ni][1]);
            __threadfence_block();
            __syncthreads();
            zero_acc(acc);
            Seg s2{Uji, npad, invdiag + ((long long)b * nt + i) * TILE * TILE, TILE, TILE / KC};
            mainloop<false, XPF_TRTRI>(acc, s2, none, smem, nh, pipe);
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 8; ++ni)
                    *reinterpret_cast<double2*>(Uji + (long long)acc_row(mi) * npad + acc_col(ni, 0)) =
                        make_double2(acc[mi][ni][0], acc[mi][ni][1]);
            __threadfence_block();
        }
    }
}

// Column mode of U = L^-T for FEW experts: one launch per block column i, tiles j < i, k-range of i-j blocks split so that
// every CTA gets about `target` blocks (split-K with last-arriver fix-up).  Task mode above has a serial chain of
// O(nt^2) tiles per row; with B*(nt-1) < 2*SMs tasks that chain, not the SM count, sets the time (B=1, n=16384: 140 ms).
__global__ void __launch_bounds__(NTHREADS, 1)
trtri_col_kernel(double* __restrict__ L, const double* __restrict__ invdiag, const double* __restrict__ invdiagT, int npad, int i,
                 int target, double* __restrict__ sk_slots, int* __restrict__ sk_counters, int slots_per_expert) {
    extern __shared__ double smem[];
    __shared__ unsigned long long pipe_bars[2 * STAGES];
    __shared__ int s_flag;
    Pipe pipe;
    pipe_init(pipe, pipe_bars);
    const int nt = npad / TILE;
    const int b = blockIdx.y;
    // blockIdx.x -> (row tile j, part, S_j); slot offset = number of parts of the tiles before j
    int bx = blockIdx.x, j = 0, S = 1, slot_off = 0;
    for (; j < i; ++j) {
        S = (i - j + target - 1) / target;
        if (bx < S) break;
        bx -= S;
        slot_off += S;
    }
    const int part = bx, depth = i - j;
    const int q0 = (int)((long long)part * depth / S), q1 = (int)((long long)(part + 1) * depth / S);
    double* Lb = L + (long long)b * npad * npad;
    const double* Ujj = invdiagT + ((long long)b * nt + j) * TILE * TILE;
    double* Urow = Lb + ((long long)j * TILE) * npad;
    const double* Li = Lb + ((long long)i * TILE) * npad;
    double acc[4][8][2];
    zero_acc(acc);
    NoHook nh;
    Seg none{nullptr, 0, nullptr, 0, 0};
    // k-block 0 of the range is U_jj (from invdiagT), k-blocks 1.. are U[j, j+1..] in the block-upper part of L
    Seg s0{Ujj, TILE, Li + (long long)j * TILE, npad, (q0 == 0) ? TILE / KC : 0};
    const int r0 = (q0 == 0) ? 1 : q0;
    Seg s1{Urow + (long long)(j + r0) * TILE, npad, Li + (long long)(j + r0) * TILE, npad, (q1 - r0) * (TILE / KC)};
    mainloop<false, XPF_TRTRI>(acc, s0, s1, smem, nh, pipe);
    if (S > 1 && !splitk_combine(acc, sk_slots + ((long long)b * slots_per_expert + slot_off) * TILE * TILE, part, S,
                                 sk_counters + (long long)b * nt + j, &s_flag))
        return;
    double* Uji = Urow + (long long)i * TILE;
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 8; ++ni)
            *reinterpret_cast<double2*>(Uji + (long long)acc_row(mi) * npad + acc_col(ni, 0)) =
                make_double2(-acc[mi][ni][0], -acc[mi][ni][1]);
    __threadfence_block();
    __syncthreads();
    zero_acc(acc);
    Seg s2{Uji, npad, invdiag + ((long long)b * nt + i) * TILE * TILE, TILE, TILE / KC};
    mainloop<false, XPF_TRTRI>(acc, s2, none, smem, nh, pipe);
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 8; ++ni)
            *reinterpret_cast<double2*>(Uji + (long long)acc_row(mi) * npad + acc_col(ni, 0)) =
                make_double2(acc[mi][ni][0], acc[mi][ni][1]);
}

// U = L^-T for all experts: persistent task mode when there are enough (expert, row) tasks to fill the GPU, column mode
// with split-K otherwise.  `work`: [256 B task counter][B*nt ints][slot tiles] (see trtri_work_bytes).
static long long trtri_slot_tiles(int nt, int sms) { return 2LL * sms + nt; }
static long long trtri_work_bytes(int B, int npad, int sms) {
    const int nt = npad / TILE;
    // the split-K slots are only used by the column mode (few experts); the task mode needs the counter alone
    // (255 experts of order 8192 would otherwise reserve 12 GB of slots beside 137 GB of factors)
    const bool task_mode = (long long)B * (nt - 1) >= 2LL * sms;
    return 256 + align256((long long)B * nt * 4) + (task_mode ? 0 : (long long)B * trtri_slot_tiles(nt, sms) * TILE * TILE * 8);
}
static int launch_trtri(double* L, const double* invdiag, const double* invdiagT, int B, int npad, void* work, cudaStream_t st,
                        int gt = 0, int seg = 0) {
    const int nt = npad / TILE;
    if (nt <= 1) return GPR_OK;
    int sms = gpr_device_sm_count();
    if (sms <= 0) return fail(GPR_ERR_CUDA, "no CUDA device%s", "");
    int* task_counter = reinterpret_cast<int*>(work);
    const long long ntasks = (long long)B * (nt - 1);
    if (ntasks >= 2LL * sms) {
        CUDA_TRY(cudaMemsetAsync(task_counter, 0, 256, st));
        CUDA_TRY(opt_in_smem(trtri_kernel, MAIN_SMEM_BYTES));
        trtri_kernel<<<sms, NTHREADS, MAIN_SMEM_BYTES, st>>>(L, invdiag, invdiagT, B, npad, task_counter, gt, seg);
        CUDA_TRY(cudaGetLastError());
        return GPR_OK;
    }
    if (seg > 0) return fail(GPR_ERR_UNSUPPORTED, "segmented trtri needs the task schedule%s (B*(nt-1) = %lld tasks)", "", ntasks);
    int* counters = reinterpret_cast<int*>(reinterpret_cast<char*>(work) + 256);
    double* slots = reinterpret_cast<double*>(reinterpret_cast<char*>(counters) + align256((long long)B * nt * 4));
    const int spe = (int)trtri_slot_tiles(nt, sms);
    CUDA_TRY(cudaMemsetAsync(counters, 0, (size_t)B * nt * 4, st));
    CUDA_TRY(opt_in_smem(trtri_col_kernel, MAIN_SMEM_BYTES));
    for (int i = (gt > 1 ? gt : 1); i < nt; ++i) {                       // block columns < gt of U are shared and in place
        const long long blocks = (long long)i * (i + 1) / 2;               // k-blocks of this column, per expert
        long long target = (blocks * B + 2LL * sms - 1) / (2LL * sms);      // ~2 CTAs per SM over all experts
        if (target < 1) target = 1;
        int nparts = 0;
        for (int j = 0; j < i; ++j) nparts += (int)((i - j + target - 1) / target);
        trtri_col_kernel<<<dim3(nparts, B), NTHREADS, MAIN_SMEM_BYTES, st>>>(L, invdiag, invdiagT, npad, i, (int)target, slots, counters, spe);
    }
    CUDA_TRY(cudaGetLastError());
    return GPR_OK;
}

__global__ void __launch_bounds__(NTHREADS, 1)
lauum_grad_kernel(const double* __restrict__ L, const double* __restrict__ invdiagT, const double* __restrict__ a,
                  const double* __restrict__ nrm, const double* __restrict__ alpha, const double* __restrict__ hp, int hp_stride,
                  int n, int npad, int d, int dpad, double* __restrict__ partials, int P, int gt, const double* __restrict__ Kinv_g,
                  int ldg, const double* __restrict__ Kinv_full = nullptr, int ldk = 0) {
    extern __shared__ double smem[];
    __shared__ unsigned long long pipe_bars[2 * STAGES];
    Pipe pipe;
    pipe_init(pipe, pipe_bars);
    __shared__ double sAi[TILE], sAj[TILE];
    __shared__ double wred[NTHREADS / 32][36];
    const int nt = npad / TILE;
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // tile index -> (i, j), i >= j, ordered by i (long k-ranges first)
    int ti = blockIdx.x, i = 0;
    while (ti >= i + 1) { ti -= i + 1; ++i; }
    const int j = ti;
    const double* Lb = L + (long long)b * npad * npad;
    const double* Uii = invdiagT + ((long long)b * nt + i) * TILE * TILE;
    const double* Ui = Lb + ((long long)i * TILE) * npad;
    const double* Uj = Lb + ((long long)j * TILE) * npad;

    double acc[4][8][2];
    zero_acc(acc);
    NoHook nh;
    // Shared leading block (gt > 0): for a tile inside it the k-blocks < gt give the communication expert's K_g^-1 tile, which is
    // the same for every expert — it is added from Kinv_g in the epilogue and only the k-blocks >= gt are multiplied here.
    const bool in_shared = i < gt;
    Seg s0{Uii, TILE, (i == j) ? Uii : Uj + (long long)i * TILE, (i == j) ? (long long)TILE : (long long)npad, in_shared ? 0 : TILE / KC};
    const int k0 = in_shared ? gt : i + 1;
    Seg s1{Ui + (long long)k0 * TILE, npad, Uj + (long long)k0 * TILE, npad, (nt - k0) * (TILE / KC)};
    if (Kinv_full) {
        // K^-1 comes materialised (gpr_grad_from_kinv: the emulated-FP64 product wrote it): the tile is read, not multiplied
        const double* kb = Kinv_full + (long long)b * ldk * ldk;
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < 8; ++ni) {
                const double2 v = *reinterpret_cast<const double2*>(kb + ((long long)i * TILE + acc_row(mi)) * ldk + j * TILE + acc_col(ni, 0));
                acc[mi][ni][0] = v.x;
                acc[mi][ni][1] = v.y;
            }
    } else {
        mainloop<false, XPF_LAUUM>(acc, s0, s1, smem, nh, pipe);
        if (in_shared) {
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 8; ++ni) {
                    const double2 v = *reinterpret_cast<const double2*>(Kinv_g + ((long long)i * TILE + acc_row(mi)) * ldg + j * TILE + acc_col(ni, 0));
                    acc[mi][ni][0] += v.x;
                    acc[mi][ni][1] += v.y;
                }
        }
        __syncthreads();                                         // the pipeline's shared memory becomes the epilogue's tile below
    }

    // epilogue: M = (K^-1 - alpha_i alpha_j^T) o K_se  -> smem tile
    double* Ms = smem;
    FeatSmem f = carve_feat(smem + TILE * XLD, dpad);
    load_feat_tile(f.sR, f.sld, a + ((long long)b * npad + i * TILE) * dpad, dpad);
    load_feat_tile(f.sC, f.sld, a + ((long long)b * npad + j * TILE) * dpad, dpad);
    cp_async_commit();
    if (tid < TILE) {
        f.nR[tid] = nrm[(long long)b * npad + i * TILE + tid];
        f.nC[tid] = nrm[(long long)b * npad + j * TILE + tid];
        sAi[tid] = alpha[(long long)b * npad + i * TILE + tid];
        sAj[tid] = alpha[(long long)b * npad + j * TILE + tid];
    }
    cp_async_wait<0>();
    __syncthreads();
    const double* hb = hp + (long long)b * hp_stride;
    double trw = 0.0;
    gen_se_tile(f.sR, f.sC, f.sld, f.nR, f.nC, dpad, hb[0] * hb[0], [&](int mi, int ni, int e, double kv) {
        const int r = acc_row(mi), c = acc_col(ni, e);
        const bool live = (i * TILE + r < n) && (j * TILE + c < n);
        const double w = acc[mi][ni][e] - sAi[r] * sAj[c];
        Ms[r * XLD + c] = live ? w * kv : 0.0;
        if (i == j && r == c && live) trw += w;
    });
    __syncthreads();
    // reduction: thread -> column c, 64 rows;  q[0] = sum M, q[1+k] = sum M (a_rk - a_ck)^2
    const int c = tid & 127, rh = tid >> 7;
    const double wgt = (i == j) ? 1.0 : 2.0;
    double q0 = 0.0;
    for (int k0 = 0; k0 < d; k0 += 16) {
        double ac[16], q[16];
        const int kn = (d - k0 < 16) ? d - k0 : 16;
#pragma unroll
        for (int k = 0; k < 16; ++k) { ac[k] = (k < kn) ? f.sC[c * f.sld + k0 + k] : 0.0; q[k] = 0.0; }
        for (int r = rh * 64; r < rh * 64 + 64; ++r) {
            const double mv = Ms[r * XLD + c];
            if (k0 == 0) q0 += mv;
            const double* ar = f.sR + r * f.sld + k0;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                if (k < kn) { const double df = ar[k] - ac[k]; q[k] += mv * df * df; }
            }
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const double v = warp_sum(q[k]);
            if (lane == 0) wred[warp][k] = v;
        }
        __syncthreads();
        if (tid < kn) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < NTHREADS / 32; ++w) s += wred[w][tid];
            partials[((long long)b * gridDim.x + blockIdx.x) * P + 1 + k0 + tid] = s * wgt;
        }
        __syncthreads();
    }
    {
        const double v0 = warp_sum(q0), v1 = warp_sum(trw);
        if (lane == 0) { wred[warp][0] = v0; wred[warp][1] = v1; }
        __syncthreads();
        if (tid == 0) {
            double s0 = 0.0, s1 = 0.0;
            for (int w = 0; w < NTHREADS / 32; ++w) { s0 += wred[w][0]; s1 += wred[w][1]; }
            double* pt = partials + ((long long)b * gridDim.x + blockIdx.x) * P;
            pt[0] = s0 * wgt;
            if (P > d + 1) pt[d + 1] = s1;
        }
    }
}

__global__ void grad_finalize_kernel(const double* __restrict__ partials, int ntiles, int P, const double* __restrict__ hp,
                                     int hp_stride, int d, double* __restrict__ grad) {
    const int b = blockIdx.x, p = threadIdx.x;
    if (p >= P) return;
    double s = 0.0;
    for (int t = 0; t < ntiles; ++t) s += partials[((long long)b * ntiles + t) * P + p];
    const double* hb = hp + (long long)b * hp_stride;
    double g;
    if (p == 0) g = s / hb[0];                                  // (1/sig) sum M
    else if (p <= d) g = (hb[p] != 0.0) ? -s / hb[p] : 0.0;     // -ls_k sum M dx^2 = -(1/ls_k) sum M da^2
    else g = hb[p] * s;                                         // sig_n tr(W)
    grad[(long long)b * P + p] = g;
}

static inline long long lml_feat_bytes(int dpad) { return (long long)TILE * XLD * 8 + feat_smem_bytes(dpad); }

extern "C" long long gpr_lml_grad_scratch_bytes(int B, int npad, int d) {
    const long long nt = npad / TILE;
    const long long ntiles = nt * (nt + 1) / 2;
    int sms = gpr_device_sm_count();
    if (sms <= 0) sms = 148;
    return trtri_work_bytes(B, npad, sms) + (long long)B * nt * TILE * TILE * 8 + align256((long long)B * ntiles * (d + 2) * 8);
}

static int lml_grad_impl(double* L, const double* invdiag, const double* a, const double* nrm, const double* alpha,
                         const double* hp, int hp_stride, int has_noise, int B, int n, int npad, int d, int dpad,
                         double* grad, void* scratch, void* stream, int gt, const double* Kinv_g, int ldg) {
    ARG_CHECK(L && invdiag && a && nrm && alpha && hp && grad && scratch);
    ARG_CHECK(B > 0 && n > 0 && d > 0 && npad % TILE == 0 && npad >= n && dpad % 4 == 0 && dpad >= d);
    const long long smem_l = lml_feat_bytes(dpad);
    if (smem_l > 227 * 1024 || d + 2 > NTHREADS)
        return fail(GPR_ERR_UNSUPPORTED, "input dimension too large for the fused gradient%s (dpad=%lld)", "", dpad);
    cudaStream_t st = (cudaStream_t)stream;
    const int nt = npad / TILE;
    const int ntiles = nt * (nt + 1) / 2;
    const int P = d + 1 + (has_noise ? 1 : 0);
    int sms = gpr_device_sm_count();
    if (sms <= 0) return fail(GPR_ERR_CUDA, "no CUDA device%s", "");
    double* invdiagT = reinterpret_cast<double*>(reinterpret_cast<char*>(scratch) + trtri_work_bytes(B, npad, sms));
    double* partials = invdiagT + (long long)B * nt * TILE * TILE;
    transpose_invdiag_kernel<<<dim3(TILE / 32, TILE / 32, B * nt), dim3(32, 8), 0, st>>>(invdiag, invdiagT);
    {
        const int rc = launch_trtri(L, invdiag, invdiagT, B, npad, scratch, st, gt);
        if (rc != GPR_OK) return rc;
    }
    const int smem = (int)(smem_l > MAIN_SMEM_BYTES ? smem_l : MAIN_SMEM_BYTES);
    CUDA_TRY(opt_in_smem(lauum_grad_kernel, smem));
    lauum_grad_kernel<<<dim3(ntiles, B), NTHREADS, smem, st>>>(L, invdiagT, a, nrm, alpha, hp, hp_stride, n, npad, d, dpad,
                                                               partials, P, gt, Kinv_g, ldg);
    grad_finalize_kernel<<<B, ((P + 31) / 32) * 32, 0, st>>>(partials, ntiles, P, hp, hp_stride, d, grad);
    CUDA_TRY(cudaGetLastError());
    return GPR_OK;
}

// y[b, i] -= A[b, i, :] . z[b, :]  — one warp per row, 8 rows per block; the right-hand-side update of the panel-wise forward solve
__global__ void __launch_bounds__(256) gemv_sub_kernel(const double* __restrict__ A, long long lda, long long a_batch_stride, int rows,
                                                       int kdim, const double* __restrict__ z, long long z_stride,
                                                       double* __restrict__ y, long long y_stride) {
    const int b = blockIdx.y, lane = threadIdx.x & 31;
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= rows) return;
    const double* a = A + (long long)b * a_batch_stride + (long long)i * lda;
    const double* zb = z + (long long)b * z_stride;
    double s = 0.0;
    for (int k = lane * 2; k < kdim; k += 64) {
        const double2 av = *reinterpret_cast<const double2*>(a + k);
        const double2 zv = *reinterpret_cast<const double2*>(zb + k);
        s = fma(av.x, zv.x, s);
        s = fma(av.y, zv.y, s);
    }
    s = warp_sum(s);
    if (lane == 0) y[(long long)b * y_stride + i] -= s;
}

extern "C" int gpr_gemv_sub(const double* A, long long lda, long long a_batch_stride, int B, int rows, int kdim, const double* z,
                            long long z_stride, double* y, long long y_stride, void* stream) {
    ARG_CHECK(A && z && y && B > 0 && rows > 0 && kdim > 0 && kdim % 2 == 0 && lda >= kdim && lda % 2 == 0 && a_batch_stride % 2 == 0 &&
              z_stride % 2 == 0);
    gemv_sub_kernel<<<dim3((rows + 7) / 8, B), 256, 0, (cudaStream_t)stream>>>(A, lda, a_batch_stride, rows, kdim, z, z_stride, y, y_stride);
    CUDA_TRY(cudaGetLastError());
    return GPR_OK;
}

// ---- emulated-FP64 gradient pieces (engine.lml_grad_emulated): K^-1 = U U^T comes from gpr_ozaki_gemm_nt ---------------------------
__global__ void triu_rows_kernel(const double* __restrict__ L, const double* __restrict__ invdiag, int npad, double* __restrict__ Ud) {
    // Ud[r][c] = U[r][c]: block column > block row from the strict block-upper part of L, diagonal blocks = inv(L_ii)^T, zeros below
    __shared__ double t[32][33];
    const int b = blockIdx.z;
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    const int br = r0 / TILE, bc = c0 / TILE, nt = npad / TILE;
    const double* Lb = L + (long long)b * npad * npad;
    double* Ub = Ud + (long long)b * npad * npad;
    if (bc < br) {
        for (int r = threadIdx.y; r < 32; r += blockDim.y) Ub[(long long)(r0 + r) * npad + c0 + threadIdx.x] = 0.0;
    } else if (bc > br) {
        for (int r = threadIdx.y; r < 32; r += blockDim.y)
            Ub[(long long)(r0 + r) * npad + c0 + threadIdx.x] = Lb[(long long)(r0 + r) * npad + c0 + threadIdx.x];
    } else {
        const double* inv = invdiag + ((long long)b * nt + br) * TILE * TILE;
        const int rr = r0 - br * TILE, cc = c0 - bc * TILE;
        for (int r = threadIdx.y; r < 32; r += blockDim.y) t[r][threadIdx.x] = inv[(cc + r) * TILE + rr + threadIdx.x];   // inv[c][r]
        __syncthreads();
        for (int r = threadIdx.y; r < 32; r += blockDim.y) Ub[(long long)(r0 + r) * npad + c0 + threadIdx.x] = t[threadIdx.x][r];
    }
}

extern "C" int gpr_triu_rows(const double* L, const double* invdiag, int B, int npad, double* Ud, void* stream) {
    ARG_CHECK(L && invdiag && Ud && B > 0 && npad > 0 && npad % TILE == 0);
    triu_rows_kernel<<<dim3(npad / 32, npad / 32, B), dim3(32, 8), 0, (cudaStream_t)stream>>>(L, invdiag, npad, Ud);
    CUDA_TRY(cudaGetLastError());
    return GPR_OK;
}

extern "C" int gpr_grad_from_kinv(const double* Kinv, int ldk, const double* a, const double* nrm, const double* alpha,
                                  const double* hp, int hp_stride, int has_noise, int B, int n, int npad, int d, int dpad,
                                  double* grad, void* scratch, void* stream) {
    ARG_CHECK(Kinv && a && nrm && alpha && hp && grad && scratch);
    ARG_CHECK(B > 0 && n > 0 && d > 0 && npad % TILE == 0 && npad >= n && dpad % 4 == 0 && dpad >= d && ldk >= npad && ldk % 2 == 0);
    const long long smem_l = lml_feat_bytes(dpad);
    if (smem_l > 227 * 1024 || d + 2 > NTHREADS)
        return fail(GPR_ERR_UNSUPPORTED, "input dimension too large for the fused gradient%s (dpad=%lld)", "", dpad);
    cudaStream_t st = (cudaStream_t)stream;
    const int nt = npad / TILE;
    const int ntiles = nt * (nt + 1) / 2;
    const int P = d + 1 + (has_noise ? 1 : 0);
    double* partials = reinterpret_cast<double*>(scratch);
    const int smem = (int)(smem_l > MAIN_SMEM_BYTES ? smem_l : MAIN_SMEM_BYTES);
    CUDA_TRY(opt_in_smem(lauum_grad_kernel, smem));
    lauum_grad_kernel<<<dim3(ntiles, B), NTHREADS, smem, st>>>(nullptr, nullptr, a, nrm, alpha, hp, hp_stride, n, npad, d, dpad,
                                                               partials, P, 0, nullptr, 0, Kinv, ldk);
    grad_finalize_kernel<<<B, ((P + 31) / 32) * 32, 0, st>>>(partials, ntiles, P, hp, hp_stride, d, grad);
    CUDA_TRY(cudaGetLastError());
    return GPR_OK;
}

extern "C" int gpr_lml_grad(double* L, const double* invdiag, const double* a, const double* nrm, const double* alpha,
                            const double* hp, int hp_stride, int has_noise, int B, int n, int npad, int d, int dpad,
                            double* grad, void* scratch, void* stream) {
    return lml_grad_impl(L, invdiag, a, nrm, alpha, hp, hp_stride, has_noise, B, n, npad, d, dpad, grad, scratch, stream, 0, nullptr, 0);
}

// Gradient for experts factorised by gpr_potrf_batched_shared WHEN the leading block copied into L also carries the
// communication expert's L_g^-T in its strict block-upper part (gpr_potri_batched / gpr_tril_inverse was run on the
// communication factor before the copy).  Kinv_g [ng, ldg] = K_g^-1 from gpr_potri_batched.  1/8 fewer flops at ng = n/2.
extern "C" int gpr_lml_grad_shared(double* L, const double* invdiag, const double* a, const double* nrm, const double* alpha,
                                   const double* hp, int hp_stride, int has_noise, int B, int n, int npad, int d, int dpad,
                                   double* grad, void* scratch, void* stream, int ng, const double* Kinv_g, int ldg) {
    ARG_CHECK(Kinv_g && ng > 0 && ng % TILE == 0 && ng < n && ldg >= ng && ldg % 2 == 0);
    return lml_grad_impl(L, invdiag, a, nrm, alpha, hp, hp_stride, has_noise, B, n, npad, d, dpad, grad, scratch, stream, ng / TILE,
                         Kinv_g, ldg);
}

// ---------------------------------------------------------------------------------------------
// Communication expert in GEMM form (shared-leading-block prediction).  The substitution of predict_kernel gives one
// serial task per 128 test points — ceil(m/128) busy SMs while the local experts wait for V_g^T.  With X = L^-1 formed once
// per fit (gpr_tril_inverse: L^-T by the blocked trtri, transposed into a lower-triangular row-major matrix), the same
// V_g^T = K* X^T is a triangular GEMM with (m/128) x (n/128) independent tiles:
//   vt_gemm_kernel     : V^T[ct, :, j] = sum_{k <= j} K*[ct, k] X[j, k]^T  + per-row partials of |V|^2 and V^T z
//   vt_finalize_kernel : mean = sum_j partials (= V^T z = K* alpha),  var = prior - sum_j |V_j|^2   (fixed order, no atomics)
// ---------------------------------------------------------------------------------------------
__global__ void tril_from_upper_kernel(const double* __restrict__ L, const double* __restrict__ invdiag, int npad,
                                       double* __restrict__ X) {
    // X[r][c] = U[c][r] for block row > block column (U = L^-T sits in the strict block-upper part of L), diagonal blocks
    // from invdiag, zero above.  32 x 32 tiles through shared memory.
    __shared__ double t[32][33];
    const int b = blockIdx.z;
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    const int br = r0 / TILE, bc = c0 / TILE, nt = npad / TILE;
    const double* Lb = L + (long long)b * npad * npad;
    double* Xb = X + (long long)b * npad * npad;
    if (bc > br) {
        for (int r = threadIdx.y; r < 32; r += blockDim.y) Xb[(long long)(r0 + r) * npad + c0 + threadIdx.x] = 0.0;
    } else if (bc == br) {
        const double* inv = invdiag + ((long long)b * nt + br) * TILE * TILE;
        for (int r = threadIdx.y; r < 32; r += blockDim.y)
            Xb[(long long)(r0 + r) * npad + c0 + threadIdx.x] = inv[(r0 - br * TILE + r) * TILE + (c0 - bc * TILE + threadIdx.x)];
    } else {
        for (int r = threadIdx.y; r < 32; r += blockDim.y) t[r][threadIdx.x] = Lb[(long long)(c0 + r) * npad + r0 + threadIdx.x];
        __syncthreads();
        for (int r = threadIdx.y; r < 32; r += blockDim.y) Xb[(long long)(r0 + r) * npad + c0 + threadIdx.x] = t[threadIdx.x][r];
    }
}

extern "C" int gpr_tril_inverse(double* L, const double* invdiag, int B, int npad, double* X, void* scratch, void* stream) {
    ARG_CHECK(L && invdiag && X && scratch && B > 0 && npad > 0 && npad % TILE == 0);
    cudaStream_t st = (cudaStream_t)stream;
    const int nt = npad / TILE;
    int sms = gpr_device_sm_count();
    if (sms <= 0) return fail(GPR_ERR_CUDA, "no CUDA device%s", "");
    double* invdiagT = reinterpret_cast<double*>(reinterpret_cast<char*>(scratch) + trtri_work_bytes(B, npad, sms));
    transpose_invdiag_kernel<<<dim3(TILE / 32, TILE / 32, B * nt), dim3(32, 8), 0, st>>>(invdiag, invdiagT);
    {
        const int rc = launch_trtri(L, invdiag, invdiagT, B, npad, scratch, st);
        if (rc != GPR_OK) return rc;
    }
    tril_from_upper_kernel<<<dim3(npad / 32, npad / 32, B), dim3(32, 8), 0, st>>>(L, invdiag, npad, X);
    CUDA_TRY(cudaGetLastError());
    return GPR_OK;
}

// ---------------------------------------------------------------------------------------------
// Emulated-FP64 prediction (opt-in, engine.predict_shared_emulated): every local expert's V_l^T = K* X_l^T is ONE GEMM against the
// rows X_l = (L^-1)[row0:, :] of the explicit inverse, run on the INT8 tensor cores by ozaki_gemm.cu.  The pieces that live here:
//   gpr_trtri_batched      U = L^-T of every expert, in place in the strict block-upper part of L (block columns < gt — the shared
//                          communication block — must already hold L_g^-T: gpr_copy_leading_block mode 2)
//   gpr_tril_inverse_rows  X[b] = rows row0.. of L_b^-1 as a dense row-major [npad - row0, npad] matrix (transposed out of U,
//                          diagonal blocks from invdiag, zeros right of the diagonal block) — the operand the digit split reads
//   gpr_vt_reduce          mean_i = mean_g + V_l^T z_l,  var_i = var_g - |V_l|^2 per test point   (gpr.py:80-85,100-104 for an
//                          expert that extends the communication expert); fixed-order sums, no atomics
// ---------------------------------------------------------------------------------------------
extern "C" int gpr_trtri_batched(double* L, const double* invdiag, int B, int npad, void* scratch, void* stream, int gt, int seg) {
    ARG_CHECK(L && invdiag && scratch && B > 0 && npad > 0 && npad % TILE == 0 && gt >= 0 && gt < npad / TILE && seg >= 0);
    cudaStream_t st = (cudaStream_t)stream;
    const int nt = npad / TILE;
    int sms = gpr_device_sm_count();
    if (sms <= 0) return fail(GPR_ERR_CUDA, "no CUDA device%s", "");
    double* invdiagT = reinterpret_cast<double*>(reinterpret_cast<char*>(scratch) + trtri_work_bytes(B, npad, sms));
    transpose_invdiag_kernel<<<dim3(TILE / 32, TILE / 32, B * nt), dim3(32, 8), 0, st>>>(invdiag, invdiagT);
    return launch_trtri(L, invdiag, invdiagT, B, npad, scratch, st, gt, seg);
}

// A window [r0, r0+nr) x [c0, c0+nc) of a factor-derived matrix as a dense row-major block out[b][nr][nc] (all multiples of 32):
//   kind 0: U = L^-T   (block column > block row: the strict block-upper part of L; diagonal blocks inv(L_ii)^T; zeros below)
//   kind 1: X = L^-1   (X[r][c] = U[c][r])
// — the operands of the products that assemble L^-T from its diagonal segments (engine.inverse_factors) and of K^-1 = U U^T.
__global__ void factor_window_kernel(const double* __restrict__ L, const double* __restrict__ invdiag, int npad, int kind, int r0,
                                     int nr, int c0, int nc, double* __restrict__ out) {
    __shared__ double t[32][33];
    const int b = blockIdx.z;
    const int wr = blockIdx.y * 32, wc = blockIdx.x * 32;          // inside the window
    const int gr = r0 + wr, gc = c0 + wc;                          // in the matrix
    const int br = gr / TILE, bc = gc / TILE, nt = npad / TILE;
    const double* Lb = L + (long long)b * npad * npad;
    double* ob = out + (long long)b * nr * nc + (long long)wr * nc + wc;
    const bool zero = kind == 0 ? bc < br : bc > br;
    const bool diag = bc == br;
    if (zero) {
        for (int r = threadIdx.y; r < 32; r += blockDim.y) ob[(long long)r * nc + threadIdx.x] = 0.0;
    } else if (diag && kind == 1) {                                // X_ii = inv(L_ii)
        const double* inv = invdiag + ((long long)b * nt + br) * TILE * TILE;
        for (int r = threadIdx.y; r < 32; r += blockDim.y)
            ob[(long long)r * nc + threadIdx.x] = inv[(gr - br * TILE + r) * TILE + (gc - bc * TILE + threadIdx.x)];
    } else if (diag) {                                             // U_ii = inv(L_ii)^T
        const double* inv = invdiag + ((long long)b * nt + br) * TILE * TILE;
        const int rr = gr - br * TILE, cc = gc - bc * TILE;
        for (int r = threadIdx.y; r < 32; r += blockDim.y) t[r][threadIdx.x] = inv[(cc + r) * TILE + rr + threadIdx.x];
        __syncthreads();
        for (int r = threadIdx.y; r < 32; r += blockDim.y) ob[(long long)r * nc + threadIdx.x] = t[threadIdx.x][r];
    } else if (kind == 0) {                                        // U above the diagonal: stored as is
        for (int r = threadIdx.y; r < 32; r += blockDim.y) ob[(long long)r * nc + threadIdx.x] = Lb[(long long)(gr + r) * npad + gc + threadIdx.x];
    } else {                                                       // X below the diagonal: U transposed
        for (int r = threadIdx.y; r < 32; r += blockDim.y) t[r][threadIdx.x] = Lb[(long long)(gc + r) * npad + gr + threadIdx.x];
        __syncthreads();
        for (int r = threadIdx.y; r < 32; r += blockDim.y) ob[(long long)r * nc + threadIdx.x] = t[threadIdx.x][r];
    }
}

extern "C" int gpr_factor_window(const double* L, const double* invdiag, int B, int npad, int kind, int r0, int nr, int c0, int nc,
                                 double* out, void* stream) {
    ARG_CHECK(L && invdiag && out && B > 0 && npad > 0 && npad % TILE == 0 && (kind == 0 || kind == 1));
    ARG_CHECK(r0 >= 0 && c0 >= 0 && nr > 0 && nc > 0 && r0 % 32 == 0 && c0 % 32 == 0 && nr % 32 == 0 && nc % 32 == 0 &&
              r0 + nr <= npad && c0 + nc <= npad);
    factor_window_kernel<<<dim3(nc / 32, nr / 32, B), dim3(32, 8), 0, (cudaStream_t)stream>>>(L, invdiag, npad, kind, r0, nr, c0, nc, out);
    CUDA_TRY(cudaGetLastError());
    return GPR_OK;
}

__global__ void tril_rows_from_upper_kernel(const double* __restrict__ L, const double* __restrict__ invdiag, int npad, int row0,
                                            double* __restrict__ X) {
    __shared__ double t[32][33];
    const int b = blockIdx.z;
    const int r0 = row0 + blockIdx.y * 32, c0 = blockIdx.x * 32;
    const int br = r0 / TILE, bc = c0 / TILE, nt = npad / TILE;
    const double* Lb = L + (long long)b * npad * npad;
    double* Xb = X + (long long)b * (npad - row0) * npad + (long long)(r0 - row0) * npad;
    if (bc > br) {
        for (int r = threadIdx.y; r < 32; r += blockDim.y) Xb[(long long)r * npad + c0 + threadIdx.x] = 0.0;
    } else if (bc == br) {
        const double* inv = invdiag + ((long long)b * nt + br) * TILE * TILE;
        for (int r = threadIdx.y; r < 32; r += blockDim.y)
            Xb[(long long)r * npad + c0 + threadIdx.x] = inv[(r0 - br * TILE + r) * TILE + (c0 - bc * TILE + threadIdx.x)];
    } else {
        for (int r = threadIdx.y; r < 32; r += blockDim.y) t[r][threadIdx.x] = Lb[(long long)(c0 + r) * npad + r0 + threadIdx.x];
        __syncthreads();
        for (int r = threadIdx.y; r < 32; r += blockDim.y) Xb[(long long)r * npad + c0 + threadIdx.x] = t[threadIdx.x][r];
    }
}

extern "C" int gpr_tril_inverse_rows(const double* L, const double* invdiag, int B, int npad, int row0, double* X, void* stream) {
    ARG_CHECK(L && invdiag && X && B > 0 && npad > 0 && npad % TILE == 0 && row0 >= 0 && row0 < npad && row0 % 32 == 0);
    tril_rows_from_upper_kernel<<<dim3(npad / 32, (npad - row0) / 32, B), dim3(32, 8), 0, (cudaStream_t)stream>>>(L, invdiag, npad, row0, X);
    CUDA_TRY(cudaGetLastError());
    return GPR_OK;
}

// one warp per (expert, test point): V [E, m, ldv] rows of length nl
__global__ void __launch_bounds__(256) vt_reduce_kernel(const double* __restrict__ V, long long ldv, const double* __restrict__ z,
                                                        long long z_stride, const double* __restrict__ mean_g,
                                                        const double* __restrict__ var_g, int m, int nl, double* __restrict__ mean,
                                                        double* __restrict__ var) {
    const int e = blockIdx.y, lane = threadIdx.x & 31;
    const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (t >= m) return;
    const double* v = V + ((long long)e * m + t) * ldv;
    const double* ze = z + (long long)e * z_stride;
    double sm = 0.0, sv = 0.0;
    for (int c = lane * 2; c < nl; c += 64) {
        const double2 a = *reinterpret_cast<const double2*>(v + c);
        const double2 b = *reinterpret_cast<const double2*>(ze + c);
        sm = fma(a.x, b.x, sm); sm = fma(a.y, b.y, sm);
        sv = fma(a.x, a.x, sv); sv = fma(a.y, a.y, sv);
    }
    for (int o = 16; o; o >>= 1) {
        sm += __shfl_xor_sync(0xffffffffu, sm, o);
        sv += __shfl_xor_sync(0xffffffffu, sv, o);
    }
    if (lane == 0) {
        mean[(long long)e * m + t] = mean_g[t] + sm;
        var[(long long)e * m + t] = var_g[t] - sv;
    }
}

extern "C" int gpr_vt_reduce(const double* V, long long ldv, const double* z, long long z_stride, const double* mean_g,
                             const double* var_g, int E, int m, int nl, double* mean, double* var, void* stream) {
    ARG_CHECK(V && z && mean_g && var_g && mean && var && E > 0 && m > 0 && nl > 0 && nl % 2 == 0 && ldv >= nl && ldv % 2 == 0);
    vt_reduce_kernel<<<dim3((m + 7) / 8, E), 256, 0, (cudaStream_t)stream>>>(V, ldv, z, z_stride, mean_g, var_g, m, nl, mean, var);
    CUDA_TRY(cudaGetLastError());
    return GPR_OK;
}

__global__ void __launch_bounds__(NTHREADS, 1)
vt_gemm_kernel(const double* __restrict__ Ks, const double* __restrict__ X, const double* __restrict__ z, int npad, int mpad,
               double* __restrict__ vt, double* __restrict__ pv, double* __restrict__ pm) {
    extern __shared__ double smem[];
    __shared__ unsigned long long pipe_bars[2 * STAGES];
    __shared__ double zj[TILE];
    __shared__ double red[2][2][TILE];
    Pipe pipe;
    pipe_init(pipe, pipe_bars);
    const int nt = npad / TILE;
    const int j = nt - 1 - blockIdx.x, ct = blockIdx.y, b = blockIdx.z;  // long k-ranges first
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < TILE) zj[tid] = z[(long long)b * npad + j * TILE + tid];
    double acc[4][8][2];
    zero_acc(acc);
    NoHook nh;
    Seg none{nullptr, 0, nullptr, 0, 0};
    Seg s0{Ks + ((long long)b * mpad + (long long)ct * TILE) * npad, npad, X + (long long)b * npad * npad + ((long long)j * TILE) * npad, npad,
           (j + 1) * (TILE / KC)};
    mainloop<false, XPF_PREDICT>(acc, s0, none, smem, nh, pipe);
    double* Vt = vt + (((long long)b * (mpad / TILE) + ct) * TILE) * npad + (long long)j * TILE;
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
        double vv = 0.0, mv = 0.0;
#pragma unroll
        for (int ni = 0; ni < 8; ++ni) {
            vv += acc[mi][ni][0] * acc[mi][ni][0] + acc[mi][ni][1] * acc[mi][ni][1];
            mv += acc[mi][ni][0] * zj[acc_col(ni, 0)] + acc[mi][ni][1] * zj[acc_col(ni, 1)];
            *reinterpret_cast<double2*>(Vt + (long long)acc_row(mi) * npad + acc_col(ni, 0)) = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
        }
        vv += __shfl_xor_sync(0xffffffffu, vv, 1);
        vv += __shfl_xor_sync(0xffffffffu, vv, 2);
        mv += __shfl_xor_sync(0xffffffffu, mv, 1);
        mv += __shfl_xor_sync(0xffffffffu, mv, 2);
        if ((lane & 3) == 0) {
            red[0][warp >> 2][acc_row(mi)] = mv;
            red[1][warp >> 2][acc_row(mi)] = vv;
        }
    }
    __syncthreads();
    if (tid < TILE) {
        const long long o = ((long long)b * nt + j) * mpad + (long long)ct * TILE + tid;
        pm[o] = red[0][0][tid] + red[0][1][tid];
        pv[o] = red[1][0][tid] + red[1][1][tid];
    }
}

__global__ void vt_finalize_kernel(const double* __restrict__ pv, const double* __restrict__ pm, const double* __restrict__ prior,
                                   int nt, int mpad, int m, double* __restrict__ mean, double* __restrict__ var, int gt_part,
                                   double* __restrict__ mean_part, double* __restrict__ var_part, int prior_per_point = 0) {
    const int b = blockIdx.y, r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    // prior variance k(x*, x*): one value per expert (stationary kernels) or one per test point (gpr.py:98,102 in general)
    const double pr = prior_per_point ? prior[(long long)b * m + r] : prior[b];
    double sm = 0.0, sv = 0.0;
    for (int j = 0; j < nt; ++j) {
        if (j == gt_part && mean_part) {  // the part of the sums that belongs to the first gt_part block rows (the shared ones)
            mean_part[(long long)b * m + r] = sm;
            var_part[(long long)b * m + r] = pr - sv;
        }
        sm += pm[((long long)b * nt + j) * mpad + r];
        sv += pv[((long long)b * nt + j) * mpad + r];
    }
    if (gt_part >= nt && mean_part) {
        mean_part[(long long)b * m + r] = sm;
        var_part[(long long)b * m + r] = pr - sv;
    }
    mean[(long long)b * m + r] = sm;
    var[(long long)b * m + r] = pr - sv;
}

static inline long long predict_gemm_layout(int B, int npad, int dpad, int mpad, long long* o_nrm, long long* o_ks, long long* o_pv,
                                            long long* o_pm) {
    long long off = align256((long long)B * mpad * dpad * 8);  // scaled test points
    *o_nrm = off; off += align256((long long)B * mpad * 8);
    *o_ks = off;  off += align256((long long)B * mpad * npad * 8);
    *o_pv = off;  off += align256((long long)B * (npad / TILE) * mpad * 8);
    *o_pm = off;  off += align256((long long)B * (npad / TILE) * mpad * 8);
    return off;
}

extern "C" long long gpr_predict_gemm_scratch_bytes(int B, int npad, int dpad, int m) {
    const int mpad = (m + TILE - 1) / TILE * TILE;
    long long a, b, c, d;
    return predict_gemm_layout(B, npad, dpad, mpad, &a, &b, &c, &d);
}

extern "C" int gpr_predict_gemm(const double* X, const double* a, const double* nrm, const double* z, const double* hp, int hp_stride,
                                const double* prior, const double* xs, int B, int n, int npad, int d, int dpad, int m, double* mean,
                                double* var, double* vt, void* scratch, void* stream, int rows_part, double* mean_part,
                                double* var_part) {
    ARG_CHECK(X && a && nrm && z && hp && prior && xs && mean && var && vt && scratch);
    ARG_CHECK((!mean_part && !var_part) || (mean_part && var_part && rows_part > 0 && rows_part % TILE == 0 && rows_part <= npad));
    ARG_CHECK(B > 0 && n > 0 && m > 0 && d > 0 && npad % TILE == 0 && npad >= n && dpad % 4 == 0 && dpad >= d);
    const int fsm = feat_smem_bytes(dpad);
    if (fsm > 227 * 1024) return fail(GPR_ERR_UNSUPPORTED, "input dimension too large for the fused SE tile%s (dpad=%lld)", "", dpad);
    cudaStream_t st = (cudaStream_t)stream;
    const int mpad = (m + TILE - 1) / TILE * TILE, nt = npad / TILE;
    long long o_nrm, o_ks, o_pv, o_pm;
    predict_gemm_layout(B, npad, dpad, mpad, &o_nrm, &o_ks, &o_pv, &o_pm);
    char* base = reinterpret_cast<char*>(scratch);
    double* as = reinterpret_cast<double*>(base);
    double* nrms = reinterpret_cast<double*>(base + o_nrm);
    double* Ks = reinterpret_cast<double*>(base + o_ks);
    double* pv = reinterpret_cast<double*>(base + o_pv);
    double* pm = reinterpret_cast<double*>(base + o_pm);
    // scaled test points (shared xs, per-expert length-scales), K* [B, mpad, npad] (padding rows / columns zero)
    se_scale_kernel<<<dim3((mpad + 127) / 128, B), 128, 0, st>>>(xs, 0, hp, hp_stride, m, d, mpad, dpad, as, nrms);
    CUDA_TRY(cudaMemsetAsync(Ks, 0, (size_t)B * mpad * npad * 8, st));
    CUDA_TRY(opt_in_smem(se_tile_kernel, fsm));
    se_tile_kernel<<<dim3(nt, mpad / TILE, B), NTHREADS, fsm, st>>>(as, nrms, mpad, a, nrm, npad, hp, hp_stride, nullptr, m, n, dpad, Ks,
                                                                    npad, (long long)mpad * npad, 1, 0, 0);
    CUDA_TRY(opt_in_smem(vt_gemm_kernel, MAIN_SMEM_BYTES));
    vt_gemm_kernel<<<dim3(nt, mpad / TILE, B), NTHREADS, MAIN_SMEM_BYTES, st>>>(Ks, X, z, npad, mpad, vt, pv, pm);
    vt_finalize_kernel<<<dim3((m + 127) / 128, B), 128, 0, st>>>(pv, pm, prior, nt, mpad, m, mean, var, rows_part / TILE, mean_part, var_part);
    CUDA_TRY(cudaGetLastError());
    return GPR_OK;
}

// ---------------------------------------------------------------------------------------------
// Dense path: any covariance the Covar protocol can express (e.g. Compose([SE, SE, WN]), covar.py:28-81) at the model level.
// The kernel matrices come materialised from the kernel-level entry points (gpr_se_build / gpr_se_cross summed by the host
// layer, as Compose.kernel does), K is factorised in place (gpr_potrf_batched, gen = 0), and
//   gpr_predict_dense : V^T = K* X^T with X = L^-1 (vt_gemm_kernel), mean = V^T z = K* alpha, var = prior[b, i] - |V_i|^2   gpr.py:76-106
//   gpr_covar_dense   : cov = K** - V^T V                                                                            gpr.py:108-120
//   gpr_trace_grad    : grad_p = 1/2 sum_ij (K^-1 - alpha alpha^T)_ij dK_p,ij  with dK materialised like the reference  loss.py:77-84
// O(P n^2) memory like the reference — the fused path above is the one for large n.
// ---------------------------------------------------------------------------------------------
extern "C" long long gpr_predict_dense_scratch_bytes(int B, int npad, int m) {
    const long long mpad = (m + TILE - 1) / TILE * TILE;
    return 2 * align256((long long)B * (npad / TILE) * mpad * 8);
}

extern "C" int gpr_predict_dense(const double* X, const double* Ks, const double* z, const double* prior, int B, int npad, int m,
                                 double* mean, double* var, double* vt, void* scratch, void* stream) {
    ARG_CHECK(X && Ks && z && prior && mean && var && vt && scratch && B > 0 && m > 0 && npad > 0 && npad % TILE == 0);
    cudaStream_t st = (cudaStream_t)stream;
    const int mpad = (m + TILE - 1) / TILE * TILE, nt = npad / TILE;
    double* pv = reinterpret_cast<double*>(scratch);
    double* pm = reinterpret_cast<double*>(reinterpret_cast<char*>(scratch) + align256((long long)B * nt * mpad * 8));
    CUDA_TRY(opt_in_smem(vt_gemm_kernel, MAIN_SMEM_BYTES));
    vt_gemm_kernel<<<dim3(nt, mpad / TILE, B), NTHREADS, MAIN_SMEM_BYTES, st>>>(Ks, X, z, npad, mpad, vt, pv, pm);
    vt_finalize_kernel<<<dim3((m + 127) / 128, B), 128, 0, st>>>(pv, pm, prior, nt, mpad, m, mean, var, 0, nullptr, nullptr, 1);
    CUDA_TRY(cudaGetLastError());
    return GPR_OK;
}

__global__ void __launch_bounds__(NTHREADS, 1)
covar_dense_kernel(const double* __restrict__ vt, const double* __restrict__ Kss, int m, int mpad, int npad, double* __restrict__ cov) {
    extern __shared__ double smem[];
    __shared__ unsigned long long pipe_bars[2 * STAGES];
    Pipe pipe;
    pipe_init(pipe, pipe_bars);
    const int tj = blockIdx.x, ti = blockIdx.y, b = blockIdx.z;
    const double* Vb = vt + (long long)b * mpad * npad;
    double acc[4][8][2];
    zero_acc(acc);
    NoHook nh;
    Seg none{nullptr, 0, nullptr, 0, 0};
    Seg s0{Vb + (long long)ti * TILE * npad, npad, Vb + (long long)tj * TILE * npad, npad, npad / KC};
    mainloop<false, XPF_PREDICT>(acc, s0, none, smem, nh, pipe);
    const double* kb = Kss + (long long)b * m * m;
    double* cb = cov + (long long)b * m * m;
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 8; ++ni)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const long long r = (long long)ti * TILE + acc_row(mi), c = (long long)tj * TILE + acc_col(ni, e);
                if (r < m && c < m) cb[r * m + c] = kb[r * m + c] - acc[mi][ni][e];
            }
}

extern "C" int gpr_covar_dense(const double* vt, const double* Kss, int B, int npad, int m, double* cov, void* stream) {
    ARG_CHECK(vt && Kss && cov && B > 0 && m > 0 && npad > 0 && npad % TILE == 0);
    const int mpad = (m + TILE - 1) / TILE * TILE;
    CUDA_TRY(opt_in_smem(covar_dense_kernel, MAIN_SMEM_BYTES));
    covar_dense_kernel<<<dim3(mpad / TILE, mpad / TILE, B), NTHREADS, MAIN_SMEM_BYTES, (cudaStream_t)stream>>>(vt, Kss, m, mpad, npad, cov);
    CUDA_TRY(cudaGetLastError());
    return GPR_OK;
}

constexpr int TG_ROWS = 8;  // rows of W o dK_p per block of trace_grad_kernel
__global__ void trace_grad_kernel(const double* __restrict__ Kinv, const double* __restrict__ alpha, int alpha_stride,
                                  const double* __restrict__ dK, int P, int n, double* __restrict__ partial, int nchunks) {
    __shared__ double red[NTHREADS / 32];
    const int chunk = blockIdx.x, p = blockIdx.y, b = blockIdx.z;
    const double* Kb = Kinv + (long long)b * n * n;
    const double* Db = dK + ((long long)b * P + p) * n * n;
    const double* al = alpha + (long long)b * alpha_stride;
    double s = 0.0;
    for (int r = chunk * TG_ROWS; r < (chunk + 1) * TG_ROWS && r < n; ++r) {
        const double ar = al[r];
        for (int c = threadIdx.x; c < n; c += NTHREADS) s += (Kb[(long long)r * n + c] - ar * al[c]) * Db[(long long)r * n + c];
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < NTHREADS / 32; ++w) t += red[w];
        partial[((long long)b * P + p) * nchunks + chunk] = t;
    }
}
extern "C" long long gpr_trace_grad_scratch_bytes(int B, int P, int n) {
    return align256((long long)B * P * ((n + TG_ROWS - 1) / TG_ROWS) * 8);
}

__global__ void trace_grad_sum_kernel(const double* __restrict__ partial, int BP, int nchunks, double* __restrict__ grad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= BP) return;
    double s = 0.0;
    for (int c = 0; c < nchunks; ++c) s += partial[(long long)i * nchunks + c];  // fixed order
    grad[i] = 0.5 * s;
}

extern "C" int gpr_trace_grad(const double* Kinv, const double* alpha, int alpha_stride, const double* dK, int B, int P, int n,
                              double* grad, void* scratch, void* stream) {
    ARG_CHECK(Kinv && alpha && dK && grad && scratch && B > 0 && P > 0 && n > 0 && alpha_stride >= n);
    cudaStream_t st = (cudaStream_t)stream;
    const int nchunks = (n + TG_ROWS - 1) / TG_ROWS;
    double* partial = reinterpret_cast<double*>(scratch);
    trace_grad_kernel<<<dim3(nchunks, P, B), NTHREADS, 0, st>>>(Kinv, alpha, alpha_stride, dK, P, n, partial, nchunks);
    trace_grad_sum_kernel<<<(B * P + 127) / 128, 128, 0, st>>>(partial, B * P, nchunks, grad);
    CUDA_TRY(cudaGetLastError());
    return GPR_OK;
}

// ---------------------------------------------------------------------------------------------
// N1: explicit inverse from the Cholesky factor (tc.cholesky_solve(identity) / cholesky_inverse, gr_bcm.py:101-112)
//   gpr_potri_batched = transpose_invdiag + trtri_kernel (U = L^-T into the block-upper part of L) + lauum_store_kernel
//   (K^-1 = U U^T, both triangles written).  grbcm_full_combine: prec = sum_c Kinv_c o (beta_c[i] + beta_c[j]) / 2.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHREADS, 1)
lauum_store_kernel(const double* __restrict__ L, const double* __restrict__ invdiagT, int n, int npad, double* __restrict__ out,
                   int ldo) {
    extern __shared__ double smem[];
    __shared__ unsigned long long pipe_bars[2 * STAGES];
    Pipe pipe;
    pipe_init(pipe, pipe_bars);
    const int nt = npad / TILE;
    const int b = blockIdx.y;
    int ti = blockIdx.x, i = 0;
    while (ti >= i + 1) { ti -= i + 1; ++i; }
    const int j = ti;
    const double* Lb = L + (long long)b * npad * npad;
    const double* Uii = invdiagT + ((long long)b * nt + i) * TILE * TILE;
    const double* Ui = Lb + ((long long)i * TILE) * npad;
    const double* Uj = Lb + ((long long)j * TILE) * npad;
    double acc[4][8][2];
    zero_acc(acc);
    NoHook nh;
    Seg s0{Uii, TILE, (i == j) ? Uii : Uj + (long long)i * TILE, (i == j) ? (long long)TILE : (long long)npad, TILE / KC};
    Seg s1{Ui + (long long)(i + 1) * TILE, npad, Uj + (long long)(i + 1) * TILE, npad, (nt - 1 - i) * (TILE / KC)};
    mainloop<false, XPF_LAUUM>(acc, s0, s1, smem, nh, pipe);
    double* ob = out + (long long)b * ldo * ldo;
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 8; ++ni)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const long long r = (long long)i * TILE + acc_row(mi), c = (long long)j * TILE + acc_col(ni, e);
                if (r < n && c < n) {
                    ob[r * ldo + c] = acc[mi][ni][e];
                    if (i != j) ob[c * ldo + r] = acc[mi][ni][e];
                }
            }
}

extern "C" long long gpr_potri_scratch_bytes(int B, int npad) {
    int sms = gpr_device_sm_count();
    if (sms <= 0) sms = 148;
    return trtri_work_bytes(B, npad, sms) + (long long)B * (npad / TILE) * TILE * TILE * 8;
}

extern "C" int gpr_potri_batched(double* L, const double* invdiag, int B, int n, int npad, double* Kinv, int ldo, void* scratch,
                                 void* stream) {
    ARG_CHECK(L && invdiag && Kinv && scratch && B > 0 && n > 0 && npad % TILE == 0 && npad >= n && ldo >= n);
    cudaStream_t st = (cudaStream_t)stream;
    const int nt = npad / TILE;
    int sms = gpr_device_sm_count();
    if (sms <= 0) return fail(GPR_ERR_CUDA, "no CUDA device%s", "");
    double* invdiagT = reinterpret_cast<double*>(reinterpret_cast<char*>(scratch) + trtri_work_bytes(B, npad, sms));
    transpose_invdiag_kernel<<<dim3(TILE / 32, TILE / 32, B * nt), dim3(32, 8), 0, st>>>(invdiag, invdiagT);
    {
        const int rc = launch_trtri(L, invdiag, invdiagT, B, npad, scratch, st);
        if (rc != GPR_OK) return rc;
    }
    CUDA_TRY(opt_in_smem(lauum_store_kernel, MAIN_SMEM_BYTES));
    lauum_store_kernel<<<dim3(nt * (nt + 1) / 2, B), NTHREADS, MAIN_SMEM_BYTES, st>>>(L, invdiagT, n, npad, Kinv, ldo);
    CUDA_TRY(cudaGetLastError());
    return GPR_OK;
}

__global__ void grbcm_full_combine_kernel(const double* __restrict__ Kinv, const double* __restrict__ beta, int nc1, int m,
                                          double* __restrict__ prec) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)m * m) return;
    const int i = idx / m, j = idx % m;
    double s = 0.0;
    for (int c = 0; c < nc1; ++c)
        s += Kinv[(long long)c * m * m + idx] * (0.5 * (beta[(long long)c * m + i] + beta[(long long)c * m + j]));  // gr_bcm.py:107-111
    prec[idx] = s;
}

extern "C" int grbcm_full_combine(const double* Kinv, const double* beta, int nc1, int m, double* prec, void* stream) {
    ARG_CHECK(Kinv && beta && prec && nc1 > 0 && m > 0);
    grbcm_full_combine_kernel<<<(unsigned)(((long long)m * m + 255) / 256), 256, 0, (cudaStream_t)stream>>>(Kinv, beta, nc1, m, prec);
    CUDA_TRY(cudaGetLastError());
    return GPR_OK;
}
