// Latency of the factorisation + inversion of ONE 128x128 diagonal tile inside a CTA (the serial link of the Cholesky's
// dependency chain): clocks of potf2 / trtri, old (8x8 register blocks) vs new (32-wide warp panels + DMMA), and the error of
// both against a host factorisation.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I pygpr_b200/csrc -I tools/microbench -o diag_tile tools/microbench/diag_tile.cu && ./diag_tile
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "gpr_core.cuh"
#include "gpr_diag_tile.cuh"
#include "diag_tile_round1.cuh"

using namespace gpr;

template <int VARIANT>
__global__ void __launch_bounds__(NTHREADS, 1) tile_kernel(const double* __restrict__ A, double* __restrict__ Lout, double* __restrict__ Xout,
                                                           long long* __restrict__ cyc, int* __restrict__ fail_out) {
    extern __shared__ double smem[];
    __shared__ double invblk[(TILE / 8) * 64];
    __shared__ int s_fail;
    double* T = smem;
    double* Bm = smem + TILE * XLD;
    const int tid = threadIdx.x;
    for (int id = tid; id < TILE * TILE; id += NTHREADS) T[(id >> 7) * XLD + (id & 127)] = A[id];
    if (tid == 0) s_fail = 1 << 30;
    __syncthreads();
    long long t0 = clock64(), t1, t2;
    if (VARIANT == 0) {
        DiagBlock blk;
        diag_block_coords(blk);
        potf2_blk(T, Bm, invblk, &s_fail, blk);
        t1 = clock64();
        for (int id = tid; id < TILE * TILE; id += NTHREADS) Lout[id] = ((id & 127) <= (id >> 7)) ? T[(id >> 7) * XLD + (id & 127)] : 0.0;
        __syncthreads();
        long long t1b = clock64();
        trtri_blk(T, Bm, invblk, blk);
        t2 = clock64() - (t1b - t1);
    } else {
        potf2_tile(T, Bm, &s_fail);
        t1 = clock64();
        for (int id = tid; id < TILE * TILE; id += NTHREADS) Lout[id] = ((id & 127) <= (id >> 7)) ? T[(id >> 7) * XLD + (id & 127)] : 0.0;
        __syncthreads();
        long long t1b = clock64();
        trtri_tile(T, Bm);
        t2 = clock64() - (t1b - t1);
    }
    for (int id = tid; id < TILE * TILE; id += NTHREADS) Xout[id] = T[(id >> 7) * XLD + (id & 127)];
    if (tid == 0) {
        cyc[0] = t1 - t0;
        cyc[1] = t2 - t1;
        *fail_out = s_fail;
    }
}

// the two single-warp phases in isolation (4 panels back to back, no updates in between: timing only)
__global__ void __launch_bounds__(NTHREADS, 1) phase_kernel(const double* __restrict__ A, long long* __restrict__ cyc) {
    extern __shared__ double smem[];
    double* T = smem;
    double* S = smem + TILE * XLD;
    for (int id = threadIdx.x; id < TILE * TILE; id += NTHREADS) T[(id >> 7) * XLD + (id & 127)] = A[id];
    __syncthreads();
    if (threadIdx.x < 32) {
        int f = 0;
        const long long t0 = clock64();
        for (int p = 0; p < 4; ++p) f += chol32_warp(T, 32 * p, S + DT_LT, S + DT_DINV + 32 * p);
        const long long t1 = clock64();
        for (int p = 0; p < 4; ++p) inv32_warp(S + DT_LT, S + DT_DINV + 32 * p, S + DT_XD + p * DT_NB * XD_LD);
        const long long t2 = clock64();
        if (threadIdx.x == 0) {
            cyc[0] = t1 - t0;
            cyc[1] = t2 - t1;
            cyc[2] = f;
        }
    }
}

static void host_chol(const std::vector<double>& A, std::vector<double>& L, std::vector<double>& X) {
    const int n = TILE;
    L.assign(n * n, 0.0);
    X.assign(n * n, 0.0);
    for (int j = 0; j < n; ++j) {
        long double s = A[j * n + j];
        for (int k = 0; k < j; ++k) s -= (long double)L[j * n + k] * L[j * n + k];
        L[j * n + j] = (double)sqrtl(s);
        for (int i = j + 1; i < n; ++i) {
            long double t = A[i * n + j];
            for (int k = 0; k < j; ++k) t -= (long double)L[i * n + k] * L[j * n + k];
            L[i * n + j] = (double)(t / L[j * n + j]);
        }
    }
    for (int c = 0; c < n; ++c)
        for (int r = c; r < n; ++r) {
            long double s = (r == c) ? 1.0L : 0.0L;
            for (int m = c; m < r; ++m) s -= (long double)L[r * n + m] * X[m * n + c];
            X[r * n + c] = (double)(s / L[r * n + r]);
        }
}

template <int VARIANT>
static void run(const char* name, const std::vector<double>& A, const std::vector<double>& Lh, const std::vector<double>& Xh) {
    double *dA, *dL, *dX;
    long long* dc;
    int* df;
    cudaMalloc(&dA, TILE * TILE * 8); cudaMalloc(&dL, TILE * TILE * 8); cudaMalloc(&dX, TILE * TILE * 8);
    cudaMalloc(&dc, 16); cudaMalloc(&df, 4);
    cudaMemcpy(dA, A.data(), TILE * TILE * 8, cudaMemcpyHostToDevice);
    const int smem = (TILE * XLD + BM_DOUBLES) * 8;
    cudaFuncSetAttribute(tile_kernel<VARIANT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int it = 0; it < 3; ++it) tile_kernel<VARIANT><<<1, NTHREADS, smem>>>(dA, dL, dX, dc, df);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    tile_kernel<VARIANT><<<1, NTHREADS, smem>>>(dA, dL, dX, dc, df);
    cudaEventRecord(e1);
    const cudaError_t e = cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    std::vector<double> L(TILE * TILE), X(TILE * TILE);
    long long c[2];
    int fail;
    cudaMemcpy(L.data(), dL, TILE * TILE * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(X.data(), dX, TILE * TILE * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(c, dc, 16, cudaMemcpyDeviceToHost);
    cudaMemcpy(&fail, df, 4, cudaMemcpyDeviceToHost);
    double eL = 0, eX = 0, mL = 0, mX = 0, upper = 0;
    for (int i = 0; i < TILE * TILE; ++i) {
        eL = fmax(eL, fabs(L[i] - Lh[i])); mL = fmax(mL, fabs(Lh[i]));
        eX = fmax(eX, fabs(X[i] - Xh[i])); mX = fmax(mX, fabs(Xh[i]));
        if ((i & 127) > (i >> 7)) upper = fmax(upper, fabs(X[i]));
    }
    printf("{\"variant\": \"%s\", \"potf2_clocks\": %lld, \"trtri_clocks\": %lld, \"kernel_us\": %.2f, \"rel_err_L\": %.3e, \"rel_err_X\": %.3e, "
           "\"max_upper_X\": %.1e, \"fail\": %d, \"status\": \"%s\"}\n",
           name, c[0], c[1], ms * 1e3, eL / mL, eX / mX, upper, fail == (1 << 30) ? 0 : fail, cudaGetErrorString(e));
    cudaFree(dA); cudaFree(dL); cudaFree(dX); cudaFree(dc); cudaFree(df);
}

int main() {
    // SE kernel matrix of 128 random points in 8-D at the bench's conditioning (length scale 2/sqrt(d), noise 0.3)
    const int n = TILE, d = 8;
    std::vector<double> x(n * d), A(n * n), Lh, Xh;
    srand(1);
    for (auto& v : x) v = rand() / (double)RAND_MAX;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            double s = 0;
            for (int k = 0; k < d; ++k) {
                const double t = (x[i * d + k] - x[j * d + k]) * 2.0 / sqrt((double)d);
                s += t * t;
            }
            A[i * n + j] = exp(-s) + (i == j ? 0.09 + 1e-7 : 0.0);
        }
    host_chol(A, Lh, Xh);
    {
        double* dA;
        long long* dc;
        cudaMalloc(&dA, TILE * TILE * 8); cudaMalloc(&dc, 24);
        cudaMemcpy(dA, A.data(), TILE * TILE * 8, cudaMemcpyHostToDevice);
        const int smem = (TILE * XLD + BM_DOUBLES) * 8;
        cudaFuncSetAttribute(phase_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        for (int it = 0; it < 2; ++it) phase_kernel<<<1, NTHREADS, smem>>>(dA, dc);
        long long c[3];
        cudaMemcpy(c, dc, 24, cudaMemcpyDeviceToHost);
        printf("{\"phase\": \"single-warp phases, 4 panels\", \"chol32_x4_clocks\": %lld, \"inv32_x4_clocks\": %lld}\n", c[0], c[1]);
        cudaFree(dA); cudaFree(dc);
    }
    run<0>("blocks8 (old)", A, Lh, Xh);
    run<1>("warp32 + dmma (new)", A, Lh, Xh);
    // a non-positive-definite tile: the first bad pivot must be reported by both
    A[40 * n + 40] = -1.0;
    host_chol(A, Lh, Xh);
    run<0>("blocks8 (old), pivot 41 <= 0", A, Lh, Xh);
    run<1>("warp32 + dmma (new), pivot 41 <= 0", A, Lh, Xh);
    return 0;
}
