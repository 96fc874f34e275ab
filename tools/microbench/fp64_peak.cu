// Microbenchmark: raw FP64 issue rates on sm_100a (B200).
//   dmma884  : mma.sync.aligned.m8n8k4.row.col.f64 (SASS DMMA.8x8x4), register-resident operands
//   dmma16816: mma.sync.aligned.m16n8k16.row.col.f64
//   dfma     : plain fma.rn.f64 chains
// Prints TFLOP/s for several warps-per-SM settings. Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int NACC>
__global__ void k_dmma884(double* out, int iters, double a0, double b0) {
    double acc[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { acc[i][0] = 0.0; acc[i][1] = 0.0; }
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(acc[i][0]), "+d"(acc[i][1]) : "d"(a), "d"(b));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i][0] + acc[i][1];
    if (s == 123.456) out[0] = s;
}

template <int NACC>
__global__ void k_dmma16816(double* out, int iters, double a0, double b0) {
    double acc[NACC][4];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.0; }
    double a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = a0 + threadIdx.x * 1e-9 + i;
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = b0 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                         : "+d"(acc[i][0]), "+d"(acc[i][1]), "+d"(acc[i][2]), "+d"(acc[i][3])
                         : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                           "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3];
    if (s == 123.456) out[0] = s;
}

template <int NACC>
__global__ void k_dfma(double* out, int iters, double a0, double b0) {
    double acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = i;
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i];
    if (s == 123.456) out[0] = s;
}

template <typename F>
float time_it(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    printf("device %s sms %d clock %d kHz\n", prop.name, sms, prop.clockRate);
    double* out; CK(cudaMalloc(&out, 8));
    const int iters = 4096;
    int wps_list[] = {4, 8, 16, 32};
    for (int wi = 0; wi < 4; ++wi) {
        int wps = wps_list[wi];
        int threads = 32 * wps > 1024 ? 1024 : 32 * wps;
        int blocks_per_sm = (32 * wps) / threads;
        int grid = sms * blocks_per_sm;
        {
            float ms = time_it([&] { k_dmma884<16><<<grid, threads>>>(out, iters, 1.0, 1e-3); });
            double flops = (double)grid * (threads / 32) * iters * 16 * 512.0;
            printf("dmma884   warps/SM %2d : %8.3f ms  %7.2f TFLOP/s\n", wps, ms, flops / ms * 1e-9);
        }
        {
            float ms = time_it([&] { k_dmma16816<4><<<grid, threads>>>(out, iters, 1.0, 1e-3); });
            double flops = (double)grid * (threads / 32) * iters * 4 * (16.0 * 8 * 16 * 2);
            printf("dmma16816 warps/SM %2d : %8.3f ms  %7.2f TFLOP/s\n", wps, ms, flops / ms * 1e-9);
        }
        {
            float ms = time_it([&] { k_dfma<16><<<grid, threads>>>(out, iters, 1.0, 1e-3); });
            double flops = (double)grid * threads * iters * 16 * 2.0;
            printf("dfma      warps/SM %2d : %8.3f ms  %7.2f TFLOP/s\n", wps, ms, flops / ms * 1e-9);
        }
    }
    CK(cudaGetLastError());
    return 0;
}
