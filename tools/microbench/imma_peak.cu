// Microbenchmark: legacy warp-level tensor-core issue rates on sm_100a (B200), to size an Ozaki-style split of the FP64
// products onto integer / low-precision tensor cores WITHOUT tcgen05:
//   imma16832 : mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32   (16*8*32*2 = 8192 int ops)
//   hmma16816 : mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 (4096 flop)
//   tf32_1688 : mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32  (2048 flop)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o imma_peak imma_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int NACC>
__global__ void k_imma(int* out, int iters) {
    int acc[NACC][4];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0;
    unsigned a[4] = {threadIdx.x, threadIdx.x + 1, threadIdx.x + 2, threadIdx.x + 3}, b[2] = {threadIdx.x * 3, threadIdx.x * 5};
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i)
            asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+r"(acc[i][0]), "+r"(acc[i][1]), "+r"(acc[i][2]), "+r"(acc[i][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    int s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3];
    if (s == 123456789) out[0] = s;
}

template <int NACC>
__global__ void k_hmma(float* out, int iters) {
    float acc[NACC][4];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
    unsigned a[4] = {0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u}, b[2] = {0x3f803f80u, 0x3f803f80u};
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(acc[i][0]), "+f"(acc[i][1]), "+f"(acc[i][2]), "+f"(acc[i][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3];
    if (s == 123.456f) out[0] = s;
}

template <int NACC>
__global__ void k_tf32(float* out, int iters) {
    float acc[NACC][4];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
    unsigned a[4] = {0x3f800000u, 0x3f800000u, 0x3f800000u, 0x3f800000u}, b[2] = {0x3f800000u, 0x3f800000u};
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(acc[i][0]), "+f"(acc[i][1]), "+f"(acc[i][2]), "+f"(acc[i][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3];
    if (s == 123.456f) out[0] = s;
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
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    void* out; cudaMalloc(&out, 64);
    const int iters = 20000;
    for (int warps : {4, 8, 16}) {
        const int threads = warps * 32, blocks = sms;
        constexpr int N = 16;
        float ms = time_it([&] { k_imma<N><<<blocks, threads>>>((int*)out, iters); });
        double ops = (double)blocks * warps * iters * N * 8192.0;
        printf("imma m16n8k32 s8   warps/SM=%2d: %8.1f TOP/s\n", warps, ops / ms * 1e-9);
        ms = time_it([&] { k_hmma<N><<<blocks, threads>>>((float*)out, iters); });
        ops = (double)blocks * warps * iters * N * 4096.0;
        printf("hmma m16n8k16 bf16 warps/SM=%2d: %8.1f TFLOP/s\n", warps, ops / ms * 1e-9);
        ms = time_it([&] { k_tf32<N><<<blocks, threads>>>((float*)out, iters); });
        ops = (double)blocks * warps * iters * N * 2048.0;
        printf("mma  m16n8k8  tf32 warps/SM=%2d: %8.1f TFLOP/s\n", warps, ops / ms * 1e-9);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
