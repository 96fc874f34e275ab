// Issue-rate microbenchmark of tcgen05.mma kind::i8 (and kind::f8f6f4 for comparison) on sm_100a:
// cycles per UMMA as a function of N, of the number of accumulators cycled through, and of where A comes from
// (shared memory descriptor vs tensor memory).  Operands are zeros: only the timing matters.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_i8_rate umma_i8_rate.cu && ./umma_i8_rate
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr uint64_t DESC_HI = (1ull << 16) | ((512ull >> 4) << 32) | (1ull << 46) | (4ull << 61);   // K-major, SWIZZLE_64B
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) { return DESC_HI | (uint64_t)((addr & 0x3FFFF) >> 4); }

template <int KIND>   // 0 = i8, 1 = f8f6f4 (e4m3)
__device__ __forceinline__ void umma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
    if (KIND == 0)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc) : "memory");
}
template <int KIND>
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc) {
    if (KIND == 0)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f8f6f4 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc) : "memory");
}

template <int KIND, int TS>
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int nacc, int na, int nb, int iters, long long* cycles) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0u;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    // instruction descriptor: i8: D=S32, A=B=S8; f8f6f4: D=F32, A=B=E4M3
    const uint32_t idesc = (KIND == 0 ? ((2u << 4) | (1u << 7) | (1u << 10)) : (1u << 4)) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
    if (threadIdx.x == 0) {
        const int a_bytes = 128 * 64, b_bytes = N * 64;           // one [rows x 64 B] K-major tile per plane
        const uint32_t a0 = base, b0 = base + 8 * a_bytes;
        // descriptors precomputed: the timed loop is 12 back-to-back UMMAs per trip, nothing else
        uint64_t ad[12], bd[12];
        uint32_t dd[12];
        for (int j = 0; j < 12; ++j) {
            dd[j] = tmem + (uint32_t)((j % nacc) * N);
            bd[j] = smem_desc(b0 + (uint32_t)((j % nb) * b_bytes + ((j / nb) & 1) * 32));
            ad[j] = smem_desc(a0 + (uint32_t)((j % na) * a_bytes + ((j / na) & 1) * 32));
        }
        const uint32_t a_tm = tmem + (uint32_t)(nacc * N);
        long long t0 = clock64();
        for (int i = 0; i < iters; i += 12) {
#pragma unroll
            for (int j = 0; j < 12; ++j) {
                if (TS) umma_ts<KIND>(dd[j], a_tm, bd[j], idesc);
                else umma_ss<KIND>(dd[j], ad[j], bd[j], idesc);
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
        long long t1 = clock64();
        cycles[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

template <int KIND, int TS>
static void run(const char* name, int N, int nacc, int na, int nb, int grid) {
    const int iters = 24000;
    long long* d;
    cudaMalloc(&d, grid * sizeof(long long));
    const int smem = 201 * 1024;
    cudaFuncSetAttribute(rate_kernel<KIND, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    rate_kernel<KIND, TS><<<grid, 128, smem>>>(N, nacc, na, nb, iters, d);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    rate_kernel<KIND, TS><<<grid, 128, smem>>>(N, nacc, na, nb, iters, d);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    long long h[148] = {0};
    cudaMemcpy(h, d, (grid < 148 ? grid : 148) * sizeof(long long), cudaMemcpyDeviceToHost);
    const double cyc = (double)h[0] / iters;
    const double ops = 2.0 * 128 * N * 32 * iters * grid;
    printf("{\"case\": \"%s\", \"N\": %d, \"accumulators\": %d, \"a_tiles\": %d, \"b_tiles\": %d, \"grid\": %d, \"cycles_per_mma\": %.2f, "
           "\"ideal_cycles\": %.1f, \"chip_tops\": %.1f, \"status\": \"%s\"}\n",
           name, N, nacc, na, nb, grid, cyc, N / 2.0, ops / (ms * 1e-3) * 1e-12, cudaGetErrorString(e));
    cudaFree(d);
}

int main() {
    for (int grid : {1, 148}) {
        for (int N : {64, 128, 256}) {
            const int nacc = 512 / N > 6 ? 6 : (512 / N > 1 ? 512 / N - (N == 256 ? 0 : 1) : 1);
            run<0, 0>("i8_ss", N, nacc, 6, 6, grid);
            run<0, 0>("i8_ss_oneA", N, nacc, 1, 6, grid);
            run<0, 1>("i8_ts", N, nacc > 1 ? nacc - (nacc * N + 32 > 512 ? 1 : 0) : 1, 1, 6, grid);
            run<1, 0>("e4m3_ss", N, nacc, 6, 6, grid);
        }
        run<0, 0>("i8_ss_1acc", 64, 1, 6, 6, grid);
    }
    return 0;
}
