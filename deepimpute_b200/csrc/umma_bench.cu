// Micro-benchmark: cycles per tcgen05.mma kind::tf32 (M = 128, K = 8, SS mode) as a function of N and of the
// operand major-ness, one CTA per SM.  Operands are whatever is in shared memory (timing only).
// usage: umma_bench
#include <cstdio>
#include <cstdlib>
#include "tc_common.cuh"
using namespace di;
using namespace di::tc;

template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(128) bench(int N, int reps, unsigned long long* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;                       // 16 KB: one 32-wide K slab of A
    uint8_t* sB = smem + 16384;               // N * 128 B
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    for (int i = threadIdx.x; i < (16384 + N * 128) / 4; i += blockDim.x) ((float*)smem)[i] = 1.0f;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc(&tmem_base, 256);
    fence_proxy_async();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = tmem_base;
    if (threadIdx.x == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                               ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r)
#pragma unroll
            for (int j = 0; j < 4; ++j)
                umma_tf32(tm, stage_desc<A_MN>(smem_u32(sA), j), stage_desc<B_MN>(smem_u32(sB), j), idesc, 1u);
        const long long t1 = clock64();
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        const long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tm, 256);
}

template <bool A_MN, bool B_MN>
void run(int N, int grid) {
    unsigned long long* d; cudaMalloc(&d, 16); unsigned long long h[2] = {0, 0};
    const int reps = 256, smem = 16384 + N * 128 + 1024;
    cudaFuncSetAttribute(bench<A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    bench<A_MN, B_MN><<<grid, 128, smem>>>(N, reps, d);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("A_%s B_%s N=%3d grid=%3d: issue %.1f cyc/MMA, complete %.1f cyc/MMA  (%s)\n", A_MN ? "MN" : "K ", B_MN ? "MN" : "K ", N, grid,
           (double)h[0] / (reps * 4), (double)h[1] / (reps * 4), cudaGetErrorString(e));
    cudaFree(d);
}

int main() {
    for (int grid : {1, 148})
        for (int N : {32, 64, 128, 256}) {
            run<false, false>(N, grid);
            run<true, false>(N, grid);
            run<true, true>(N, grid);
        }
    return 0;
}
