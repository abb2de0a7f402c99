// Stand-alone check of the building block the round-2 main loop needs: tcgen05.mma
// kind::tf32 with the A operand in TENSOR MEMORY (written from registers with tcgen05.st) and B in shared memory.
//   D[128][64] = A[128][64] * B[64][64]^T, integer-valued inputs (exact in tf32), compared bit-for-bit with the CPU.
// Hypothesis under test: A lives in TMEM as lane = row m, one 32-bit column per k; the MMA of K step j reads the 8
// columns [a_col0 + 8 j, a_col0 + 8 j + 8).  Usage: umma_probe_ts [a_major_bit 0|1] [col_stride_per_kstep (default 8)]
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "tc_common.cuh"

using namespace di;
using namespace di::tc;

constexpr int M = 128, N = 64, K = 64;
constexpr uint32_t A_COL0 = 64, TMEM_COLS = 128;   // accumulator in columns 0..63, A in columns 64..127

// D[tmem] (+)= A[tmem] * B[smem]; issued by one thread
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// 32 consecutive 32-bit columns of this thread's TMEM lane (warp w writes lanes 32*(w%4) .. +31)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float v[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
        "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        ::"r"(taddr),
          "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
          "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
          "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
          "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
          "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
          "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
          "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
          "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
        : "memory");
}

__global__ void __launch_bounds__(128) probe_ts(const float* __restrict__ A, const __grid_constant__ CUtensorMap mapB, float* D,
                                                int a_major_bit, int col_stride) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sB = smem;                       // 2 stages x 8 KB, K-major, 128-byte swizzle
    __shared__ uint64_t full_bar, mma_bar;
    __shared__ uint32_t tmem_base;
    const int warp = threadIdx.x / 32, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(&full_bar, 1); mbar_init(&mma_bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(&tmem_base, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = tmem_base;
    const int row = warp * 32 + lane;

    // every thread puts its row of A (64 values) into its TMEM lane, columns A_COL0 .. A_COL0 + 63
    for (int k0 = 0; k0 < K; k0 += 32) {
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = A[row * K + k0 + i];
        tmem_st32(tm + ((uint32_t)(warp * 32) << 16) + A_COL0 + k0, v);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp == 0 && elect_one()) {
        mbar_arrive_expect_tx(&full_bar, 2 * 8192);
        for (int kb = 0; kb < 2; ++kb) load_stage<false>(sB + kb * 8192, &mapB, &full_bar, 0, kb * BLOCK_K, N);
        mbar_wait(&full_bar, 0);
        tc_fence_after();
        const uint32_t idesc = make_idesc_tf32(M, N, a_major_bit != 0, false);
        for (int kb = 0; kb < 2; ++kb)
            for (int j = 0; j < 4; ++j)
                umma_tf32_ts(tm, tm + A_COL0 + (uint32_t)((kb * 4 + j) * col_stride),
                             stage_desc<false>(smem_u32(sB + kb * 8192), j), idesc, (kb | j) ? 1u : 0u);
        umma_commit(&mma_bar);
    }
    __syncwarp();
    mbar_wait(&mma_bar, 0);
    tc_fence_after();
    for (int c = 0; c < N; c += 16) {
        float v[16];
        tmem_ld16(tm + ((uint32_t)(warp * 32) << 16) + c, v);
        for (int i = 0; i < 16; ++i) D[row * N + c + i] = v[i];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, TMEM_COLS);
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)

int main(int argc, char** argv) {
    const int a_major_bit = argc > 1 ? atoi(argv[1]) : 0, col_stride = argc > 2 ? atoi(argv[2]) : 8;
    std::vector<float> A(M * K), B(N * K), Dref(M * N, 0.f), D(M * N, -1.f);
    srand(7);
    for (auto& x : A) x = (float)(rand() % 7 - 3);
    for (auto& x : B) x = (float)(rand() % 7 - 3);
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { float a = 0; for (int k = 0; k < K; ++k) a += A[m * K + k] * B[n * K + k]; Dref[m * N + n] = a; }
    float *dA, *dB, *dD;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xFF, D.size() * 4));
    CUtensorMap mB;
    if (!make_map_2d(&mB, dB, N, K, K, N)) { printf("tensor map encode failed\n"); return 2; }
    const int smem = 2 * 8192 + 1024;
    CK(cudaFuncSetAttribute(probe_ts, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    probe_ts<<<1, 128, smem>>>(dA, mB, dD, a_major_bit, col_stride);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    int bad = 0; double maxerr = 0;
    for (int i = 0; i < M * N; ++i) { double e = fabs((double)D[i] - Dref[i]); if (!(e == 0)) ++bad; if (e > maxerr) maxerr = e; }
    printf("A in TMEM (a_major bit %d, %d columns per K step): mismatches=%d/%d maxerr=%g  D[0..3]=%g %g %g %g ref=%g %g %g %g | "
           "D[64*64..]=%g %g ref=%g %g\n", a_major_bit, col_stride, bad, M * N, maxerr, D[0], D[1], D[2], D[3],
           Dref[0], Dref[1], Dref[2], Dref[3], D[64 * N], D[64 * N + 1], Dref[64 * N], Dref[64 * N + 1]);
    return bad ? 1 : 0;
}
