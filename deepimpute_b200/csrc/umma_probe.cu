// Stand-alone check of the tcgen05/TMA building blocks in tc_common.cuh on a real B200:
// D[128][64] = A[128][64] * B[64][64]^T-or-not for every operand major-ness, integer-valued inputs (exact in
// tf32), compared bit-for-bit with the CPU.  Usage: umma_probe <a_mn 0|1> <b_mn 0|1> [lboA sboA advA lboB sboB advB]
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "tc_common.cuh"

using namespace di;
using namespace di::tc;

constexpr int M = 128, N = 64, K = 64;

struct Over { int lbo, sbo, adv, type; };

template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                                             float* D, Over oa, Over ob) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;                       // 2 stages x 16 KB
    uint8_t* sB = smem + 2 * 16384;           // 2 stages x 8 KB
    __shared__ uint64_t full_bar, mma_bar;
    __shared__ uint32_t tmem_base;
    const int warp = threadIdx.x / 32;
    if (threadIdx.x == 0) { mbar_init(&full_bar, 1); mbar_init(&mma_bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(&tmem_base, 64);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = tmem_base;
    if (warp == 0 && elect_one()) {
        mbar_arrive_expect_tx(&full_bar, 2 * (16384 + 8192));
        for (int kb = 0; kb < 2; ++kb) {
            load_stage<A_MN>(sA + kb * 16384, &mapA, &full_bar, 0, kb * BLOCK_K, M);
            load_stage<B_MN>(sB + kb * 8192, &mapB, &full_bar, 0, kb * BLOCK_K, N);
        }
        mbar_wait(&full_bar, 0);
        tc_fence_after();
        const uint32_t idesc = make_idesc_tf32(M, N, A_MN, B_MN);
        for (int kb = 0; kb < 2; ++kb)
            for (int j = 0; j < 4; ++j) {
                const uint64_t da = make_smem_desc(smem_u32(sA + kb * 16384) + j * oa.adv, oa.lbo, oa.sbo, oa.type);
                const uint64_t db = make_smem_desc(smem_u32(sB + kb * 8192) + j * ob.adv, ob.lbo, ob.sbo, ob.type);
                umma_tf32(tm, da, db, idesc, (kb | j) ? 1u : 0u);
            }
        umma_commit(&mma_bar);
    }
    __syncwarp();
    mbar_wait(&mma_bar, 0);
    tc_fence_after();
    const int row = warp * 32 + (threadIdx.x & 31);
    for (int c = 0; c < N; c += 16) {
        float v[16];
        tmem_ld16(tm + ((uint32_t)(warp * 32) << 16) + c, v);
        for (int i = 0; i < 16; ++i) D[row * N + c + i] = v[i];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 64);
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)

template <bool A_MN, bool B_MN>
int run(Over oa, Over ob) {
    std::vector<float> A(M * K), B(N * K), Dref(M * N, 0.f), D(M * N, -1.f);
    srand(7);
    for (auto& x : A) x = (float)(rand() % 7 - 3);
    for (auto& x : B) x = (float)(rand() % 7 - 3);
    // logical A[m][k], B[n][k]
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { float a = 0; for (int k = 0; k < K; ++k) a += A[m * K + k] * B[n * K + k]; Dref[m * N + n] = a; }
    std::vector<float> Ag(M * K), Bg(N * K);
    for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) { if (A_MN) Ag[k * M + m] = A[m * K + k]; else Ag[m * K + k] = A[m * K + k]; }
    for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) { if (B_MN) Bg[k * N + n] = B[n * K + k]; else Bg[n * K + k] = B[n * K + k]; }
    float *dA, *dB, *dD;
    CK(cudaMalloc(&dA, Ag.size() * 4)); CK(cudaMalloc(&dB, Bg.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
    CK(cudaMemcpy(dA, Ag.data(), Ag.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, Bg.data(), Bg.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xFF, D.size() * 4));
    CUtensorMap mA, mB;
    bool ok = A_MN ? make_map_2d(&mA, dA, K, M, M, 32, true) : make_map_2d(&mA, dA, M, K, K, M);
    ok = ok && (B_MN ? make_map_2d(&mB, dB, K, N, N, 32, true) : make_map_2d(&mB, dB, N, K, K, N));
    if (!ok) { printf("tensor map encode failed\n"); return 2; }
    const int smem = 2 * 16384 + 2 * 8192 + 1024;
    CK(cudaFuncSetAttribute(probe<A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    probe<A_MN, B_MN><<<1, 128, smem>>>(mA, mB, dD, oa, ob);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    int bad = 0; double maxerr = 0;
    for (int i = 0; i < M * N; ++i) { double e = fabs((double)D[i] - Dref[i]); if (!(e == 0)) ++bad; if (e > maxerr) maxerr = e; }
    printf("a_mn=%d b_mn=%d A(lbo=%d sbo=%d adv=%d type=%d) B(lbo=%d sbo=%d adv=%d type=%d): mismatches=%d/%d maxerr=%g  D[0..3]=%g %g %g %g ref=%g %g %g %g\n",
           (int)A_MN, (int)B_MN, oa.lbo, oa.sbo, oa.adv, oa.type, ob.lbo, ob.sbo, ob.adv, ob.type, bad, M * N, maxerr,
           D[0], D[1], D[2], D[3], Dref[0], Dref[1], Dref[2], Dref[3]);
    return bad ? 1 : 0;
}

int main(int argc, char** argv) {
    const int a_mn = argc > 1 ? atoi(argv[1]) : 0, b_mn = argc > 2 ? atoi(argv[2]) : 0;
    Over oa = a_mn ? Over{(int)MN_BOX_BYTES, 512, 1024, (int)SWIZZLE_128B_BASE32B} : Over{16, 1024, 32, (int)SWIZZLE_128B};
    Over ob = b_mn ? Over{(int)MN_BOX_BYTES, 512, 1024, (int)SWIZZLE_128B_BASE32B} : Over{16, 1024, 32, (int)SWIZZLE_128B};
    if (argc > 8) { oa.lbo = atoi(argv[3]); oa.sbo = atoi(argv[4]); oa.adv = atoi(argv[5]); ob.lbo = atoi(argv[6]); ob.sbo = atoi(argv[7]); ob.adv = atoi(argv[8]); }
    if (!a_mn && !b_mn) return run<false, false>(oa, ob);
    if (a_mn && !b_mn) return run<true, false>(oa, ob);
    if (!a_mn && b_mn) return run<false, true>(oa, ob);
    return run<true, true>(oa, ob);
}
