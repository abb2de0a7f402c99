// sm_100a building blocks used by kernels_tc.cu: mbarrier, TMA (cp.async.bulk.tensor), TMEM allocation,
// tcgen05.mma (kind::tf32) with shared-memory matrix descriptors, tcgen05.ld.  Inline PTX only.
//
// Shared-memory operand layouts (fp32/tf32 elements, tiles written by TMA, verified bit-exact by umma_probe.cu):
//   K-major  operand (K contiguous in memory):  one TMA box = [rows = MN extent][32 elements = 128 B], 128-byte
//            swizzle (16-byte chunks); 8-row groups are 1024 B apart (SBO); one MMA (K = 8) reads 32 B of every
//            row, so successive MMAs inside the 128-byte atom advance the descriptor start address by 32 B.
//   MN-major operand (M or N contiguous in memory): one TMA box = [rows = K extent][32 elements along MN].
//            For 32-bit operands the tensor core accepts only the "128-byte swizzle with 32-byte atomicity"
//            layout here (TMA: CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, descriptor layout type 1): the atom is
//            32 MN elements x 4 K rows (512 B), 32-byte chunks XOR-ed with (row & 3).  Groups of 4 K rows are
//            512 B apart (SBO), successive 32-element MN chunks one box apart (LBO = box bytes); one MMA (K = 8)
//            reads two atoms = 1024 B, so successive MMAs advance the start address by 1024 B.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace di {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n"
        ".reg .b32 rx;\n"
        ".reg .pred px;\n"
        "elect.sync rx|px, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, px;\n"
        "}\n" : "=r"(pred));
    return pred != 0;
}

// ---------------------------------------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must neither hang the GPU nor kill the context.  After ~0.5 s the waiter records
// who gave up (g_mbar_timeout: tag << 24 | blockIdx.z << 12 | blockIdx.y << 4 | warp) and carries on, so the kernel
// terminates with garbage results; the host reads the word at its next synchronisation point and fails the call
// with that information (di_api.cu: sync_check).
__device__ unsigned int g_mbar_timeout = 0;
__device__ __noinline__ void mbar_wait_slow(uint64_t* bar, uint32_t parity, uint32_t tag) {
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 1000000000ll) {
            atomicCAS(&g_mbar_timeout, 0u, (tag << 24) | ((blockIdx.z & 0xFFFu) << 12) | ((blockIdx.y & 0xFFu) << 4) |
                                               ((threadIdx.x >> 5) & 0xFu) | 0x80000000u);
            return;
        }
    }
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, uint32_t tag = 0) {
    if (mbar_try_wait(bar, parity)) return;
    mbar_wait_slow(bar, parity, tag);
}

// --------------------------------------------------------------------------------------------------- TMA
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
// 2-D tiled load: c0 = element offset along the contiguous dimension, c1 = row
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int32_t c0, int32_t c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

// 2-D tiled store shared -> global (bulk async group); out-of-bounds parts of the box are not written
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int32_t c0, int32_t c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until at most N of this thread's bulk groups are still READING their shared-memory source
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
// wait until at most N of this thread's bulk groups are incomplete (writes performed)
template <int N>
__device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// -------------------------------------------------------------------------------------------------- TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {   // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {     // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ------------------------------------------------------------------------------------------- descriptors
constexpr uint32_t SWIZZLE_128B = 2;            // 128-byte swizzle, 16-byte atomicity (K-major operands)
constexpr uint32_t SWIZZLE_128B_BASE32B = 1;    // 128-byte swizzle, 32-byte atomicity (MN-major 32-bit operands)

// Shared-memory matrix descriptor (sm_100 format: 14-bit address/LBO/SBO fields in 16-byte units, version 1
// at bit 46, swizzle mode at bits 61..63).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout_type = SWIZZLE_128B) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout_type << 61;
    return d;
}

// Instruction descriptor, kind::tf32, fp32 accumulate: c_format F32 (1) @4, a/b_format TF32 (2) @7/@10,
// a_major @15, b_major @16 (1 = MN-major), N >> 3 @17, M >> 4 @24.
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N, bool a_mn, bool b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]; issued by one thread
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// completion of all prior MMAs of this thread arrives on the mbarrier (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- operand staging helpers: one pipeline stage holds a [MN extent] x [32 K] fp32 slab (BLOCK_K = 32) ----
constexpr int BLOCK_K = 32;                 // fp32 elements per stage along K = one 128-byte swizzle row
constexpr int UMMA_K = 8;                   // K per tcgen05.mma kind::tf32
constexpr uint32_t MN_BOX_BYTES = 32 * BLOCK_K * 4;   // MN-major box: 32 K-rows x 128 B

// issue the TMA loads of one stage; mn0/k0 are element coordinates in the global tensor
template <bool MN_MAJOR>
__device__ __forceinline__ void load_stage(uint8_t* smem, const CUtensorMap* map, uint64_t* bar, int mn0, int k0, int mn_extent) {
    if constexpr (MN_MAJOR) {
        for (int c = 0; c < mn_extent / 32; ++c) tma_load_2d(smem + c * MN_BOX_BYTES, map, bar, mn0 + 32 * c, k0);
    } else {
        tma_load_2d(smem, map, bar, k0, mn0);       // box rows == mn_extent
    }
}
// descriptor of the kstep-th (0..3) K=8 slice of a stage
template <bool MN_MAJOR>
__device__ __forceinline__ uint64_t stage_desc(uint32_t saddr, int kstep) {
    if constexpr (MN_MAJOR) return make_smem_desc(saddr + kstep * 1024, MN_BOX_BYTES, 512, SWIZZLE_128B_BASE32B);
    else return make_smem_desc(saddr + kstep * UMMA_K * 4, 16, 1024, SWIZZLE_128B);
}

// 16 consecutive fp32 accumulator columns of this thread's TMEM lane (warp w reads lanes 32*(w%4) .. +31)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float v[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// 8 consecutive fp32 accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float v[8]) {
    uint32_t r[8];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

}  // namespace tc

// host: 0 if no mbarrier wait has timed out since the last call, else the recorded word (and clears it)
unsigned int tc_take_timeout_word();

// ---------------------------------------------------------------------------------- host: tensor maps
// cuTensorMapEncodeTiled is fetched through the runtime so that nothing links against libcuda.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// fp32 row-major [rows][cols] with row pitch `pitch` floats; box = 32 floats (128 B, the swizzle span) x box_rows.
// Out-of-bounds parts of a box are filled with zeros, which is what lets ragged K / M / N tails work.
// mn_major selects the swizzle an MN-major 32-bit operand needs (32-byte atomicity), see the header comment.
inline bool make_map_2d(CUtensorMap* map, const float* base, uint64_t rows, uint64_t cols, uint64_t pitch, uint32_t box_rows,
                        bool mn_major = false) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {pitch * sizeof(float)};
    cuuint32_t box[2] = {32, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE,
                    mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// fp32 row-major [rows][cols], un-swizzled box of box_cols x box_rows elements: plain tiles that threads index
// directly in shared memory (weight / moment tiles of the Adam epilogue, target tiles of the loss epilogue).
inline bool make_map_plain(CUtensorMap* map, const float* base, uint64_t rows, uint64_t cols, uint64_t pitch,
                           uint32_t box_cols, uint32_t box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {pitch * sizeof(float)};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

}  // namespace di
