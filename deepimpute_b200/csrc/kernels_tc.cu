// tcgen05 / TMA path (DI_MATH_TF32): every GEMM of the training step and of inference runs on the 5th-generation
// tensor cores (kind::tf32: fp32 operands read straight from HBM by TMA, truncated to TF32 by the tensor core, fp32
// accumulation in TMEM), batched over sub-networks with blockIdx.z, with the layer-specific work in the epilogue:
//
//   FWD1  h    = dropout(relu(X W1 + b1))                                   A = W1 (MN-major)  B = X    (K-major)
//   FWD2  yhat = softplus(h W2 + b2); wMSE; dz2; b2 <- Adam(sum_b dz2)       A = W2 (MN-major)  B = h    (K-major)
//   BWD   dz1  = (dz2 W2^T) * relu/dropout mask; b1 <- Adam(sum_b dz1)       A = W2 (K-major)   B = dz2  (K-major)
//   ADAM  W   <- Adam(in^T dout), dW never leaves TMEM/registers             A = dout (MN-major) B = in (MN-major)
//         (ADAM2: in = h, dout = dz2, W = W2;  ADAM1: in = X, dout = dz1, W = W1)
//
// Orientation: the M dimension of every MMA (the 128 TMEM lanes) is the layer's OUTPUT feature index -- the
// contiguous dimension of the Keras-layout weights W[in][out] and of every activation row.  One epilogue thread
// therefore owns one feature: its bias is a scalar, the bias gradient is a private sum over the batch columns, one
// Philox call yields the keep-bits of 4 consecutive batch rows, and every global access of a warp is 32 consecutive
// floats (128 B, coalesced) with no shared-memory transpose.  The batch (or the cell tile at inference) is the MMA
// N dimension, so B = 64 costs N = 64 rather than a half-empty M.
//
// The ADAM kernel is the HBM-bound one (24 B per parameter per step: w, m, v read and written).  Its epilogue never
// touches global memory from registers: a ring of TMA bulk loads streams [8 rows x 128 features] tiles of w, m, v into
// shared memory ahead of use, the epilogue threads update them in place against the accumulator columns read from
// TMEM, and TMA bulk stores drain them -- enough bytes in flight per SM to cover HBM latency without registers.
//
// Warp roles (192 threads): warps 0-3 epilogue (warp w reads TMEM lanes 32w..32w+31), warp 4 TMA producer,
// warp 5 MMA issuer.  smem ring of 32-wide K slabs, full/empty mbarriers, one tmem_full barrier.  X3 kernels have
// four more warps (6-9): all eight non-producer warps compute the residual slabs during the main loop and then split
// the epilogue (two warps per TMEM lane quadrant, half of the batch columns each).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <utility>

#include "engine.h"
#include "tc_common.cuh"

namespace di {

using namespace tc;

namespace {

enum { TC_FWD1 = 0, TC_FWD2 = 1, TC_BWD = 2 };

constexpr int TILE_M = 128;              // TMEM lanes = output features per CTA
constexpr int NTHREADS = 192;        // warps 0-3 epilogue, 4 TMA, 5 MMA
constexpr int NTHREADS_X3 = 320;     // ... plus warps 6-9: extra residual converters (8 converter warps with 0-3)
constexpr int NCONV = 256;           // converter threads of an X3 CTA
constexpr int MAX_STAGES = 6;
constexpr uint32_t A_STAGE_BYTES = TILE_M * BLOCK_K * 4;       // 16 KB

// cells per CTA at inference (UMMA N): Engine::infer_tile, 128 or 256
constexpr int ADAM_TILE = 128;       // input features per CTA in the weight-gradient kernel (UMMA N)
constexpr int AD_R = 8;              // weight rows per streamed chunk
constexpr int AD_STAGES = 3;         // chunks of w/m/v in the dedicated part of the shared-memory ring
constexpr int AD_MAX_RING = 12;      // ... plus the stages that reuse the operand buffers once the MMAs are done

struct TcParams {
    const SubnetDesc* desc;
    int S, H, O, Hp, Op;
    int n_cols;                 // UMMA N: batch rows (FWD*/BWD) or input-feature tile width (ADAM)
    int tmem_cols;              // power of two >= (nacc + lo_acc) * n_cols
    int row_tiles;              // persistent inference kernel: 128-cell tiles per sub-network and feature tile
    int conv_teams;             // converter-warp kernels: 1 = all eight converter warps share every slab, 2 = two teams of four
                                // warps take alternate slabs (the conversion of one slab is a latency chain, not a throughput limit)
    int nacc;                   // X3: accumulators the a b products of successive K blocks rotate through (>= 1) ...
    int lo_acc;                 // ... and 1 if the small products a_lo b + a b_lo have an accumulator of their own (see acc_sum)
    int stages;                 // ring of raw (hi) slabs written by TMA
    int lo_stages;              // X3: ring of residual (lo) slabs written by the converter warps
    int which;                  // ADAM: 1 = W1 (in = X, dout = dz1), 2 = W2 (in = h, dout = dz2)
    int nkb_adam;               // ADAM: K blocks = padded batch rows / 32
    int ad_kg;                  // ADAM ring kernel: K blocks resident per pass (operand buffers hold ad_kg of them)
    int64_t row0;               // first row of the batch / cell tile group inside the B-operand tensor
    int64_t rows_per_block_y;   // inference: blockIdx.y / m_tiles selects a cell tile of n_cols rows
    int m_tiles;                // feature tiles per sub-network
    int aux_cols;               // > 0: the epilogue's side operand (Y tile / h tile) is staged by TMA, row pitch in floats
    int64_t aux_row0;
    int wbox;                   // ADAM: features per weight-tile row in shared memory (min(128, out_dim)), W1 tiles
    int wbox2;                  // ... W2 tiles
    int nx1;                    // ADAM: blockIdx.x < nx1 -> W1 tiles (in = X, dout = dz1), else W2 tiles (in = h, dout = dz2)
    // epilogue operands
    const float* Y; int64_t ldy;                 // packed targets
    float* Hact; int64_t ldh;                    // [rows][S*Hp]
    float* DZ2; float* DZ1;                      // [Bp][S*Op], [Bp][S*Hp]
    float *Hlo, *DZ2lo, *DZ1lo;                  // TF32 residual twins of h / dz2 / dz1 (nullptr: not wanted)
    float *b1, *mb1, *vb1, *b2, *mb2, *vb2;
    float *W1, *mW1, *vW1, *W2, *mW2, *vW2;      // ADAM, direct mode: updated values go to global memory from registers
    float *W1lo, *W2lo;                          // ADAM: residual twins W - trunc_tf32(W), rewritten with every update (nullptr: not kept)
    float* kpart;                                // LT split K: partial accumulators [S][m_tiles][KS][n_cols][128]
    unsigned int* kcount;                        // ... and arrival counters [S][m_tiles] (zero between launches)
    int adam_direct;                             // 1: registers -> st.global; 0: in place in the ring + TMA stores
    int ad_nded, ad_stride;                      // one-CTA-per-SM ADAM kernel: dedicated chunk stages, bytes per stage
    int ts_wbox, ts_acol0;                       // TS kernels: features per row of the plain weight tile; first TMEM column of the weight slabs
    int ad_generic;                              // experiment: run-time pitches in the ADAM chunk update
    int pdl_early;                               // release the dependent grid right after this one's own wait (experiment)
    int pdl_prefetch;                            // fetch what the previous grid does not write before waiting for it
    int pdl_lead;                                // > 0: release the dependent grid this many K blocks before the main loop ends
    float* out; int64_t ld_out;                  // inference output
    double* loss;
    int n_valid;                                 // real rows of the batch / chunk
    int training;
    uint32_t step; uint64_t seed; uint32_t drop_thresh; float keep_scale;
    float inv_norm;
    AdamParams adam;
    // sub-network group of this launch: blockIdx.z counts from s_base
    int s_base;
    // epoch-graph mode: the launch is a node of a graph that is replayed every epoch, so what changes from epoch to
    // epoch is read from device memory: the dropout counter is *step_base + step (step = position in the epoch)
    // and Adam's bias-corrected rate is lr_table[step]
    const uint32_t* step_base;
    const float* lr_table;
    // debugging (DEEPIMPUTE_B200_TRACE=1): CTA (0,0,0) records clock64() at pipeline events, 256 slots per kernel
    unsigned long long* trace;
};

// DI_TRACE: for code run by ONE elected thread.  DI_TRACE_T0: for code run by whole warps -- thread 0 records and the
// warp re-converges explicitly, because the .sync.aligned tcgen05 instructions that follow need all 32 lanes together
// (a bare `if (threadIdx.x == 0)` in front of them left warp 0 diverged: launch failures and hung mbarriers).
#define DI_TRACE(slot) do { if (p.trace && blockIdx.y == 0 && blockIdx.z == 0 && blockIdx.x == 0) p.trace[(slot)] = clock64(); } while (0)
#define DI_TRACE_T0(slot) do { if (p.trace && threadIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && blockIdx.x == 0) p.trace[(slot)] = clock64(); __syncwarp(); } while (0)

// Programmatic dependent launch: the kernels of an optimiser step form a chain on one stream.  Every kernel waits here
// (after its prologue: barriers, TMEM allocation) for the previous grid to complete and flush, and releases its own
// dependents once its accumulator is complete, so that the next kernel's launch latency and prologue overlap this
// kernel's epilogue.  Without the launch attribute both instructions are no-ops.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_release() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ uint32_t dropout_step(const TcParams& p) { return p.step_base ? *p.step_base + p.step : p.step; }
__device__ __forceinline__ AdamParams adam_of(const TcParams& p) {
    AdamParams a = p.adam;
    if (p.lr_table) a.lr_t = p.lr_table[p.step];
    return a;
}

__device__ __forceinline__ uint32_t idesc_for(int n_cols, bool a_mn, bool b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
           ((uint32_t)(n_cols >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
}

// residual of the tensor core's operand truncation: a - (a with the low 13 mantissa bits cleared)
__device__ __forceinline__ float tf32_residual(float a) {
    return a - __uint_as_float(__float_as_uint(a) & 0xFFFFE000u);
}

// Short accumulation chains.  The tensor core adds every product into its fp32 accumulator with TRUNCATION: a chain of n
// MMAs into one accumulator is biased by about n/2 ulp of the accumulator (measured: 17 K blocks x 12 MMAs = 204 MMAs
// leave FWD1 1.2e-5 from the fp32 sum, and ReLU gates + Adam's m / sqrt(v) amplify that into a tail of 1e-3 deviations
// after a few hundred steps; the same products summed by four CTAs of 51 MMAs each stay within 1e-5 of the fp32 path,
// profiles/r02e_accumulation.md).  So the compensated kernels never build one long chain: the a b products of K block kb
// go to accumulator kb % nacc, the small products a_lo b + a b_lo -- two thirds of all MMAs, 2^-11 of the magnitude --
// to an accumulator of their own where their truncation is invisible, and the epilogue adds the nacc + 1 tiles in fp32.
// taddr: lane base of the first accumulator; accumulator a sits n_cols columns further per step, the small-product
// accumulator after the nacc-th; `used` = min(nacc, K blocks issued) accumulators hold data
__device__ __forceinline__ void acc_sum16(uint32_t taddr, int c, int n_cols, int nacc, int used, bool lo_used, float (&v)[16]) {
    tmem_ld16(taddr + c, v);
    for (int a = 1; a < used; ++a) {
        float t[16];
        tmem_ld16(taddr + a * n_cols + c, t);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] += t[i];
    }
    if (lo_used) {
        float t[16];
        tmem_ld16(taddr + nacc * n_cols + c, t);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] += t[i];
    }
}

// softplus and sigmoid of z from one exponential (output layer, multinet.py:145, and its derivative)
__device__ __forceinline__ void softplus_sigmoid(float z, float& sp, float& sg) {
    const float e = __expf(-fabsf(z));
    const float l = (e < 1e-4f) ? e * (1.0f - 0.5f * e) : __logf(1.0f + e);
    sp = fmaxf(z, 0.f) + l;
    const float r = __fdividef(1.0f, 1.0f + e);
    sg = (z >= 0.f) ? r : e * r;
}

// ---- A operand in tensor memory (verified bit-exact by csrc/umma_probe_ts.cu: lane = row m, one 32-bit column per k,
// K step j of a slab reads columns [col0 + 8 j, col0 + 8 j + 8), a_major bit 0) --------------------------------------
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// 16 consecutive 32-bit columns of this thread's TMEM lane (warp w writes lanes 32*(w%4) .. +31)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        ::"r"(taddr),
          "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
          "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
          "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
          "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
        : "memory");
}
constexpr int TS_COLS = 2 * BLOCK_K;     // TMEM columns of one weight slab: 32 of W (the tensor core truncates) + 32 of W_lo

// ============================================================================================ FWD1 / FWD2 / BWD
// X3 = error-compensated "3xTF32" products for the forward GEMMs: the tensor core truncates an fp32 operand a to
// its upper 19 bits (a_hi); the epilogue warps, idle during the main loop, write the residual a_lo = a - a_hi of
// every staged slab next to it (element-wise, so the swizzled layout carries over unchanged) and the MMA warp issues
// a_lo*b_hi + a_hi*b_lo + a_hi*b_hi into the same accumulator.  The dropped a_lo*b_lo term is 2^-20 relative, i.e.
// fp32-level products on the tensor cores, at 3x the (negligible) MMA time and 2x the staging memory.
// TS (X3 FWD1 / FWD2 only): the weight slab never becomes an MMA operand in shared memory.  TMA delivers it as a plain
// [32 k][128 features] tile, the converter warps move it into tensor memory (W and W_lo side by side, tcgen05.st) and
// the MMAs take A from there: per K block the tensor core then reads 24 KB of shared memory (the activation slabs)
// instead of 72 KB, and the converters write 8 KB instead of 24 KB.
template <int OP, bool X3, bool TS = false>
__global__ void __launch_bounds__(X3 ? NTHREADS_X3 : NTHREADS, X3 ? 1 : 2) tc_kernel(const __grid_constant__ CUtensorMap mapA,
                                                                  const __grid_constant__ CUtensorMap mapB,
                                                                  const __grid_constant__ CUtensorMap mapC, const TcParams p) {
    constexpr bool A_MN = (OP != TC_BWD);

    const int s = blockIdx.z + p.s_base;
    const SubnetDesc d = p.desc[s];
    const int m_tile = (int)(blockIdx.y % p.m_tiles);
    const int row_tile = (int)(blockIdx.y / p.m_tiles);
    const int m0 = m_tile * TILE_M;                       // first output feature of this CTA
    const int64_t row0 = p.row0 + (int64_t)row_tile * p.rows_per_block_y;

    int out_dim, nkb;
    int a_c0, a_c1, b_c0, b_c1, c_c0 = 0;    // element coordinates of K block 0 (c0 = contiguous dim, c1 = row)
    if constexpr (OP == TC_FWD1) {
        out_dim = p.Hp; nkb = d.Pp / BLOCK_K;
        a_c0 = m0; a_c1 = (int)d.coff;                    // W1 [PT][Hp]: rows = k (predictor), cols = feature
        b_c0 = (int)d.coff; b_c1 = (int)row0;             // X  [rows][PT]
    } else if constexpr (OP == TC_FWD2) {
        out_dim = p.Op; nkb = p.Hp / BLOCK_K;
        a_c0 = m0; a_c1 = s * p.Hp;                       // W2 [S*Hp][Op]
        b_c0 = s * p.Hp; b_c1 = (int)row0;                // h  [rows][S*Hp]
        c_c0 = s * p.Op + m0;                             // Y tile
    } else {
        out_dim = p.Hp; nkb = p.Op / BLOCK_K;
        a_c0 = 0; a_c1 = s * p.Hp + m0;                   // W2 rows = hidden unit, k = output gene (contiguous)
        b_c0 = s * p.Op; b_c1 = 0;                        // dz2 [Bp][S*Op]
        c_c0 = s * p.Hp + m0;                             // h tile
    }
    if (m0 >= out_dim) return;                            // whole tile is padding (uniform per CTA)

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t b_stage_bytes = (uint32_t)p.n_cols * BLOCK_K * 4;
    // one slab: [A | B], what TMA writes per K block (TS: the A part is a plain [32][ts_wbox] tile)
    const uint32_t hi_bytes = ((TS && A_MN) ? (uint32_t)(BLOCK_K * p.ts_wbox * 4) : A_STAGE_BYTES) + b_stage_bytes;
    const uint32_t stage_bytes = A_STAGE_BYTES + b_stage_bytes;
    const uint32_t lo_stage_bytes = TS ? b_stage_bytes : stage_bytes;  // TS: only the activation slab has a residual in smem
    const int stages = p.stages;
    const int lo_stages = X3 ? p.lo_stages : 0;
    // shared memory: [raw ring: stages slabs][X3: residual ring: lo_stages slabs][aux tile]
    uint8_t* lo_base = smem + (size_t)stages * stage_bytes;
    const float* aux = reinterpret_cast<const float*>(lo_base + (size_t)lo_stages * lo_stage_bytes);
    __shared__ uint64_t full_bar[MAX_STAGES], empty_bar[MAX_STAGES], lo_ready[MAX_STAGES], lo_free[MAX_STAGES], tmem_full_bar, aux_bar;
    __shared__ uint32_t tmem_base_slot;
    __shared__ double red[8];
    __shared__ float gpart[TILE_M];      // X3: per-feature partial sums of the upper epilogue warps (bias gradients)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    DI_TRACE_T0(0);

    if (threadIdx.x == 0) {
        for (int i = 0; i < stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < lo_stages; ++i) { mbar_init(&lo_ready[i], NCONV / p.conv_teams); mbar_init(&lo_free[i], 1); }
        mbar_init(&tmem_full_bar, 1);
        mbar_init(&aux_bar, 1);
        fence_barrier_init();
    }
    if (warp == 4 && lane == 0) { prefetch_tensormap(&mapA); prefetch_tensormap(&mapB); }
    if (warp == 0) tmem_alloc(&tmem_base_slot, (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_slot;
    // the producer warp waits for the previous grid later: the operand that grid does not write is fetched first
    if (warp != 4) pdl_wait();
    if (p.pdl_early && threadIdx.x == 0) pdl_release();
    DI_TRACE_T0(1);

    // residual pass (X3): lo = a - trunc19(a) for every float of the slab TMA just delivered, written to the next slab
    // of the residual ring (element-wise: the swizzled layout carries over unchanged).  Run by 8 warps (0-3, 6-9).
    auto convert = [&](int cid_all) {
        // two teams (warps 0-3, warps 6-9): team t converts slabs t, t + 2, ...; one team: everybody converts every slab
        const int teams = p.conv_teams, CT = NCONV / teams;           // threads per slab
        const int team = teams == 2 ? (cid_all >> 7) : 0, cid = teams == 2 ? (cid_all & 127) : cid_all;
        const int per_thread = (int)(hi_bytes / 16) / CT;             // float4 per thread; hi_bytes is a multiple of 4 KB
        for (int kb = team; kb < nkb; kb += teams) {
            const int lstages = X3 ? lo_stages : 1;       // (plain TF32 never instantiates a call of this lambda)
            const int st = kb % stages, ls = kb % lstages;
            mbar_wait(&full_bar[st], (kb / stages) & 1, 1);
            if (kb >= lstages) mbar_wait(&lo_free[ls], ((kb / lstages) - 1) & 1, 11);
            if constexpr (TS) {
                // weights: this thread owns feature `quad * 32 + lane` (the TMEM lane its warp may write) and half of the
                // slab's 32 k; W and its residual go to tensor memory, columns [ls * 64, +32) and [ls * 64 + 32, +32)
                const int hw = (int)(threadIdx.x >> 5), quad = hw & 3, ml = quad * 32 + (cid & 31);
                const int h_first = teams == 2 ? 0 : (hw >= 6 ? 1 : 0), h_last = teams == 2 ? 1 : h_first;   // halves of the 32 k this thread converts
                for (int half = h_first; half <= h_last; ++half) {
                float hi[16], lo[16];
                if constexpr (A_MN) {
                    // FWD1 / FWD2: plain [32 k][wbox features] tile; a warp reads 32 consecutive features of one k
                    const uint32_t src = smem_u32(smem + (size_t)st * stage_bytes) + (uint32_t)(half * 16 * p.ts_wbox + ml) * 4u;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        hi[i] = 0.f;
                        if (ml < p.ts_wbox) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(hi[i]) : "r"(src + (uint32_t)(i * p.ts_wbox) * 4u));
                    }
                } else {
                    // BWD: K-major tile written by TMA with the 128-byte swizzle: row = feature (128 B = 32 k), the 16-byte
                    // chunk c of row m sits at chunk c ^ (m & 7); this thread's 16 k are chunks 4 half .. 4 half + 3
                    // (eight consecutive rows hit eight different chunks: conflict-free 16-byte loads)
                    const uint32_t rowb = smem_u32(smem + (size_t)st * stage_bytes) + (uint32_t)ml * 128u;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const uint32_t chunk = (uint32_t)((half * 4 + c) ^ (ml & 7));
                        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                                     : "=f"(hi[4 * c]), "=f"(hi[4 * c + 1]), "=f"(hi[4 * c + 2]), "=f"(hi[4 * c + 3]) : "r"(rowb + chunk * 16u));
                    }
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) lo[i] = tf32_residual(hi[i]);
                const uint32_t ta = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(p.ts_acol0 + ls * TS_COLS + half * 16);
                tmem_st16(ta, hi);
                tmem_st16(ta + BLOCK_K, lo);
                }
                // activations: residual slab in shared memory, as in the SS path but for the B part only
                const uint32_t bhi = smem_u32(smem + (size_t)st * stage_bytes + A_STAGE_BYTES) + cid * 16;
                const uint32_t blo = smem_u32(lo_base + (size_t)ls * lo_stage_bytes) + cid * 16;
                for (uint32_t off = 0; off + cid * 16 < b_stage_bytes; off += CT * 16) {
                    float4 v;
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(bhi + off));
                    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};"
                                 ::"r"(blo + off), "f"(tf32_residual(v.x)), "f"(tf32_residual(v.y)), "f"(tf32_residual(v.z)),
                                   "f"(tf32_residual(v.w)) : "memory");
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                fence_proxy_async();                      // generic-proxy writes -> visible to the tensor core
                tc_fence_before();                        // tensor-memory stores ordered before the arrive
                mbar_arrive(&lo_ready[ls]);
                if (kb < 40) DI_TRACE_T0(128 + kb);
                continue;
            }
            const uint32_t hi = smem_u32(smem + (size_t)st * stage_bytes) + cid * 16;
            const uint32_t lo = smem_u32(lo_base + (size_t)ls * stage_bytes) + cid * 16;
            for (int i0 = 0; i0 < per_thread; i0 += 4) {
                float4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (i0 + u < per_thread)
                        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                                     : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "r"(hi + (i0 + u) * (CT * 16)));
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (i0 + u < per_thread)
                        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};"
                                     ::"r"(lo + (i0 + u) * (CT * 16)), "f"(tf32_residual(v[u].x)), "f"(tf32_residual(v[u].y)),
                                       "f"(tf32_residual(v[u].z)), "f"(tf32_residual(v[u].w)) : "memory");
            }
            fence_proxy_async();                          // generic-proxy writes -> visible to the tensor core
            mbar_arrive(&lo_ready[ls]);
            if (kb < 40) DI_TRACE_T0(128 + kb);
        }
    };

    if (warp == 4) {
        // ===== TMA producer =====
        if (elect_one()) {
            auto load_a = [&](int kb, int st) {
                uint8_t* sa = smem + (size_t)st * stage_bytes;
                if constexpr (TS && A_MN) tma_load_2d(sa, &mapA, &full_bar[st], a_c0, a_c1 + kb * BLOCK_K);   // plain [32 k][wbox] tile
                else if constexpr (A_MN) load_stage<true>(sa, &mapA, &full_bar[st], a_c0, a_c1 + kb * BLOCK_K, TILE_M);
                else tma_load_2d(sa, &mapA, &full_bar[st], a_c0 + kb * BLOCK_K, a_c1);
            };
            auto load_b = [&](int kb, int st) {
                tma_load_2d(smem + (size_t)st * stage_bytes + A_STAGE_BYTES, &mapB, &full_bar[st], b_c0 + kb * BLOCK_K, b_c1);
            };
            auto load_aux = [&]() {                       // side operand of the epilogue, needed only after the MMAs
                mbar_arrive_expect_tx(&aux_bar, (uint32_t)(p.n_cols * p.aux_cols * 4));
                tma_load_2d((void*)aux, &mapC, &aux_bar, c_c0, (int32_t)p.aux_row0);
            };
            // What the grid ahead of this one in the step's chain does NOT write can be fetched before waiting for
            // it: FWD1 follows ADAM (writes W1, not the staged batch X); FWD2 follows FWD1 (writes h, not W2 or the
            // Y tile); BWD follows FWD2 (writes dz2, not W2 or the h tile).  Every earlier grid has completed by
            // the time this one was released (the releasing grid had passed its own wait).
            constexpr bool A_FIRST = (OP != TC_FWD1);
            const int npre = p.pdl_prefetch ? min(stages, nkb) : 0;
            for (int kb = 0; kb < npre; ++kb) {
                mbar_arrive_expect_tx(&full_bar[kb], hi_bytes);
                if (kb < 40) DI_TRACE(8 + kb);
                if constexpr (A_FIRST) load_a(kb, kb); else load_b(kb, kb);
            }
            if (npre && p.aux_cols > 0) load_aux();
            pdl_wait();
            for (int kb = 0; kb < npre; ++kb) {
                if constexpr (A_FIRST) load_b(kb, kb); else load_a(kb, kb);
            }
            for (int kb = npre; kb < nkb; ++kb) {
                const int st = kb % stages;
                if (kb >= stages) mbar_wait(&empty_bar[st], ((kb / stages) - 1) & 1, 2);
                if (kb < 40) DI_TRACE(8 + kb);
                mbar_arrive_expect_tx(&full_bar[st], hi_bytes);
                load_a(kb, st);
                load_b(kb, st);
                if (kb == 0 && p.aux_cols > 0) load_aux();
            }
        }
    } else if (warp == 5) {
        // ===== MMA issuer =====
        if (elect_one()) {
            const uint32_t idesc = idesc_for(p.n_cols, A_MN, false);
            for (int kb = 0; kb < nkb; ++kb) {
                const int st = kb % stages;
                mbar_wait(&full_bar[st], (kb / stages) & 1, 3);
                const int ls = X3 ? kb % lo_stages : 0;
                if (kb < 40) DI_TRACE(48 + kb);
                if constexpr (X3) mbar_wait(&lo_ready[ls], (kb / lo_stages) & 1, 10);
                tc_fence_after();
                if (kb < 40) DI_TRACE(88 + kb);
                const uint32_t sa = smem_u32(smem + (size_t)st * stage_bytes);
                const uint32_t sb = sa + A_STAGE_BYTES;
                if constexpr (TS) {
                    const uint32_t idesc_ts = idesc_for(p.n_cols, false, false);
                    const uint32_t sb_lo = smem_u32(lo_base + (size_t)ls * lo_stage_bytes);
                    const uint32_t ta = tmem + (uint32_t)(p.ts_acol0 + ls * TS_COLS);
                    // short chains (see acc_sum16): a b -> accumulator kb % nacc, the small products -> their own one
                    const uint32_t d_hi = tmem + (uint32_t)((kb % p.nacc) * p.n_cols);
                    const uint32_t d_lo = p.lo_acc ? tmem + (uint32_t)(p.nacc * p.n_cols) : d_hi;
                    const uint32_t first_lo = p.lo_acc ? (kb ? 1u : 0u) : (kb >= p.nacc ? 1u : 0u);   // accumulate flag of the block's first MMA
#pragma unroll
                    for (int j = 0; j < BLOCK_K / UMMA_K; ++j) {
                        umma_tf32_ts(d_lo, ta + BLOCK_K + j * UMMA_K, stage_desc<false>(sb, j), idesc_ts, j ? 1u : first_lo);
                        umma_tf32_ts(d_lo, ta + j * UMMA_K, stage_desc<false>(sb_lo, j), idesc_ts, 1u);
                        umma_tf32_ts(d_hi, ta + j * UMMA_K, stage_desc<false>(sb, j), idesc_ts, (!p.lo_acc || kb >= p.nacc || j) ? 1u : 0u);
                    }
                } else if constexpr (X3) {
                    const uint32_t sa_lo = smem_u32(lo_base + (size_t)ls * stage_bytes), sb_lo = sa_lo + A_STAGE_BYTES;
                    const uint32_t d_hi = tmem + (uint32_t)((kb % p.nacc) * p.n_cols);
                    const uint32_t d_lo = p.lo_acc ? tmem + (uint32_t)(p.nacc * p.n_cols) : d_hi;
                    const uint32_t first_lo = p.lo_acc ? (kb ? 1u : 0u) : (kb >= p.nacc ? 1u : 0u);   // accumulate flag of the block's first MMA
#pragma unroll
                    for (int j = 0; j < BLOCK_K / UMMA_K; ++j) {
                        umma_tf32(d_lo, stage_desc<A_MN>(sa_lo, j), stage_desc<false>(sb, j), idesc, j ? 1u : first_lo);
                        umma_tf32(d_lo, stage_desc<A_MN>(sa, j), stage_desc<false>(sb_lo, j), idesc, 1u);
                        umma_tf32(d_hi, stage_desc<A_MN>(sa, j), stage_desc<false>(sb, j), idesc, (!p.lo_acc || kb >= p.nacc || j) ? 1u : 0u);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < BLOCK_K / UMMA_K; ++j)
                        umma_tf32(tmem, stage_desc<A_MN>(sa, j), stage_desc<false>(sb, j), idesc, (kb | j) ? 1u : 0u);
                }
                umma_commit(&empty_bar[st]);
                if constexpr (X3) umma_commit(&lo_free[ls]);
                if (p.pdl_lead > 0 && kb == nkb - 1 - p.pdl_lead) pdl_release();
            }
            umma_commit(&tmem_full_bar);
        }
    } else {
        // ===== epilogue: a warp reads the TMEM lanes of its quadrant (warp % 4) = output features m0 + 32 quad .. =====
        // X3: the four converter-only warps (6-9) share the epilogue with warps 0-3: each quadrant has two warps, the
        // lower one takes the first half of the batch columns, the upper one the second half
        const int quad = warp & 3;
        const bool upper = warp >= 6;
        const int fl = quad * 32 + lane;                  // feature inside the tile
        const int f = m0 + fl;                            // output feature of this thread
        const bool f_ok = f < out_dim;
        const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16);
        const int ncol = p.n_cols;
        const int c_lo = (X3 && upper) ? ncol / 2 : 0;
        const int c_hi = (X3 && !upper) ? ncol / 2 : ncol;
        // global operands of the epilogue are fetched now, a whole main loop before they are needed: bias, the Adam
        // state of the bias this thread will update, the per-epoch tables of the epoch graph
        const int64_t bias_i = (int64_t)s * ((OP == TC_FWD2) ? p.Op : p.Hp) + f;
        float bias = 0.f, bw = 0.f, bm = 0.f, bv = 0.f;
        if (f_ok) {
            if constexpr (OP == TC_FWD1) bias = p.b1[bias_i];
            if constexpr (OP == TC_FWD2) bias = p.b2[bias_i];
            if (p.training && !upper) {
                if constexpr (OP == TC_FWD2) { bw = bias; bm = p.mb2[bias_i]; bv = p.vb2[bias_i]; }
                if constexpr (OP == TC_BWD) { bw = p.b1[bias_i]; bm = p.mb1[bias_i]; bv = p.vb1[bias_i]; }
            }
        }
        const AdamParams adam_b = adam_of(p);
        const uint32_t dstep = (OP == TC_FWD1) ? dropout_step(p) : 0u;
        if constexpr (X3) convert(upper ? (warp - 2) * 32 + lane : warp * 32 + lane);
        if (p.aux_cols > 0) mbar_wait(&aux_bar, 0, 5);
        DI_TRACE_T0(2);
        mbar_wait(&tmem_full_bar, 0, 4);
        tc_fence_after();
        if (threadIdx.x == 0) pdl_release();
        __syncwarp();
        DI_TRACE_T0(3);

        if constexpr (OP == TC_FWD1) {
            const bool drop = p.training && p.drop_thresh;
            float* hrow = p.Hact + row0 * p.ldh + (int64_t)s * p.Hp + f;
            float* hlo = (p.training && p.Hlo) ? p.Hlo + (int64_t)s * p.Hp + f : nullptr;   // training h starts at row 0
            for (int c = c_lo; c < c_hi; c += 16) {
                float v[16];
                __syncwarp();
                acc_sum16(taddr, c, ncol, p.nacc, min(p.nacc, nkb), X3 && p.lo_acc != 0, v);
                if (!f_ok) continue;
                // four independent Philox calls first (instruction-level parallelism), then the 16 activations
                uint32_t w[4][4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    w[q][0] = w[q][1] = w[q][2] = w[q][3] = 0xFFFFFFFFu;
                    if (drop) dropout_words((uint32_t)f, (uint32_t)((c >> 2) + q), (uint32_t)d.gid, dstep, p.seed, w[q]);
                }
                float a[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    a[i] = fmaxf(v[i] + bias, 0.f);
                    if (drop) a[i] = (w[i >> 2][i & 3] >= p.drop_thresh) ? a[i] * p.keep_scale : 0.f;
                }
                float* dst = hrow + (int64_t)c * p.ldh;
#pragma unroll
                for (int i = 0; i < 16; ++i) dst[(int64_t)i * p.ldh] = a[i];
                if (hlo) {
                    float* dlo = hlo + (int64_t)c * p.ldh;
#pragma unroll
                    for (int i = 0; i < 16; ++i) dlo[(int64_t)i * p.ldh] = tf32_residual(a[i]);
                }
            }
        } else if constexpr (OP == TC_FWD2) {
            const int64_t bi = bias_i;
            float part = 0.f, gsum = 0.f;
            const int rows_left = p.n_valid - row_tile * ncol;
            for (int c = c_lo; c < c_hi; c += 16) {
                float v[16], y[16];
                __syncwarp();
                acc_sum16(taddr, c, ncol, p.nacc, min(p.nacc, nkb), X3 && p.lo_acc != 0, v);
                if (!f_ok) continue;
                if (p.aux_cols > 0) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) y[i] = aux[(c + i) * p.aux_cols + fl];
                } else if (p.Y) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) y[i] = __ldg(p.Y + (row0 + c + i) * p.ldy + bi);
                }
                // phase 1: the 16 transcendental chains side by side (branch-free)
                float yh[16], sg[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) softplus_sigmoid(v[i] + bias, yh[i], sg[i]);
                // phase 2: consumers
                if (p.out && f < p.O) {
                    float* dst = p.out + ((int64_t)row_tile * ncol + c) * p.ld_out + (int64_t)s * p.O + f;
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (c + i < rows_left) dst[(int64_t)i * p.ld_out] = yh[i];
                }
                if (p.Y) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) { const float diff = y[i] - yh[i]; part += y[i] * diff * diff; }
                    if (p.training) {
                        float g[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) { g[i] = 2.0f * y[i] * (yh[i] - y[i]) * sg[i] * p.inv_norm; gsum += g[i]; }
                        const int64_t ld2 = (int64_t)p.S * p.Op;
                        float* dst = p.DZ2 + (int64_t)c * ld2 + bi;
#pragma unroll
                        for (int i = 0; i < 16; ++i) dst[(int64_t)i * ld2] = g[i];
                        if (p.DZ2lo) {
                            float* dlo = p.DZ2lo + (int64_t)c * ld2 + bi;
#pragma unroll
                            for (int i = 0; i < 16; ++i) dlo[(int64_t)i * ld2] = tf32_residual(g[i]);
                        }
                    }
                }
            }
            if (p.training || p.loss) {
                double dpart = (double)part;
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) dpart += __shfl_xor_sync(0xffffffffu, dpart, off);
                if (lane == 0) red[upper ? 4 + quad : quad] = dpart;
                if (X3 && upper) gpart[fl] = gsum;
                named_bar_sync(1, X3 ? 256 : 128);
                if (!upper) {
                    if constexpr (X3) gsum += gpart[fl];
                    if (p.training && f_ok) {
                        adam_update_fast(gsum, bw, bm, bv, adam_b);
                        p.b2[bi] = bw; p.mb2[bi] = bm; p.vb2[bi] = bv;
                    }
                    if (p.loss && threadIdx.x == 0) {
                        double tot = red[0] + red[1] + red[2] + red[3];
                        if constexpr (X3) tot += red[4] + red[5] + red[6] + red[7];
                        atomicAdd(p.loss, tot);
                    }
                }
            }
        } else {
            const int64_t bi = bias_i;
            float gsum = 0.f;
            for (int c = c_lo; c < c_hi; c += 16) {
                float v[16], h[16];
                __syncwarp();
                acc_sum16(taddr, c, ncol, p.nacc, min(p.nacc, nkb), X3 && p.lo_acc != 0, v);
                if (!f_ok) continue;
                if (p.aux_cols > 0) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) h[i] = aux[(c + i) * p.aux_cols + fl];
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) h[i] = p.Hact[(int64_t)(c + i) * p.S * p.Hp + bi];
                }
                float g[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) { g[i] = (h[i] > 0.f) ? v[i] * p.keep_scale : 0.f; gsum += g[i]; }
                const int64_t ld1 = (int64_t)p.S * p.Hp;
                float* dst = p.DZ1 + (int64_t)c * ld1 + bi;
#pragma unroll
                for (int i = 0; i < 16; ++i) dst[(int64_t)i * ld1] = g[i];
                if (p.DZ1lo) {
                    float* dlo = p.DZ1lo + (int64_t)c * ld1 + bi;
#pragma unroll
                    for (int i = 0; i < 16; ++i) dlo[(int64_t)i * ld1] = tf32_residual(g[i]);
                }
            }
            if constexpr (X3) {
                if (upper) gpart[fl] = gsum;
                named_bar_sync(1, 256);
                if (!upper) gsum += gpart[fl];
            }
            if (!upper && f_ok) {
                adam_update_fast(gsum, bw, bm, bv, adam_b);
                p.b1[bi] = bw; p.mb1[bi] = bm; p.vb1[bi] = bv;
            }
        }
    }

    DI_TRACE_T0(4);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, (uint32_t)p.tmem_cols);
    DI_TRACE_T0(5);
}

// ============================================================================ FWD1 / FWD2 / BWD, every operand by TMA
// LT ("lo by TMA") form of the error-compensated kernels, the default of DI_MATH_TF32X3.  The residual twin of every
// operand already exists in HBM when the kernel starts: activations get theirs from the epilogue (h, dz2) or the
// staging gather (X) that produces them -- the ADAM kernel needs those twins anyway -- and the weights get theirs from
// the ADAM epilogue, which writes W_lo = W - trunc_tf32(W) next to every updated W.  A K block is then four plain TMA
// tiles [A | A_lo | B | B_lo] and twelve MMAs; nobody touches the slabs between TMA and the tensor core, so the
// converter warps of tc_kernel<.., true> (1250 cycles per K block, profiles/traces/r01i) and their tensor-memory
// staging are gone and the K loop runs at the pace of the MMAs.  All sixteen non-producer warps belong to the epilogue:
// four per TMEM lane quadrant, each a quarter of the batch columns.
// Warp roles (576 threads): warps 0-15 epilogue, warp 16 TMA producer, warp 17 MMA issuer.
// Split K (gridDim.x = KS > 1; used when a GPU holds so few sub-networks that most SMs would idle while one CTA walks
// the whole K loop): CTA x accumulates K blocks [x nkb / KS, (x + 1) nkb / KS), writes its accumulator to a scratch
// tile in global memory (L2) and counts itself in; the CTA that arrives last adds the KS partial tiles in index order
// -- the sum does not depend on which CTA that is -- and runs the epilogue.  Nobody waits for anybody, so the CTAs of a
// tile need not be co-resident.
constexpr int LT_EPI_WARPS = 16;
constexpr int LT_THREADS = (LT_EPI_WARPS + 2) * 32;
constexpr int LT_MAX_STAGES = 6;

struct LtMaps { CUtensorMap A, Alo, B, Blo, C; };

template <int OP>
__global__ void __launch_bounds__(LT_THREADS, 1) tc_lt_kernel(const __grid_constant__ LtMaps maps, const TcParams p) {
    constexpr bool A_MN = (OP != TC_BWD);

    const int s = blockIdx.z + p.s_base;
    const SubnetDesc d = p.desc[s];
    const int m_tile = (int)(blockIdx.y % p.m_tiles);
    const int row_tile = (int)(blockIdx.y / p.m_tiles);
    const int m0 = m_tile * TILE_M;
    const int64_t row0 = p.row0 + (int64_t)row_tile * p.rows_per_block_y;

    int out_dim, nkb_all;
    int a_c0, a_c1, b_c0, b_c1, c_c0 = 0;
    if constexpr (OP == TC_FWD1) {
        out_dim = p.Hp; nkb_all = d.Pp / BLOCK_K;
        a_c0 = m0; a_c1 = (int)d.coff;
        b_c0 = (int)d.coff; b_c1 = (int)row0;
    } else if constexpr (OP == TC_FWD2) {
        out_dim = p.Op; nkb_all = p.Hp / BLOCK_K;
        a_c0 = m0; a_c1 = s * p.Hp;
        b_c0 = s * p.Hp; b_c1 = (int)row0;
        c_c0 = s * p.Op + m0;
    } else {
        out_dim = p.Hp; nkb_all = p.Op / BLOCK_K;
        a_c0 = 0; a_c1 = s * p.Hp + m0;
        b_c0 = s * p.Op; b_c1 = 0;
        c_c0 = s * p.Hp + m0;
    }
    if (m0 >= out_dim) return;                            // whole tile is padding (uniform per CTA and per cluster)
    // split K: this CTA's share of the K blocks
    const int ks = (int)gridDim.x, kx = (int)blockIdx.x;
    const int kb_begin = (int)((int64_t)nkb_all * kx / ks), kb_end = (int)((int64_t)nkb_all * (kx + 1) / ks);
    const int nkb = kb_end - kb_begin;

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t b_bytes = (uint32_t)p.n_cols * BLOCK_K * 4;
    const uint32_t stage_bytes = 2 * A_STAGE_BYTES + 2 * b_bytes;          // [A | A_lo | B | B_lo]
    const int stages = p.stages;
    const float* aux = reinterpret_cast<const float*>(smem + (size_t)stages * stage_bytes);
    // after the main loop the ring is dead: its first bytes hold the partial bias-gradient sums of the epilogue
    // (3 x 128 floats); split K: the partial accumulators of the other CTAs of the cluster follow at +2 KB
    float* gpart = reinterpret_cast<float*>(smem);
    __shared__ uint64_t full_bar[LT_MAX_STAGES], empty_bar[LT_MAX_STAGES], tmem_full_bar, aux_bar;
    __shared__ uint32_t tmem_base_slot, arrival_slot;
    __shared__ double red[LT_EPI_WARPS];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    DI_TRACE_T0(0);

    if (threadIdx.x == 0) {
        for (int i = 0; i < stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        mbar_init(&tmem_full_bar, 1);
        mbar_init(&aux_bar, 1);
        fence_barrier_init();
    }
    if (warp == LT_EPI_WARPS && lane == 0) {
        prefetch_tensormap(&maps.A); prefetch_tensormap(&maps.Alo); prefetch_tensormap(&maps.B); prefetch_tensormap(&maps.Blo);
    }
    if (warp == 0) tmem_alloc(&tmem_base_slot, (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_slot;
    if (warp != LT_EPI_WARPS) pdl_wait();                 // the producer waits later: see below
    // few sub-networks per GPU: SMs are idle, so the dependent grid may as well take its seats now, run its prologue
    // and fetch its weights while this grid works (it still waits for this grid's completion before touching its output)
    if (p.pdl_early && threadIdx.x == 0) pdl_release();
    DI_TRACE_T0(1);

    if (warp == LT_EPI_WARPS) {
        // ===== TMA producer =====
        if (elect_one()) {
            auto load_a = [&](int kb, int st) {
                uint8_t* sa = smem + (size_t)st * stage_bytes;
                const int k = (kb_begin + kb) * BLOCK_K;
                if constexpr (A_MN) {
                    load_stage<true>(sa, &maps.A, &full_bar[st], a_c0, a_c1 + k, TILE_M);
                    load_stage<true>(sa + A_STAGE_BYTES, &maps.Alo, &full_bar[st], a_c0, a_c1 + k, TILE_M);
                } else {
                    tma_load_2d(sa, &maps.A, &full_bar[st], a_c0 + k, a_c1);
                    tma_load_2d(sa + A_STAGE_BYTES, &maps.Alo, &full_bar[st], a_c0 + k, a_c1);
                }
            };
            auto load_b = [&](int kb, int st) {
                uint8_t* sb = smem + (size_t)st * stage_bytes + 2 * A_STAGE_BYTES;
                const int k = (kb_begin + kb) * BLOCK_K;
                tma_load_2d(sb, &maps.B, &full_bar[st], b_c0 + k, b_c1);
                tma_load_2d(sb + b_bytes, &maps.Blo, &full_bar[st], b_c0 + k, b_c1);
            };
            auto load_aux = [&]() {
                mbar_arrive_expect_tx(&aux_bar, (uint32_t)(p.n_cols * p.aux_cols * 4));
                tma_load_2d((void*)aux, &maps.C, &aux_bar, c_c0, (int32_t)p.aux_row0);
            };
            // what the grid ahead in the step's chain does NOT write is fetched before waiting for it: FWD1 follows
            // ADAM (writes W1 / W1_lo, not the staged batch); FWD2 follows FWD1 (writes h / h_lo, not W2 or the Y
            // tile); BWD follows FWD2 (writes dz2 / dz2_lo, not W2 or the h tile)
            constexpr bool A_FIRST = (OP != TC_FWD1);
            const bool want_aux = p.aux_cols > 0;        // split K: any CTA of the tile may be the one that runs the epilogue
            const int npre = p.pdl_prefetch ? min(stages, nkb) : 0;
            for (int kb = 0; kb < npre; ++kb) {
                mbar_arrive_expect_tx(&full_bar[kb], stage_bytes);
                if (kb < 40) DI_TRACE(8 + kb);
                if constexpr (A_FIRST) load_a(kb, kb); else load_b(kb, kb);
            }
            if (npre && want_aux) load_aux();
            pdl_wait();
            for (int kb = 0; kb < npre; ++kb) {
                if constexpr (A_FIRST) load_b(kb, kb); else load_a(kb, kb);
            }
            for (int kb = npre; kb < nkb; ++kb) {
                const int st = kb % stages;
                if (kb >= stages) mbar_wait(&empty_bar[st], ((kb / stages) - 1) & 1, 2);
                if (kb < 40) DI_TRACE(8 + kb);
                mbar_arrive_expect_tx(&full_bar[st], stage_bytes);
                load_a(kb, st);
                load_b(kb, st);
                if (kb == 0 && want_aux) load_aux();
            }
            if (nkb == 0 && !npre && want_aux) load_aux();
        }
    } else if (warp == LT_EPI_WARPS + 1) {
        // ===== MMA issuer: per K step a_lo b + a b_lo + a b (small terms first) =====
        if (elect_one()) {
            const uint32_t idesc = idesc_for(p.n_cols, A_MN, false);
            for (int kb = 0; kb < nkb; ++kb) {
                const int st = kb % stages;
                mbar_wait(&full_bar[st], (kb / stages) & 1, 3);
                tc_fence_after();
                if (kb < 40) DI_TRACE(48 + kb);
                const uint32_t sa = smem_u32(smem + (size_t)st * stage_bytes);
                const uint32_t sa_lo = sa + A_STAGE_BYTES, sb = sa + 2 * A_STAGE_BYTES, sb_lo = sb + b_bytes;
                // short chains (see acc_sum16): a b -> accumulator kb % nacc, the small products -> their own one
                const uint32_t d_hi = tmem + (uint32_t)((kb % p.nacc) * p.n_cols);
                const uint32_t d_lo = p.lo_acc ? tmem + (uint32_t)(p.nacc * p.n_cols) : d_hi;
                const bool fresh_hi = kb < p.nacc;            // first K block of this accumulator
#pragma unroll
                for (int j = 0; j < BLOCK_K / UMMA_K; ++j) {
                    if (p.lo_acc) {
                        umma_tf32(d_lo, stage_desc<A_MN>(sa_lo, j), stage_desc<false>(sb, j), idesc, (kb | j) ? 1u : 0u);
                        umma_tf32(d_lo, stage_desc<A_MN>(sa, j), stage_desc<false>(sb_lo, j), idesc, 1u);
                        umma_tf32(d_hi, stage_desc<A_MN>(sa, j), stage_desc<false>(sb, j), idesc, (!fresh_hi || j) ? 1u : 0u);
                    } else {
                        umma_tf32(d_hi, stage_desc<A_MN>(sa_lo, j), stage_desc<false>(sb, j), idesc, (!fresh_hi || j) ? 1u : 0u);
                        umma_tf32(d_hi, stage_desc<A_MN>(sa, j), stage_desc<false>(sb_lo, j), idesc, 1u);
                        umma_tf32(d_hi, stage_desc<A_MN>(sa, j), stage_desc<false>(sb, j), idesc, 1u);
                    }
                }
                umma_commit(&empty_bar[st]);
            }
            umma_commit(&tmem_full_bar);
        }
    } else {
        // ===== epilogue: warp -> TMEM lane quadrant (warp % 4) and column group (warp / 4) =====
        const int quad = warp & 3, cg = warp >> 2;
        const int fl = quad * 32 + lane;
        const int f = m0 + fl;
        const bool f_ok = f < out_dim;
        const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16);
        const int ncol = p.n_cols;
        const int cpg = ((ncol + 63) / 64) * 16;          // columns per group: a multiple of 16 (tcgen05.ld.x16)
        const int c_lo = min(ncol, cg * cpg), c_hi = min(ncol, c_lo + cpg);
        const bool first = cg == 0;                       // this group finishes the per-feature sums
        const int64_t bias_i = (int64_t)s * ((OP == TC_FWD2) ? p.Op : p.Hp) + f;
        float bias = 0.f, bw = 0.f, bm = 0.f, bv = 0.f;
        if (f_ok) {
            if constexpr (OP == TC_FWD1) bias = p.b1[bias_i];
            if constexpr (OP == TC_FWD2) bias = p.b2[bias_i];
            if (p.training && first) {
                if constexpr (OP == TC_FWD2) { bw = bias; bm = p.mb2[bias_i]; bv = p.vb2[bias_i]; }
                if constexpr (OP == TC_BWD) { bw = p.b1[bias_i]; bm = p.mb1[bias_i]; bv = p.vb1[bias_i]; }
            }
        }
        const AdamParams adam_b = adam_of(p);
        const uint32_t dstep = (OP == TC_FWD1) ? dropout_step(p) : 0u;
        // dropout keep-bits of this thread's columns: forty Philox rounds per 16 columns that depend on nothing the
        // main loop produces, so they are drawn while the tensor core works (bit i: column c_lo + i is kept)
        uint64_t keep = ~0ull;
        if constexpr (OP == TC_FWD1) {
            if (p.training && p.drop_thresh && f_ok) {
                keep = 0ull;
                for (int c = c_lo; c < c_hi; c += 4) {
                    uint32_t w[4];
                    dropout_words((uint32_t)f, (uint32_t)(c >> 2), (uint32_t)d.gid, dstep, p.seed, w);
#pragma unroll
                    for (int i = 0; i < 4; ++i) keep |= (uint64_t)(w[i] >= p.drop_thresh ? 1u : 0u) << (c - c_lo + i);
                }
            }
        }
        if (p.aux_cols > 0) mbar_wait(&aux_bar, 0, 5);
        DI_TRACE_T0(2);
        if (nkb > 0) mbar_wait(&tmem_full_bar, 0, 4);
        tc_fence_after();
        // split K: park this CTA's accumulator in the scratch tile, count in, and carry on only as the last arrival
        const float* ksum = nullptr;                      // != nullptr: the epilogue sums the KS scratch tiles instead of reading TMEM
        if (ks > 1) {
            const int64_t tile = ((int64_t)s * p.m_tiles + m_tile) * ks;
            float* mine = p.kpart + (tile + kx) * (int64_t)ncol * TILE_M + fl;
            for (int c = c_lo; c < c_hi; c += 16) {
                float v[16];
                __syncwarp();
                if (nkb > 0) acc_sum16(taddr, c, ncol, p.nacc, min(p.nacc, nkb), p.lo_acc != 0, v);
#pragma unroll
                for (int i = 0; i < 16; ++i) mine[(int64_t)(c + i) * TILE_M] = nkb > 0 ? v[i] : 0.f;
            }
            // the CTA barrier orders every thread's stores before thread 0's release at GPU scope (cumulativity), and
            // thread 0's acquire before every thread's loads after the second barrier: one atomic instead of 512 fences
            named_bar_sync(1, LT_EPI_WARPS * 32);
            if (threadIdx.x == 0) {
                unsigned int* cnt = p.kcount + (int64_t)s * p.m_tiles + m_tile;
                unsigned int prev;
                asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(prev) : "l"(cnt) : "memory");
                if (prev == (unsigned int)(ks - 1)) *cnt = 0u;     // everybody has arrived: ready for the next launch
                arrival_slot = prev;
            }
            named_bar_sync(1, LT_EPI_WARPS * 32);
            if (arrival_slot == (unsigned int)(ks - 1)) ksum = p.kpart + tile * (int64_t)ncol * TILE_M + fl;
        }
        const bool run_epilogue = ks == 1 || ksum != nullptr;
        if (threadIdx.x == 0) pdl_release();
        __syncwarp();
        DI_TRACE_T0(3);
        auto load_acc = [&](int c, float (&v)[16]) {
            __syncwarp();
            if (ksum) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = 0.f;
                for (int k = 0; k < ks; ++k) {
                    const float* src = ksum + ((int64_t)k * ncol + c) * TILE_M;
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] += __ldcg(src + (int64_t)i * TILE_M);
                }
            } else {
                acc_sum16(taddr, c, ncol, p.nacc, min(p.nacc, nkb), p.lo_acc != 0, v);
            }
        };

        if (!run_epilogue) {
            // another CTA of this tile finishes it
        } else if constexpr (OP == TC_FWD1) {
            const bool drop = p.training && p.drop_thresh;
            float* hrow = p.Hact + row0 * p.ldh + (int64_t)s * p.Hp + f;
            float* hlo = p.Hlo ? p.Hlo + row0 * p.ldh + (int64_t)s * p.Hp + f : nullptr;
            for (int c = c_lo; c < c_hi; c += 16) {
                float v[16];
                load_acc(c, v);
                if (!f_ok) continue;
                float a[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    a[i] = fmaxf(v[i] + bias, 0.f);
                    if (drop) a[i] = ((keep >> (c - c_lo + i)) & 1ull) ? a[i] * p.keep_scale : 0.f;
                }
                float* dst = hrow + (int64_t)c * p.ldh;
#pragma unroll
                for (int i = 0; i < 16; ++i) dst[(int64_t)i * p.ldh] = a[i];
                if (hlo) {
                    float* dlo = hlo + (int64_t)c * p.ldh;
#pragma unroll
                    for (int i = 0; i < 16; ++i) dlo[(int64_t)i * p.ldh] = tf32_residual(a[i]);
                }
            }
        } else if constexpr (OP == TC_FWD2) {
            const int64_t bi = bias_i;
            float part = 0.f, gsum = 0.f;
            const int rows_left = p.n_valid - row_tile * ncol;
            for (int c = c_lo; c < c_hi; c += 16) {
                float v[16], y[16];
                load_acc(c, v);
                if (!f_ok) continue;
                if (p.aux_cols > 0) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) y[i] = aux[(c + i) * p.aux_cols + fl];
                } else if (p.Y) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) y[i] = __ldg(p.Y + (row0 + c + i) * p.ldy + bi);
                }
                float yh[16], sg[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) softplus_sigmoid(v[i] + bias, yh[i], sg[i]);
                if (p.out && f < p.O) {
                    float* dst = p.out + ((int64_t)row_tile * ncol + c) * p.ld_out + (int64_t)s * p.O + f;
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (c + i < rows_left) dst[(int64_t)i * p.ld_out] = yh[i];
                }
                if (p.Y) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) { const float diff = y[i] - yh[i]; part += y[i] * diff * diff; }
                    if (p.training) {
                        float g[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) { g[i] = 2.0f * y[i] * (yh[i] - y[i]) * sg[i] * p.inv_norm; gsum += g[i]; }
                        const int64_t ld2 = (int64_t)p.S * p.Op;
                        float* dst = p.DZ2 + (int64_t)c * ld2 + bi;
#pragma unroll
                        for (int i = 0; i < 16; ++i) dst[(int64_t)i * ld2] = g[i];
                        if (p.DZ2lo) {
                            float* dlo = p.DZ2lo + (int64_t)c * ld2 + bi;
#pragma unroll
                            for (int i = 0; i < 16; ++i) dlo[(int64_t)i * ld2] = tf32_residual(g[i]);
                        }
                    }
                }
            }
            if (p.training || p.loss) {
                double dpart = (double)part;
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) dpart += __shfl_xor_sync(0xffffffffu, dpart, off);
                if (lane == 0) red[warp] = dpart;
                if (!first) gpart[(cg - 1) * TILE_M + fl] = gsum;
                named_bar_sync(1, LT_EPI_WARPS * 32);
                if (first) {
                    gsum += gpart[fl] + gpart[TILE_M + fl] + gpart[2 * TILE_M + fl];
                    if (p.training && f_ok) {
                        adam_update_fast(gsum, bw, bm, bv, adam_b);
                        p.b2[bi] = bw; p.mb2[bi] = bm; p.vb2[bi] = bv;
                    }
                    if (p.loss && threadIdx.x == 0) {
                        double tot = 0.0;
#pragma unroll
                        for (int i = 0; i < LT_EPI_WARPS; ++i) tot += red[i];
                        atomicAdd(p.loss, tot);
                    }
                }
            }
        } else {
            const int64_t bi = bias_i;
            float gsum = 0.f;
            for (int c = c_lo; c < c_hi; c += 16) {
                float v[16], h[16];
                load_acc(c, v);
                if (!f_ok) continue;
                if (p.aux_cols > 0) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) h[i] = aux[(c + i) * p.aux_cols + fl];
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) h[i] = p.Hact[(int64_t)(c + i) * p.S * p.Hp + bi];
                }
                float g[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) { g[i] = (h[i] > 0.f) ? v[i] * p.keep_scale : 0.f; gsum += g[i]; }
                const int64_t ld1 = (int64_t)p.S * p.Hp;
                float* dst = p.DZ1 + (int64_t)c * ld1 + bi;
#pragma unroll
                for (int i = 0; i < 16; ++i) dst[(int64_t)i * ld1] = g[i];
                if (p.DZ1lo) {
                    float* dlo = p.DZ1lo + (int64_t)c * ld1 + bi;
#pragma unroll
                    for (int i = 0; i < 16; ++i) dlo[(int64_t)i * ld1] = tf32_residual(g[i]);
                }
            }
            if (!first) gpart[(cg - 1) * TILE_M + fl] = gsum;
            named_bar_sync(1, LT_EPI_WARPS * 32);
            if (first && f_ok) {
                gsum += gpart[fl] + gpart[TILE_M + fl] + gpart[2 * TILE_M + fl];
                adam_update_fast(gsum, bw, bm, bv, adam_b);
                p.b1[bi] = bw; p.mb1[bi] = bm; p.vb1[bi] = bv;
            }
        }
    }

    DI_TRACE_T0(4);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, (uint32_t)p.tmem_cols);
    DI_TRACE_T0(5);
}

// ================================================================== inference: persistent, epilogue under the main loop
// The two forward layers over all cells (model.predict, multinet.py:278-280, and the validation pass of every epoch).
// One CTA per SM walks tiles t = blockIdx.x, + gridDim.x, ... of [128 features x 128 cells] (row tile fastest, then feature
// tile, then sub-network).  Operands as in tc_lt_kernel: [W | W_lo | act | act_lo] slabs by TMA, twelve MMAs per K block.
// Tensor memory holds TWO accumulator stages of two tiles each (a b products / the small products, 4 x 128 = 512 columns):
// while the sixteen epilogue warps drain stage i & 1 (bias, relu or softplus, optional wMSE, stores), the MMA warp is
// already filling the other stage with the next tile, and the TMA ring never drains between tiles.  In tc_lt_kernel (one
// tile per CTA) the tensor pipe idles during prologue, first-slab latency and the whole epilogue: 44 % / 33 % active in
// FWD1 / FWD2 (profiles/r02_full_infer_lt.md).
constexpr int INF_TILE = 128;

template <int OP>
__global__ void __launch_bounds__(LT_THREADS, 1) tc_lt_infer_kernel(const __grid_constant__ LtMaps maps, const TcParams p) {
    static_assert(OP == TC_FWD1 || OP == TC_FWD2, "inference runs the two forward layers");
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    constexpr uint32_t b_bytes = INF_TILE * BLOCK_K * 4;
    constexpr uint32_t stage_bytes = 2 * A_STAGE_BYTES + 2 * b_bytes;           // 64 KB
    const int stages = p.stages;
    __shared__ uint64_t full_bar[LT_MAX_STAGES], empty_bar[LT_MAX_STAGES], tmem_full_bar[2], tmem_empty_bar[2];
    __shared__ uint32_t tmem_base_slot;
    __shared__ double red[2][LT_EPI_WARPS];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full_bar[i], 1); mbar_init(&tmem_empty_bar[i], LT_EPI_WARPS); }
        fence_barrier_init();
    }
    if (warp == LT_EPI_WARPS && lane == 0) {
        prefetch_tensormap(&maps.A); prefetch_tensormap(&maps.Alo); prefetch_tensormap(&maps.B); prefetch_tensormap(&maps.Blo);
    }
    if (warp == 0) tmem_alloc(&tmem_base_slot, 512u);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_slot;

    const int row_tiles = p.row_tiles, m_tiles = p.m_tiles;
    const int n_tiles = p.S * m_tiles * row_tiles;
    // tile t -> (sub-network, feature tile, row tile) and its operand coordinates
    auto tile_of = [&](int t, int& s, int& m0, int& row_tile, int& nkb, int& a_c1, int& b_c0) {
        row_tile = t % row_tiles;
        const int mt = (t / row_tiles) % m_tiles;
        s = t / (row_tiles * m_tiles) + p.s_base;
        m0 = mt * TILE_M;
        if constexpr (OP == TC_FWD1) { const SubnetDesc d = p.desc[s]; nkb = d.Pp / BLOCK_K; a_c1 = (int)d.coff; b_c0 = (int)d.coff; }
        else { nkb = p.Hp / BLOCK_K; a_c1 = s * p.Hp; b_c0 = s * p.Hp; }
    };

    if (warp == LT_EPI_WARPS) {
        // ===== TMA producer: the ring runs across tiles =====
        if (elect_one()) {
            int it = 0;
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                int s, m0, row_tile, nkb, a_c1, b_c0;
                tile_of(t, s, m0, row_tile, nkb, a_c1, b_c0);
                const int b_c1 = (int)(p.row0 + (int64_t)row_tile * INF_TILE);
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int st = it % stages;
                    if (it >= stages) mbar_wait(&empty_bar[st], ((it / stages) - 1) & 1, 2);
                    mbar_arrive_expect_tx(&full_bar[st], stage_bytes);
                    uint8_t* sa = smem + (size_t)st * stage_bytes;
                    const int k = kb * BLOCK_K;
                    load_stage<true>(sa, &maps.A, &full_bar[st], m0, a_c1 + k, TILE_M);
                    load_stage<true>(sa + A_STAGE_BYTES, &maps.Alo, &full_bar[st], m0, a_c1 + k, TILE_M);
                    tma_load_2d(sa + 2 * A_STAGE_BYTES, &maps.B, &full_bar[st], b_c0 + k, b_c1);
                    tma_load_2d(sa + 2 * A_STAGE_BYTES + b_bytes, &maps.Blo, &full_bar[st], b_c0 + k, b_c1);
                }
            }
        }
    } else if (warp == LT_EPI_WARPS + 1) {
        // ===== MMA issuer =====
        if (elect_one()) {
            const uint32_t idesc = idesc_for(INF_TILE, true, false);
            int it = 0, i = 0;
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++i) {
                int s, m0, row_tile, nkb, a_c1, b_c0;
                tile_of(t, s, m0, row_tile, nkb, a_c1, b_c0);
                const int acc = i & 1;
                mbar_wait(&tmem_empty_bar[acc], ((i >> 1) & 1) ^ 1, 12);     // the epilogue has drained this stage (passes at once the first time)
                tc_fence_after();
                const uint32_t d_hi = tmem + (uint32_t)(acc * 2 * INF_TILE), d_lo = d_hi + INF_TILE;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int st = it % stages;
                    mbar_wait(&full_bar[st], (it / stages) & 1, 3);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + (size_t)st * stage_bytes);
                    const uint32_t sa_lo = sa + A_STAGE_BYTES, sb = sa + 2 * A_STAGE_BYTES, sb_lo = sb + b_bytes;
#pragma unroll
                    for (int j = 0; j < BLOCK_K / UMMA_K; ++j) {
                        umma_tf32(d_lo, stage_desc<true>(sa_lo, j), stage_desc<false>(sb, j), idesc, (kb | j) ? 1u : 0u);
                        umma_tf32(d_lo, stage_desc<true>(sa, j), stage_desc<false>(sb_lo, j), idesc, 1u);
                        umma_tf32(d_hi, stage_desc<true>(sa, j), stage_desc<false>(sb, j), idesc, (kb | j) ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[st]);
                }
                umma_commit(&tmem_full_bar[acc]);
            }
        }
    } else {
        // ===== epilogue: warp -> TMEM lane quadrant (warp % 4), 32 of the 128 cell columns (warp / 4) =====
        const int quad = warp & 3, cg = warp >> 2;
        const int fl = quad * 32 + lane;
        int i = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++i) {
            int s, m0, row_tile, nkb, a_c1, b_c0;
            tile_of(t, s, m0, row_tile, nkb, a_c1, b_c0);
            const int acc = i & 1;
            const int out_dim = (OP == TC_FWD1) ? p.Hp : p.Op;
            const int f = m0 + fl;
            const bool f_ok = f < out_dim;
            const int64_t bias_i = (int64_t)s * out_dim + f;
            const float bias = f_ok ? ((OP == TC_FWD1) ? p.b1[bias_i] : p.b2[bias_i]) : 0.f;
            const int64_t row0 = p.row0 + (int64_t)row_tile * INF_TILE;
            const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * 2 * INF_TILE);
            mbar_wait(&tmem_full_bar[acc], (i >> 1) & 1, 4);
            tc_fence_after();
            float part = 0.f;
            const int rows_left = p.n_valid - row_tile * INF_TILE;
#pragma unroll 1
            for (int c = cg * 32; c < cg * 32 + 32; c += 16) {
                float v[16];
                __syncwarp();
                acc_sum16(taddr, c, INF_TILE, 1, 1, true, v);
                if (m0 >= out_dim || !f_ok) continue;
                if constexpr (OP == TC_FWD1) {
                    float* dst = p.Hact + (row0 + c) * p.ldh + (int64_t)s * p.Hp + f;
                    float* dlo = p.Hlo + (row0 + c) * p.ldh + (int64_t)s * p.Hp + f;
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        const float a = fmaxf(v[k] + bias, 0.f);
                        dst[(int64_t)k * p.ldh] = a;
                        dlo[(int64_t)k * p.ldh] = tf32_residual(a);
                    }
                } else {
                    float yh[16], sg[16];
#pragma unroll
                    for (int k = 0; k < 16; ++k) softplus_sigmoid(v[k] + bias, yh[k], sg[k]);
                    if (p.out && f < p.O) {
                        float* dst = p.out + ((int64_t)row_tile * INF_TILE + c) * p.ld_out + (int64_t)s * p.O + f;
#pragma unroll
                        for (int k = 0; k < 16; ++k)
                            if (c + k < rows_left) dst[(int64_t)k * p.ld_out] = yh[k];
                    }
                    if (p.Y) {
#pragma unroll
                        for (int k = 0; k < 16; ++k) {
                            const float y = __ldg(p.Y + (row0 + c + k) * p.ldy + bias_i);
                            const float diff = y - yh[k];
                            part += y * diff * diff;
                        }
                    }
                }
            }
            // this warp has read its part of the stage: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
            if constexpr (OP == TC_FWD2) {
                if (p.loss) {
                    double dpart = (double)part;
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) dpart += __shfl_xor_sync(0xffffffffu, dpart, off);
                    if (lane == 0) red[acc][warp] = dpart;
                    named_bar_sync(1, LT_EPI_WARPS * 32);
                    if (threadIdx.x == 0) {
                        double tot = 0.0;
#pragma unroll
                        for (int w = 0; w < LT_EPI_WARPS; ++w) tot += red[acc][w];
                        atomicAdd(p.loss, tot);
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512u);
}

// ============================================================================================ ADAM (weight update)
// One CTA: dW tile [128 output features (lanes)] x [n_cols input features (columns)] = dout^T in, K = padded batch.
// shared memory: operands (all K blocks at once) | ring of AD_STAGES x {w, m, v} x [AD_R rows][wbox floats]
// X3: the error-compensated product needs dout_hi in_hi + dout_hi in_lo + dout_lo in_hi.  The residual twins already
// exist in global memory (written by the kernels that produced dout / in), so the three terms are three
// load -> MMA rounds through the SAME operand buffers: no second copy in shared memory, two CTAs per SM as before.
// Both weight matrices are updated by ONE launch: blockIdx.x < nx1 are W1 tiles, the rest W2 tiles (fewer, larger
// waves than two launches, and one kernel boundary less on the step's critical path).
struct AdamMaps { CUtensorMap A, B, Alo, Blo, W, M, V; };   // dout, in, their residual twins, and the w / m / v tiles

template <bool X3>
__global__ void __launch_bounds__(NTHREADS, 2) tc_adam_kernel(const __grid_constant__ AdamMaps maps1,
                                                              const __grid_constant__ AdamMaps maps2, const TcParams p) {
    const int s = blockIdx.z + p.s_base;
    const SubnetDesc d = p.desc[s];
    const bool second = (int)blockIdx.x >= p.nx1;
    const AdamMaps* mp = second ? &maps2 : &maps1;
    const CUtensorMap &mapA = mp->A, &mapB = mp->B, &mapAlo = mp->Alo, &mapBlo = mp->Blo, &mapW = mp->W, &mapM = mp->M, &mapV = mp->V;
    const int m0 = blockIdx.y * TILE_M;
    const int n0 = ((int)blockIdx.x - (second ? p.nx1 : 0)) * p.n_cols;
    int out_dim, in_dim, a_c0, b_c0, b_c1;
    int64_t row_base;                                     // first weight row of this sub-network
    if (!second) { out_dim = p.Hp; in_dim = d.Pp; a_c0 = s * p.Hp + m0; b_c0 = (int)d.coff + n0; b_c1 = (int)p.row0; row_base = d.coff; }
    else { out_dim = p.Op; in_dim = p.Hp; a_c0 = s * p.Op + m0; b_c0 = s * p.Hp + n0; b_c1 = 0; row_base = (int64_t)s * p.Hp; }
    if (m0 >= out_dim || n0 >= in_dim) return;
    const int nkb = p.nkb_adam;
    const int ncols_ok = min(p.n_cols, in_dim - n0);      // multiple of 32: chunks of AD_R rows are whole
    const int nchunks = ncols_ok / AD_R;

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t b_block_bytes = (uint32_t)p.n_cols * BLOCK_K * 4;
    const int KG = p.ad_kg;                               // K blocks per pass
    const int ppr = (nkb + KG - 1) / KG;                  // passes per compensation round
    uint8_t* sA = smem;
    uint8_t* sB = smem + (size_t)KG * A_STAGE_BYTES;
    float* wring = reinterpret_cast<float*>(sB + (size_t)KG * b_block_bytes);
    const int wbox = second ? p.wbox2 : p.wbox;
    const int tile_floats = AD_R * wbox;                  // one tensor, one chunk
    const uint32_t chunk_bytes = 3u * tile_floats * 4u;
    // ring stages: AD_STAGES dedicated ones, then as many as fit in the operand buffers, which are dead once the
    // accumulator is complete -- the deeper ring is what keeps enough bytes in flight to cover HBM latency
    const uint32_t ops_bytes = (uint32_t)KG * (A_STAGE_BYTES + b_block_bytes);
    const int ring = min(AD_MAX_RING, AD_STAGES + (int)(ops_bytes / chunk_bytes));
    auto stage_ptr = [&](int st) -> float* {
        return st < AD_STAGES ? wring + (size_t)st * 3 * tile_floats
                              : reinterpret_cast<float*>(smem + (size_t)(st - AD_STAGES) * chunk_bytes);
    };
    __shared__ uint64_t ops_bar, mma_bar, tmem_full_bar, wfull[AD_MAX_RING], wdone[AD_MAX_RING];
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    DI_TRACE_T0(0);
    if (threadIdx.x == 0) {
        mbar_init(&ops_bar, 1); mbar_init(&mma_bar, 1); mbar_init(&tmem_full_bar, 1);
        for (int i = 0; i < ring; ++i) { mbar_init(&wfull[i], 1); mbar_init(&wdone[i], 128); }
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&tmem_base_slot, (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_slot;
    pdl_wait();
    if (p.pdl_early && threadIdx.x == 0) pdl_release();
    __syncwarp();

    auto load_chunk = [&](int c) {                        // one elected thread
        const int st = c % ring;
        float* ws = stage_ptr(st);
        const int32_t r = (int32_t)(row_base + n0 + c * AD_R);
        mbar_arrive_expect_tx(&wfull[st], chunk_bytes);
        tma_load_2d(ws, &mapW, &wfull[st], m0, r);
        tma_load_2d(ws + tile_floats, &mapM, &wfull[st], m0, r);
        tma_load_2d(ws + 2 * tile_floats, &mapV, &wfull[st], m0, r);
    };

    if (warp == 4) {
        // ===== TMA: operands, then the w/m/v ring (loads ahead of the epilogue, stores behind it) =====
        if (elect_one()) {
            // passes: X3 rounds (dout_hi in_lo, dout_hi in_hi, dout_lo in_hi) x groups of KG K blocks; a pass reloads
            // the operand buffers once the MMAs of the previous pass have read them (plain TF32: one round, Blo == B)
            constexpr int ROUNDS = X3 ? 3 : 1;
            for (int pi = 0; pi < ROUNDS * ppr; ++pi) {
                const int round = pi / ppr, kb0 = (pi % ppr) * KG, cnt = min(KG, nkb - kb0);
                if (pi > 0) mbar_wait(&mma_bar, (pi - 1) & 1, 9);
                mbar_arrive_expect_tx(&ops_bar, (uint32_t)cnt * (A_STAGE_BYTES + b_block_bytes));
                const CUtensorMap* ma = (round == 2) ? &mapAlo : &mapA;
                const CUtensorMap* mb = (X3 && round == 0) ? &mapBlo : &mapB;
                for (int k = 0; k < cnt; ++k) {
                    load_stage<true>(sA + (size_t)k * A_STAGE_BYTES, ma, &ops_bar, a_c0, (kb0 + k) * BLOCK_K, TILE_M);
                    load_stage<true>(sB + (size_t)k * b_block_bytes, mb, &ops_bar, b_c0, b_c1 + (kb0 + k) * BLOCK_K, p.n_cols);
                }
                if (pi == 0) for (int c = 0; c < min(AD_STAGES, nchunks); ++c) load_chunk(c);
            }
            if (ring > AD_STAGES) {                                // the operand buffers join the ring
                mbar_wait(&tmem_full_bar, 0, 4);
                for (int c = AD_STAGES; c < min(ring, nchunks); ++c) load_chunk(c);
            }
            if (p.adam_direct) {
                // the epilogue threads store their results to global memory themselves: a stage is free again as soon
                // as all 128 of them hold its chunk in registers
                for (int c = 0; c + ring < nchunks; ++c) {
                    mbar_wait(&wdone[c % ring], (c / ring) & 1, 8);
                    load_chunk(c + ring);
                }
            } else {
            for (int c = 0; c < nchunks; ++c) {
                const int st = c % ring;
                float* ws = stage_ptr(st);
                mbar_wait(&wdone[st], (c / ring) & 1, 8);           // all 128 epilogue threads updated this chunk
                if (c < 40) DI_TRACE(48 + c);
                const int32_t r = (int32_t)(row_base + n0 + c * AD_R);
                tma_store_2d(&mapW, ws, m0, r);
                tma_store_2d(&mapM, ws + tile_floats, m0, r);
                tma_store_2d(&mapV, ws + 2 * tile_floats, m0, r);
                bulk_commit();
                // refill one iteration late: by now the PREVIOUS chunk's stores have read their stage, so the wait
                // does not stall this thread (it only has to keep the most recent group in flight)
                if (c >= 1 && c - 1 + ring < nchunks) {
                    bulk_wait_read<1>();
                    load_chunk(c - 1 + ring);
                }
            }
            }
            bulk_wait<0>();
        }
    } else if (warp == 5) {
        if (elect_one()) {
            const uint32_t idesc = idesc_for(p.n_cols, true, true);
            constexpr int ROUNDS = X3 ? 3 : 1;
            const int npass = ROUNDS * ppr;
            for (int pi = 0; pi < npass; ++pi) {
                const int cnt = min(KG, nkb - (pi % ppr) * KG);
                mbar_wait(&ops_bar, pi & 1, 6);
                tc_fence_after();
                // X3: round 1 (dout_hi in_hi) accumulates in columns [0, n_cols), rounds 0 and 2 (the small products) in
                // [n_cols, 2 n_cols): short chains, see acc_sum16
                const int round = pi / ppr;
                const uint32_t dst = (X3 && p.lo_acc && round != 1) ? tmem + (uint32_t)p.n_cols : tmem;
                const bool fresh = (pi % ppr) == 0 && (round == 0 || (X3 && p.lo_acc && round == 1));   // first pass of this accumulator
                for (int k = 0; k < cnt; ++k) {
                    const uint32_t sa = smem_u32(sA + (size_t)k * A_STAGE_BYTES);
                    const uint32_t sb = smem_u32(sB + (size_t)k * b_block_bytes);
#pragma unroll
                    for (int j = 0; j < BLOCK_K / UMMA_K; ++j)
                        umma_tf32(dst, stage_desc<true>(sa, j), stage_desc<true>(sb, j), idesc, (!fresh || k || j) ? 1u : 0u);
                }
                if (pi + 1 < npass) umma_commit(&mma_bar);
            }
            umma_commit(&tmem_full_bar);
        }
    } else {
        const int fl = warp * 32 + lane;
        const bool f_ok = (m0 + fl) < out_dim;
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
        const AdamParams adam = adam_of(p);
        DI_TRACE_T0(2);
        mbar_wait(&tmem_full_bar, 0, 4);
        tc_fence_after();
        if (threadIdx.x == 0) pdl_release();
        __syncwarp();
        DI_TRACE_T0(3);
        for (int c = 0; c < nchunks; ++c) {
            const int st = c % ring;
            float* ws = stage_ptr(st);
            float g[AD_R];
            __syncwarp();
            tmem_ld8(taddr + c * AD_R, g);
            if (X3 && p.lo_acc) {
                float gl[AD_R];
                tmem_ld8(taddr + p.n_cols + c * AD_R, gl);
#pragma unroll
                for (int r = 0; r < AD_R; ++r) g[r] += gl[r];
            }
            mbar_wait(&wfull[st], (c / ring) & 1, 7);
            if (c < 40) DI_TRACE_T0(8 + c);
            if (f_ok) {
                // all 24 loads, then 8 independent updates, then all 24 stores: written with shared-space
                // instructions on explicit register arrays so that no store can alias (and serialise) a later load
                const uint32_t base = smem_u32(ws) + (uint32_t)fl * 4u;
                const uint32_t row_b = (uint32_t)wbox * 4u, tile_b = (uint32_t)tile_floats * 4u;
                float w[AD_R], m[AD_R], v[AD_R];
#pragma unroll
                for (int r = 0; r < AD_R; ++r) {
                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(w[r]) : "r"(base + r * row_b));
                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(m[r]) : "r"(base + tile_b + r * row_b));
                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v[r]) : "r"(base + 2 * tile_b + r * row_b));
                }
#pragma unroll
                for (int r = 0; r < AD_R; ++r) adam_update_fast(g[r], w[r], m[r], v[r], adam);
                if (p.adam_direct) {
                    // the updates consumed every loaded value, so the stage can be refilled; results leave from
                    // registers: a warp writes 128 contiguous bytes per instruction, 24 independent stores in flight
                    mbar_arrive(&wdone[st]);
                    const int64_t off = (row_base + n0 + (int64_t)c * AD_R) * out_dim + m0 + fl;
                    float* gw = (second ? p.W2 : p.W1) + off;
                    float* gm = (second ? p.mW2 : p.mW1) + off;
                    float* gv = (second ? p.vW2 : p.vW1) + off;
#pragma unroll
                    for (int r = 0; r < AD_R; ++r) {
                        gw[(int64_t)r * out_dim] = w[r];
                        gm[(int64_t)r * out_dim] = m[r];
                        gv[(int64_t)r * out_dim] = v[r];
                    }
                    if (float* gl0 = second ? p.W2lo : p.W1lo) {
                        float* gl = gl0 + off;
#pragma unroll
                        for (int r = 0; r < AD_R; ++r) gl[(int64_t)r * out_dim] = tf32_residual(w[r]);
                    }
                    if (c < 40 && threadIdx.x == 0) DI_TRACE(48 + c);
                    continue;
                }
                if (float* gl0 = second ? p.W2lo : p.W1lo) {
                    float* gl = gl0 + (row_base + n0 + (int64_t)c * AD_R) * out_dim + m0 + fl;
#pragma unroll
                    for (int r = 0; r < AD_R; ++r) gl[(int64_t)r * out_dim] = tf32_residual(w[r]);
                }
#pragma unroll
                for (int r = 0; r < AD_R; ++r) {
                    asm volatile("st.shared.f32 [%0], %1;" ::"r"(base + r * row_b), "f"(w[r]) : "memory");
                    asm volatile("st.shared.f32 [%0], %1;" ::"r"(base + tile_b + r * row_b), "f"(m[r]) : "memory");
                    asm volatile("st.shared.f32 [%0], %1;" ::"r"(base + 2 * tile_b + r * row_b), "f"(v[r]) : "memory");
                }
            } else if (p.adam_direct) {
                mbar_arrive(&wdone[st]);                  // padding feature: nothing to update, the stage still needs 128 arrivals
                continue;
            }
            fence_proxy_async();                          // generic-proxy writes -> visible to the TMA store
            mbar_arrive(&wdone[st]);
        }
        DI_TRACE_T0(4);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, (uint32_t)p.tmem_cols);
    DI_TRACE_T0(5);
}

// One [AD_R rows] x {w, m, v} chunk of one feature: shared memory -> registers -> Adam -> global memory.
// WBOX / LD > 0: row pitch of the shared tile (floats) and of the weight matrix known at compile time.
template <int WBOX, int LD>
__device__ __forceinline__ void adam_chunk(uint32_t base, const float (&g)[AD_R], const AdamParams& adam,
                                           float* __restrict__ gw, float* __restrict__ gm, float* __restrict__ gv,
                                           float* __restrict__ gl, int wbox_rt, int ld_rt) {
    const uint32_t row_b = (WBOX > 0 ? (uint32_t)WBOX : (uint32_t)wbox_rt) * 4u;
    const uint32_t tile_b = row_b * AD_R;
    const int64_t ld = LD > 0 ? (int64_t)LD : (int64_t)ld_rt;
    float w[AD_R], m[AD_R], v[AD_R];
#pragma unroll
    for (int r = 0; r < AD_R; ++r) {
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(w[r]) : "r"(base + r * row_b));
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(m[r]) : "r"(base + tile_b + r * row_b));
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v[r]) : "r"(base + 2 * tile_b + r * row_b));
    }
#pragma unroll
    for (int r = 0; r < AD_R; ++r) adam_update_fast(g[r], w[r], m[r], v[r], adam);
#pragma unroll
    for (int r = 0; r < AD_R; ++r) {
        gw[r * ld] = w[r];
        gm[r * ld] = m[r];
        gv[r * ld] = v[r];
    }
    if (gl) {                                             // residual twin of the new weights (operand of the LT kernels)
#pragma unroll
        for (int r = 0; r < AD_R; ++r) gl[r * ld] = tf32_residual(w[r]);
    }
}

// ------------------------------------------------------------------------------------------ ADAM, one CTA per SM
// Same tile and arithmetic as tc_adam_kernel, laid out for ONE resident CTA per SM that never waits on a buffer:
//   * every operand set is resident at once (X3: dout_hi, in_lo, in_hi, dout_lo), so the three compensation rounds
//     are issued back to back -- the ring kernel reloads two of its buffers between rounds and spends a third of its
//     life (about 10k of 31k cycles, in-kernel trace) waiting for TMA before the accumulator is complete;
//   * every [8 rows x 128 features] x {w, m, v} chunk of the tile owns a shared-memory stage (n_ded dedicated ones
//     are filled while the MMAs run, the rest reuse the operand area once the accumulator is complete), so there is no
//     refill protocol at all;
//   * 4 x G epilogue warps (G per TMEM lane quadrant, G = 2..4 chosen at launch) take chunks round-robin and write w, m, v to global memory
//     straight from registers: the update of one chunk is a latency chain (TMEM load, 24 shared loads, 16 MUFU ops,
//     24 stores) of about 900 cycles per warp, so the number of warps sets the pace.
// (4 G + 2) warps: 0 .. 4G-1 epilogue, then the TMA warp, then the MMA warp.  Needs 4 (X3) or 2 operand sets of nkb x 16 KB each:
// used when that fits (batch <= 64 in X3 mode), otherwise the ring kernel above runs.
constexpr int AD_MAX_GROUPS = 4;                          // epilogue warp groups (4 warps each) taking chunks round-robin
constexpr int NTHREADS_BIG = (4 * AD_MAX_GROUPS + 2) * 32;   // + TMA warp + MMA warp; the launch may use fewer groups
constexpr int AD_MAX_CHUNKS = ADAM_TILE / AD_R;

template <bool X3>
__global__ void __launch_bounds__(NTHREADS_BIG, 1) tc_adam_big_kernel(const __grid_constant__ AdamMaps maps1,
                                                                       const __grid_constant__ AdamMaps maps2, const TcParams p) {
    const int s = blockIdx.z + p.s_base;
    const SubnetDesc d = p.desc[s];
    const bool second = (int)blockIdx.x >= p.nx1;
    const AdamMaps* mp = second ? &maps2 : &maps1;
    const CUtensorMap &mapA = mp->A, &mapB = mp->B, &mapAlo = mp->Alo, &mapBlo = mp->Blo, &mapW = mp->W, &mapM = mp->M, &mapV = mp->V;
    const int m0 = blockIdx.y * TILE_M;
    const int n0 = ((int)blockIdx.x - (second ? p.nx1 : 0)) * p.n_cols;
    int out_dim, in_dim, a_c0, b_c0, b_c1;
    int64_t row_base;
    if (!second) { out_dim = p.Hp; in_dim = d.Pp; a_c0 = s * p.Hp + m0; b_c0 = (int)d.coff + n0; b_c1 = (int)p.row0; row_base = d.coff; }
    else { out_dim = p.Op; in_dim = p.Hp; a_c0 = s * p.Op + m0; b_c0 = s * p.Hp + n0; b_c1 = 0; row_base = (int64_t)s * p.Hp; }
    if (m0 >= out_dim || n0 >= in_dim) return;
    const int nkb = p.nkb_adam;
    const int nchunks = min(p.n_cols, in_dim - n0) / AD_R;

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    constexpr int NSETS = X3 ? 4 : 2;
    const uint32_t set_bytes = (uint32_t)nkb * A_STAGE_BYTES;       // n_cols == TILE_M: A and B sets have one size
    uint8_t* set0 = smem;                                           // dout_hi
    uint8_t* set1 = smem + set_bytes;                               // in_lo   (plain TF32: in)
    uint8_t* set2 = smem + 2 * (size_t)set_bytes;                   // in_hi
    uint8_t* set3 = smem + 3 * (size_t)set_bytes;                   // dout_lo
    uint8_t* ded = smem + (size_t)NSETS * set_bytes;                // dedicated chunk stages
    const int wbox = second ? p.wbox2 : p.wbox;
    const int tile_floats = AD_R * wbox;
    const uint32_t chunk_bytes = 3u * tile_floats * 4u;
    auto stage_ptr = [&](int c) -> float* {
        return reinterpret_cast<float*>(c < p.ad_nded ? ded + (size_t)c * p.ad_stride : smem + (size_t)(c - p.ad_nded) * p.ad_stride);
    };
    __shared__ uint64_t ops_bar[2], tmem_full_bar, wfull[AD_MAX_CHUNKS];
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ngroups = (((int)blockDim.x >> 5) - 2) >> 2;
    DI_TRACE_T0(0);
    if (threadIdx.x == 0) {
        mbar_init(&ops_bar[0], 1); mbar_init(&ops_bar[1], 1); mbar_init(&tmem_full_bar, 1);
        for (int i = 0; i < AD_MAX_CHUNKS; ++i) mbar_init(&wfull[i], 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&tmem_base_slot, (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_slot;
    if (warp != 4 * ngroups) pdl_wait();                   // the producer warp waits later (see below)
    if (p.pdl_early && threadIdx.x == 0) pdl_release();
    __syncwarp();

    if (warp == 4 * ngroups) {
        if (elect_one()) {
            auto load_a = [&](const CUtensorMap* m, uint8_t* dst, uint64_t* bar) {
                for (int kb = 0; kb < nkb; ++kb)
                    load_stage<true>(dst + (size_t)kb * A_STAGE_BYTES, m, bar, a_c0, kb * BLOCK_K, TILE_M);
            };
            auto load_b = [&](const CUtensorMap* m, uint8_t* dst, uint64_t* bar) {
                for (int kb = 0; kb < nkb; ++kb)
                    load_stage<true>(dst + (size_t)kb * A_STAGE_BYTES, m, bar, b_c0, b_c1 + kb * BLOCK_K, p.n_cols);
            };
            auto load_chunk = [&](int c) {
                float* ws = stage_ptr(c);
                const int32_t r = (int32_t)(row_base + n0 + c * AD_R);
                mbar_arrive_expect_tx(&wfull[c], chunk_bytes);
                tma_load_2d(ws, &mapW, &wfull[c], m0, r);
                tma_load_2d(ws + tile_floats, &mapM, &wfull[c], m0, r);
                tma_load_2d(ws + 2 * tile_floats, &mapV, &wfull[c], m0, r);
            };
            // ADAM follows BWD, which writes dz1 (the `dout` of the W1 tiles) and nothing else this kernel reads: the
            // w / m / v chunks, the `in` operands and -- for W2 tiles -- dout = dz2 are fetched before waiting for it
            const bool a_free = p.pdl_prefetch && second;
            const bool early = p.pdl_prefetch != 0;
            if constexpr (X3) {
                mbar_arrive_expect_tx(&ops_bar[0], 3u * set_bytes);
                mbar_arrive_expect_tx(&ops_bar[1], set_bytes);
            } else {
                mbar_arrive_expect_tx(&ops_bar[0], 2u * set_bytes);
            }
            auto load_in = [&]() {
                if constexpr (X3) { load_b(&mapBlo, set1, &ops_bar[0]); load_b(&mapB, set2, &ops_bar[0]); }
                else load_b(&mapB, set1, &ops_bar[0]);
            };
            auto load_dout = [&]() {
                load_a(&mapA, set0, &ops_bar[0]);
                if constexpr (X3) load_a(&mapAlo, set3, &ops_bar[1]);
            };
            if (early) {
                load_in();
                if (a_free) load_dout();
                for (int c = 0; c < min(p.ad_nded, nchunks); ++c) load_chunk(c);
                pdl_wait();
                if (!a_free) load_dout();
            } else {
                pdl_wait();
                load_dout(); load_in();
                for (int c = 0; c < min(p.ad_nded, nchunks); ++c) load_chunk(c);
            }
            if (nchunks > p.ad_nded) {                             // the operand area becomes chunk stages
                mbar_wait(&tmem_full_bar, 0, 4);
                for (int c = p.ad_nded; c < nchunks; ++c) load_chunk(c);
            }
        }
    } else if (warp == 4 * ngroups + 1) {
        if (elect_one()) {
            const uint32_t idesc = idesc_for(p.n_cols, true, true);
            auto mma_round = [&](uint32_t dst, const uint8_t* a, const uint8_t* b, bool first) {
                for (int kb = 0; kb < nkb; ++kb) {
                    const uint32_t sa = smem_u32(a + (size_t)kb * A_STAGE_BYTES);
                    const uint32_t sb = smem_u32(b + (size_t)kb * A_STAGE_BYTES);
#pragma unroll
                    for (int j = 0; j < BLOCK_K / UMMA_K; ++j)
                        umma_tf32(dst, stage_desc<true>(sa, j), stage_desc<true>(sb, j), idesc, (!first || kb || j) ? 1u : 0u);
                }
            };
            // X3: the hi hi product and the two small products accumulate separately (columns [0, 128) and [128, 256),
            // see acc_sum16); the epilogue adds the two tiles
            const uint32_t d_lo = p.lo_acc ? tmem + (uint32_t)ADAM_TILE : tmem;
            mbar_wait(&ops_bar[0], 0, 6);
            tc_fence_after();
            if constexpr (X3) {
                mma_round(d_lo, set0, set1, true);                     // dout_hi in_lo
                mma_round(tmem, set0, set2, p.lo_acc != 0);            // dout_hi in_hi
                mbar_wait(&ops_bar[1], 0, 6);
                tc_fence_after();
                mma_round(d_lo, set3, set2, false);                    // dout_lo in_hi
            } else {
                mma_round(tmem, set0, set1, true);                     // dout in
            }
            umma_commit(&tmem_full_bar);
        }
    } else {
        const int quad = warp & 3, grp = warp >> 2;       // TMEM lane quadrant of this warp; chunks c = grp, grp + ngroups, ...
        const int fl = quad * 32 + lane;
        const bool f_ok = (m0 + fl) < out_dim;
        const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16);
        const AdamParams adam = adam_of(p);
        const bool tracer = p.trace && quad == 0 && lane == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
        DI_TRACE_T0(2);
        mbar_wait(&tmem_full_bar, 0, 4);
        tc_fence_after();
        if (threadIdx.x == 0) pdl_release();
        __syncwarp();
        DI_TRACE_T0(3);
        float* gw0 = second ? p.W2 : p.W1;
        float* gm0 = second ? p.mW2 : p.mW1;
        float* gv0 = second ? p.vW2 : p.vW1;
        float* gl0 = second ? p.W2lo : p.W1lo;
        for (int c = grp; c < nchunks; c += ngroups) {
            float g[AD_R];
            __syncwarp();
            tmem_ld8(taddr + c * AD_R, g);
            if (X3 && p.lo_acc) {
                float gl[AD_R];
                tmem_ld8(taddr + ADAM_TILE + c * AD_R, gl);
#pragma unroll
                for (int r = 0; r < AD_R; ++r) g[r] += gl[r];
            }
            mbar_wait(&wfull[c], 0, 7);
            if (tracer && c < 40) p.trace[8 + c] = clock64();
            __syncwarp();
            if (f_ok) {
                const uint32_t base = smem_u32(stage_ptr(c)) + (uint32_t)fl * 4u;
                const int64_t off = (row_base + n0 + (int64_t)c * AD_R) * out_dim + m0 + fl;
                // the epilogue is instruction-bound: with the row pitch known at compile time (the default topology:
                // 128-wide tiles of W1 [.., 256] and W2 [.., 512]) every shared load and global store addresses
                // base + immediate instead of computing 48 addresses per chunk
                float* gl = gl0 ? gl0 + off : nullptr;
                if (p.ad_generic) adam_chunk<0, 0>(base, g, adam, gw0 + off, gm0 + off, gv0 + off, gl, wbox, out_dim);
                else if (wbox == TILE_M && out_dim == 256) adam_chunk<TILE_M, 256>(base, g, adam, gw0 + off, gm0 + off, gv0 + off, gl, 0, 0);
                else if (wbox == TILE_M && out_dim == 512) adam_chunk<TILE_M, 512>(base, g, adam, gw0 + off, gm0 + off, gv0 + off, gl, 0, 0);
                else adam_chunk<0, 0>(base, g, adam, gw0 + off, gm0 + off, gv0 + off, gl, wbox, out_dim);
            }
            if (tracer && c < 40) p.trace[48 + c] = clock64();
        }
        __syncwarp();
        DI_TRACE_T0(4);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, (uint32_t)p.tmem_cols);
    DI_TRACE_T0(5);
}

// ------------------------------------------------------------------------------ ADAM, persistent: tiles overlap
// tc_adam_big_kernel spends the first 45 % of a tile's life (operand loads over a shared L2, then the MMAs) with the
// w / m / v stream idle, and the rest with the tensor core idle.  Here one CTA walks several tiles
// (t = blockIdx.x, + gridDim.x, ...) and overlaps the two halves: tensor memory holds two accumulator stages, so while
// the sixteen epilogue warps stream tile i's chunks through an 8-stage ring (load by TMA, update in registers, store),
// the TMA warp already fetches tile i + 1's operand sets into the operand area -- free as soon as tile i's MMAs have
// been committed -- and the MMA warp fills the other stage.  Same tile, operands, arithmetic and summation order as the
// resident kernel: results are bit-identical.  Used where throughput counts (many sub-networks per GPU); the resident
// kernel, one tile per CTA, stays for the latency-bound regime.
constexpr int ADP_RING = 8;                               // chunk stages

template <bool X3>
__global__ void __launch_bounds__(NTHREADS_BIG, 1) tc_adam_pers_kernel(const __grid_constant__ AdamMaps maps1,
                                                                        const __grid_constant__ AdamMaps maps2, const TcParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    constexpr int NSETS = X3 ? 4 : 2;
    const int nkb = p.nkb_adam;
    const uint32_t set_bytes = (uint32_t)nkb * A_STAGE_BYTES;
    uint8_t* set0 = smem;                                           // dout_hi
    uint8_t* set1 = smem + set_bytes;                               // in_lo   (plain TF32: in)
    uint8_t* set2 = smem + 2 * (size_t)set_bytes;                   // in_hi
    uint8_t* set3 = smem + 3 * (size_t)set_bytes;                   // dout_lo
    uint8_t* ring = smem + (size_t)NSETS * set_bytes;
    __shared__ uint64_t ops_full[2], ops_empty, tmem_full_bar[2], tmem_empty_bar[2], wfull[ADP_RING], wdone[ADP_RING];
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int TMA_WARP = 4 * AD_MAX_GROUPS, MMA_WARP = TMA_WARP + 1;
    if (threadIdx.x == 0) {
        mbar_init(&ops_full[0], 1); mbar_init(&ops_full[1], 1); mbar_init(&ops_empty, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full_bar[i], 1); mbar_init(&tmem_empty_bar[i], 4 * AD_MAX_GROUPS); }
        for (int i = 0; i < ADP_RING; ++i) { mbar_init(&wfull[i], 1); mbar_init(&wdone[i], 4); }
        fence_barrier_init();
    }
    const uint32_t acc_cols = (uint32_t)ADAM_TILE * (p.lo_acc ? 2u : 1u);     // columns of one accumulator stage
    if (warp == 0) tmem_alloc(&tmem_base_slot, (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_slot;
    if (warp != TMA_WARP) pdl_wait();
    if (p.pdl_early && threadIdx.x == 0) pdl_release();
    __syncwarp();

    // tile t of this launch -> sub-network, weight matrix, tile origin.  tiles per sub-network: nx1 * mh of W1, then nx2 * mo of W2
    const int mh = (p.Hp + TILE_M - 1) / TILE_M, mo = (p.Op + TILE_M - 1) / TILE_M, nx2 = (p.Hp + ADAM_TILE - 1) / ADAM_TILE;
    const int tps = p.nx1 * mh + nx2 * mo;
    const int n_tiles = p.row_tiles * tps;                // row_tiles: sub-networks of this launch
    struct Tile { int s, m0, n0, out_dim, a_c0, b_c0, b_c1, nch, wbox; bool second; int64_t row_base; };
    auto tile_of = [&](int t, Tile& T) {
        T.s = t / tps + p.s_base;
        int r = t % tps;
        T.second = r >= p.nx1 * mh;
        const SubnetDesc d = p.desc[T.s];
        int in_dim;
        if (!T.second) {
            T.n0 = (r / mh) * ADAM_TILE; T.m0 = (r % mh) * TILE_M; T.out_dim = p.Hp; in_dim = d.Pp;
            T.a_c0 = T.s * p.Hp + T.m0; T.b_c0 = (int)d.coff + T.n0; T.b_c1 = (int)p.row0; T.row_base = d.coff; T.wbox = p.wbox;
        } else {
            r -= p.nx1 * mh;
            T.n0 = (r / mo) * ADAM_TILE; T.m0 = (r % mo) * TILE_M; T.out_dim = p.Op; in_dim = p.Hp;
            T.a_c0 = T.s * p.Op + T.m0; T.b_c0 = T.s * p.Hp + T.n0; T.b_c1 = 0; T.row_base = (int64_t)T.s * p.Hp; T.wbox = p.wbox2;
        }
        T.nch = (T.m0 < T.out_dim && T.n0 < in_dim) ? min(ADAM_TILE, in_dim - T.n0) / AD_R : 0;    // 0: nothing to do (ragged edge)
    };

    if (warp == TMA_WARP) {
        if (elect_one()) {
            auto load_ops = [&](const Tile& T, bool first_tile) {
                const AdamMaps* mp = T.second ? &maps2 : &maps1;
                auto load_a = [&](const CUtensorMap* m, uint8_t* dst, uint64_t* bar) {
                    for (int kb = 0; kb < nkb; ++kb) load_stage<true>(dst + (size_t)kb * A_STAGE_BYTES, m, bar, T.a_c0, kb * BLOCK_K, TILE_M);
                };
                auto load_b = [&](const CUtensorMap* m, uint8_t* dst, uint64_t* bar) {
                    for (int kb = 0; kb < nkb; ++kb) load_stage<true>(dst + (size_t)kb * A_STAGE_BYTES, m, bar, T.b_c0, T.b_c1 + kb * BLOCK_K, ADAM_TILE);
                };
                if constexpr (X3) { mbar_arrive_expect_tx(&ops_full[0], 3u * set_bytes); mbar_arrive_expect_tx(&ops_full[1], set_bytes); }
                else mbar_arrive_expect_tx(&ops_full[0], 2u * set_bytes);
                // (the launch's first tile: ADAM follows BWD, which writes dz1 -- the dout of W1 tiles -- and nothing else
                // read here, so everything but that is requested before waiting for it)
                if constexpr (X3) { load_b(&mp->Blo, set1, &ops_full[0]); load_b(&mp->B, set2, &ops_full[0]); }
                else load_b(&mp->B, set1, &ops_full[0]);
                const bool wait_first = first_tile && !(p.pdl_prefetch && T.second);
                if (first_tile && !wait_first) { /* dout = dz2: written two grids back */ }
                if (wait_first) pdl_wait();
                load_a(&mp->A, set0, &ops_full[0]);
                if constexpr (X3) load_a(&mp->Alo, set3, &ops_full[1]);
                if (first_tile && !wait_first) pdl_wait();
            };
            int g = 0, i = 0;                             // chunks issued so far (ring position), non-empty tiles so far
            Tile T, Tn;
            int t = blockIdx.x;
            for (; t < n_tiles; t += gridDim.x) { tile_of(t, T); if (T.nch) break; }
            if (t < n_tiles) load_ops(T, true); else pdl_wait();
            while (t < n_tiles) {
                int tn = t + gridDim.x;                   // next non-empty tile of this CTA
                for (; tn < n_tiles; tn += gridDim.x) { tile_of(tn, Tn); if (Tn.nch) break; }
                const AdamMaps* mp = T.second ? &maps2 : &maps1;
                const int tile_floats = AD_R * T.wbox;
                auto load_chunk = [&](int c) {
                    const int slot = g % ADP_RING;
                    if (g >= ADP_RING) mbar_wait(&wdone[slot], ((g / ADP_RING) - 1) & 1, 8);
                    float* ws = reinterpret_cast<float*>(ring + (size_t)slot * p.ad_stride);
                    const int32_t r = (int32_t)(T.row_base + T.n0 + c * AD_R);
                    mbar_arrive_expect_tx(&wfull[slot], 3u * (uint32_t)tile_floats * 4u);
                    tma_load_2d(ws, &mp->W, &wfull[slot], T.m0, r);
                    tma_load_2d(ws + tile_floats, &mp->M, &wfull[slot], T.m0, r);
                    tma_load_2d(ws + 2 * tile_floats, &mp->V, &wfull[slot], T.m0, r);
                    ++g;
                };
                int c = 0;
                for (; c < min(ADP_RING, T.nch); ++c) load_chunk(c);
                // this tile's MMAs have read the operand area: the next tile's operands may land
                mbar_wait(&ops_empty, i & 1, 9);
                if (tn < n_tiles) load_ops(Tn, false);
                for (; c < T.nch; ++c) load_chunk(c);
                t = tn; T = Tn; ++i;
            }
        }
    } else if (warp == MMA_WARP) {
        if (elect_one()) {
            const uint32_t idesc = idesc_for(ADAM_TILE, true, true);
            int i = 0;
            Tile T;
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                tile_of(t, T);
                if (!T.nch) continue;
                const int acc = i & 1;
                mbar_wait(&tmem_empty_bar[acc], ((i >> 1) & 1) ^ 1, 12);
                tc_fence_after();
                const uint32_t d_hi = tmem + (uint32_t)acc * acc_cols, d_lo = p.lo_acc ? d_hi + ADAM_TILE : d_hi;
                auto mma_round = [&](uint32_t dst, const uint8_t* a, const uint8_t* b, bool first) {
                    for (int kb = 0; kb < nkb; ++kb) {
                        const uint32_t sa = smem_u32(a + (size_t)kb * A_STAGE_BYTES);
                        const uint32_t sb = smem_u32(b + (size_t)kb * A_STAGE_BYTES);
#pragma unroll
                        for (int j = 0; j < BLOCK_K / UMMA_K; ++j)
                            umma_tf32(dst, stage_desc<true>(sa, j), stage_desc<true>(sb, j), idesc, (!first || kb || j) ? 1u : 0u);
                    }
                };
                mbar_wait(&ops_full[0], i & 1, 6);
                tc_fence_after();
                if constexpr (X3) {
                    mma_round(d_lo, set0, set1, true);                     // dout_hi in_lo
                    mma_round(d_hi, set0, set2, p.lo_acc != 0);            // dout_hi in_hi
                    mbar_wait(&ops_full[1], i & 1, 6);
                    tc_fence_after();
                    mma_round(d_lo, set3, set2, false);                    // dout_lo in_hi
                } else {
                    mma_round(d_hi, set0, set1, true);
                }
                umma_commit(&ops_empty);
                umma_commit(&tmem_full_bar[acc]);
                ++i;
            }
        }
    } else {
        const int quad = warp & 3, grp = warp >> 2;
        const int fl = quad * 32 + lane;
        const AdamParams adam = adam_of(p);
        int g0 = 0, i = 0, n_live = 0;
        Tile T;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) { tile_of(t, T); n_live += T.nch ? 1 : 0; }
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            tile_of(t, T);
            if (!T.nch) continue;
            const int acc = i & 1;
            const bool f_ok = (T.m0 + fl) < T.out_dim;
            const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)acc * acc_cols;
            mbar_wait(&tmem_full_bar[acc], (i >> 1) & 1, 4);
            tc_fence_after();
            if (i + 1 == n_live && threadIdx.x == 0 && !p.pdl_early) pdl_release();     // the last accumulator of this CTA is complete
            __syncwarp();
            float* gw0 = T.second ? p.W2 : p.W1;
            float* gm0 = T.second ? p.mW2 : p.mW1;
            float* gv0 = T.second ? p.vW2 : p.vW1;
            float* gl0 = T.second ? p.W2lo : p.W1lo;
            for (int c = grp; c < T.nch; c += AD_MAX_GROUPS) {
                const int g = g0 + c, slot = g % ADP_RING;
                float gr[AD_R];
                __syncwarp();
                tmem_ld8(taddr + c * AD_R, gr);
                if (X3 && p.lo_acc) {
                    float gl[AD_R];
                    tmem_ld8(taddr + ADAM_TILE + c * AD_R, gl);
#pragma unroll
                    for (int r = 0; r < AD_R; ++r) gr[r] += gl[r];
                }
                mbar_wait(&wfull[slot], (g / ADP_RING) & 1, 7);
                __syncwarp();
                if (f_ok) {
                    const uint32_t base = smem_u32(ring + (size_t)slot * p.ad_stride) + (uint32_t)fl * 4u;
                    const int64_t off = (T.row_base + T.n0 + (int64_t)c * AD_R) * T.out_dim + T.m0 + fl;
                    float* gl = gl0 ? gl0 + off : nullptr;
                    // the loads of adam_chunk are complete before its stores are issued (they feed them), so the stage
                    // may be handed back right after the call
                    if (T.wbox == TILE_M && T.out_dim == 256) adam_chunk<TILE_M, 256>(base, gr, adam, gw0 + off, gm0 + off, gv0 + off, gl, 0, 0);
                    else if (T.wbox == TILE_M && T.out_dim == 512) adam_chunk<TILE_M, 512>(base, gr, adam, gw0 + off, gm0 + off, gv0 + off, gl, 0, 0);
                    else adam_chunk<0, 0>(base, gr, adam, gw0 + off, gm0 + off, gv0 + off, gl, T.wbox, T.out_dim);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&wdone[slot]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
            g0 += T.nch; ++i;
        }
        if (n_live == 0 && threadIdx.x == 0 && !p.pdl_early) pdl_release();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, (uint32_t)p.tmem_cols);
}

// --------------------------------------------------------------- ADAM, persistent, any batch: operands in passes
// The persistent kernel above needs all four operand sets of a tile resident at once (batch <= 64).  For larger batches
// (BASELINE.json configs[4]: 256) the K loop of the weight-gradient GEMM runs in passes of ADQ_KG K blocks through the
// same 128 KB operand area: the operand producer refills it as soon as the MMAs of the previous pass are committed and
// the accumulators (a b / small products, see acc_sum16) carry across passes.  A SECOND producer warp owns the w / m / v
// chunk ring, so that a full ring never delays the operands of the next pass.  Everything else -- two accumulator stages
// in tensor memory, the epilogue of tile i under the passes of tile i + 1 -- is tc_adam_pers_kernel.
// Replaces the ring kernel (tc_adam_kernel), which reloads its operand buffers twelve times per tile with the
// w / m / v stream idle: 191 us per full-width launch at configs[4].
constexpr int ADQ_KG = 2;
constexpr int NTHREADS_PERS2 = (4 * AD_MAX_GROUPS + 3) * 32;

__global__ void __launch_bounds__(NTHREADS_PERS2, 1) tc_adam_pers2_kernel(const __grid_constant__ AdamMaps maps1,
                                                                          const __grid_constant__ AdamMaps maps2, const TcParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int nkb = p.nkb_adam;
    const int npass = (nkb + ADQ_KG - 1) / ADQ_KG;
    constexpr uint32_t set_bytes = ADQ_KG * A_STAGE_BYTES;
    uint8_t* set0 = smem;                                           // dout_hi
    uint8_t* set1 = smem + set_bytes;                               // in_lo
    uint8_t* set2 = smem + 2 * (size_t)set_bytes;                   // in_hi
    uint8_t* set3 = smem + 3 * (size_t)set_bytes;                   // dout_lo
    uint8_t* ring = smem + 4 * (size_t)set_bytes;
    __shared__ uint64_t ops_full, ops_empty, tmem_full_bar[2], tmem_empty_bar[2], wfull[ADP_RING], wdone[ADP_RING];
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int OPS_WARP = 4 * AD_MAX_GROUPS, MMA_WARP = OPS_WARP + 1, CHUNK_WARP = OPS_WARP + 2;
    if (threadIdx.x == 0) {
        mbar_init(&ops_full, 1); mbar_init(&ops_empty, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full_bar[i], 1); mbar_init(&tmem_empty_bar[i], 4 * AD_MAX_GROUPS); }
        for (int i = 0; i < ADP_RING; ++i) { mbar_init(&wfull[i], 1); mbar_init(&wdone[i], 4); }
        fence_barrier_init();
    }
    constexpr uint32_t acc_cols = 2 * ADAM_TILE;          // a b tile + small-product tile
    if (warp == 0) tmem_alloc(&tmem_base_slot, 512u);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_slot;
    // the chunk producer reads only what the previous optimiser step wrote: it does not wait for the grid ahead
    if (warp != CHUNK_WARP) pdl_wait();
    if (p.pdl_early && threadIdx.x == 0) pdl_release();
    __syncwarp();

    const int mh = (p.Hp + TILE_M - 1) / TILE_M, mo = (p.Op + TILE_M - 1) / TILE_M, nx2 = (p.Hp + ADAM_TILE - 1) / ADAM_TILE;
    const int tps = p.nx1 * mh + nx2 * mo;
    const int n_tiles = p.row_tiles * tps;
    struct Tile { int s, m0, n0, out_dim, a_c0, b_c0, b_c1, nch, wbox; bool second; int64_t row_base; };
    auto tile_of = [&](int t, Tile& T) {
        T.s = t / tps + p.s_base;
        int r = t % tps;
        T.second = r >= p.nx1 * mh;
        const SubnetDesc d = p.desc[T.s];
        int in_dim;
        if (!T.second) {
            T.n0 = (r / mh) * ADAM_TILE; T.m0 = (r % mh) * TILE_M; T.out_dim = p.Hp; in_dim = d.Pp;
            T.a_c0 = T.s * p.Hp + T.m0; T.b_c0 = (int)d.coff + T.n0; T.b_c1 = (int)p.row0; T.row_base = d.coff; T.wbox = p.wbox;
        } else {
            r -= p.nx1 * mh;
            T.n0 = (r / mo) * ADAM_TILE; T.m0 = (r % mo) * TILE_M; T.out_dim = p.Op; in_dim = p.Hp;
            T.a_c0 = T.s * p.Op + T.m0; T.b_c0 = T.s * p.Hp + T.n0; T.b_c1 = 0; T.row_base = (int64_t)T.s * p.Hp; T.wbox = p.wbox2;
        }
        T.nch = (T.m0 < T.out_dim && T.n0 < in_dim) ? min(ADAM_TILE, in_dim - T.n0) / AD_R : 0;
    };

    if (warp == OPS_WARP) {
        // ===== operand producer: (tile, pass) after (tile, pass), one refill of the operand area each =====
        if (elect_one()) {
            int q = 0;                                    // passes issued so far
            Tile T;
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                tile_of(t, T);
                if (!T.nch) continue;
                const AdamMaps* mp = T.second ? &maps2 : &maps1;
                for (int pass = 0; pass < npass; ++pass, ++q) {
                    const int kb0 = pass * ADQ_KG, cnt = min(ADQ_KG, nkb - kb0);
                    if (q > 0) mbar_wait(&ops_empty, (q - 1) & 1, 9);          // the MMAs of the previous pass have read the area
                    mbar_arrive_expect_tx(&ops_full, 4u * (uint32_t)cnt * A_STAGE_BYTES);
                    for (int k = 0; k < cnt; ++k) {
                        const int row = (kb0 + k) * BLOCK_K;
                        load_stage<true>(set0 + (size_t)k * A_STAGE_BYTES, &mp->A, &ops_full, T.a_c0, row, TILE_M);
                        load_stage<true>(set3 + (size_t)k * A_STAGE_BYTES, &mp->Alo, &ops_full, T.a_c0, row, TILE_M);
                        load_stage<true>(set1 + (size_t)k * A_STAGE_BYTES, &mp->Blo, &ops_full, T.b_c0, T.b_c1 + row, ADAM_TILE);
                        load_stage<true>(set2 + (size_t)k * A_STAGE_BYTES, &mp->B, &ops_full, T.b_c0, T.b_c1 + row, ADAM_TILE);
                    }
                }
            }
        }
    } else if (warp == CHUNK_WARP) {
        // ===== chunk producer: the w / m / v ring, tile after tile =====
        if (elect_one()) {
            int g = 0;
            Tile T;
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                tile_of(t, T);
                if (!T.nch) continue;
                const AdamMaps* mp = T.second ? &maps2 : &maps1;
                const int tile_floats = AD_R * T.wbox;
                for (int c = 0; c < T.nch; ++c, ++g) {
                    const int slot = g % ADP_RING;
                    if (g >= ADP_RING) mbar_wait(&wdone[slot], ((g / ADP_RING) - 1) & 1, 8);
                    float* ws = reinterpret_cast<float*>(ring + (size_t)slot * p.ad_stride);
                    const int32_t r = (int32_t)(T.row_base + T.n0 + c * AD_R);
                    mbar_arrive_expect_tx(&wfull[slot], 3u * (uint32_t)tile_floats * 4u);
                    tma_load_2d(ws, &mp->W, &wfull[slot], T.m0, r);
                    tma_load_2d(ws + tile_floats, &mp->M, &wfull[slot], T.m0, r);
                    tma_load_2d(ws + 2 * tile_floats, &mp->V, &wfull[slot], T.m0, r);
                }
            }
        }
    } else if (warp == MMA_WARP) {
        if (elect_one()) {
            const uint32_t idesc = idesc_for(ADAM_TILE, true, true);
            int i = 0, q = 0;
            Tile T;
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                tile_of(t, T);
                if (!T.nch) continue;
                const int acc = i & 1;
                mbar_wait(&tmem_empty_bar[acc], ((i >> 1) & 1) ^ 1, 12);
                tc_fence_after();
                const uint32_t d_hi = tmem + (uint32_t)acc * acc_cols, d_lo = d_hi + ADAM_TILE;
                for (int pass = 0; pass < npass; ++pass, ++q) {
                    const int cnt = min(ADQ_KG, nkb - pass * ADQ_KG);
                    mbar_wait(&ops_full, q & 1, 6);
                    tc_fence_after();
                    auto mma_round = [&](uint32_t dst, const uint8_t* a, const uint8_t* b, bool fresh) {
                        for (int k = 0; k < cnt; ++k) {
                            const uint32_t sa = smem_u32(a + (size_t)k * A_STAGE_BYTES);
                            const uint32_t sb = smem_u32(b + (size_t)k * A_STAGE_BYTES);
#pragma unroll
                            for (int j = 0; j < BLOCK_K / UMMA_K; ++j)
                                umma_tf32(dst, stage_desc<true>(sa, j), stage_desc<true>(sb, j), idesc, (!fresh || k || j) ? 1u : 0u);
                        }
                    };
                    mma_round(d_lo, set0, set1, pass == 0);                // dout_hi in_lo
                    mma_round(d_hi, set0, set2, pass == 0);                // dout_hi in_hi
                    mma_round(d_lo, set3, set2, false);                    // dout_lo in_hi
                    umma_commit(&ops_empty);
                }
                umma_commit(&tmem_full_bar[acc]);
                ++i;
            }
        }
    } else {
        const int quad = warp & 3, grp = warp >> 2;
        const int fl = quad * 32 + lane;
        const AdamParams adam = adam_of(p);
        int g0 = 0, i = 0, n_live = 0;
        Tile T;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) { tile_of(t, T); n_live += T.nch ? 1 : 0; }
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            tile_of(t, T);
            if (!T.nch) continue;
            const int acc = i & 1;
            const bool f_ok = (T.m0 + fl) < T.out_dim;
            const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)acc * acc_cols;
            mbar_wait(&tmem_full_bar[acc], (i >> 1) & 1, 4);
            tc_fence_after();
            if (i + 1 == n_live && threadIdx.x == 0 && !p.pdl_early) pdl_release();
            __syncwarp();
            float* gw0 = T.second ? p.W2 : p.W1;
            float* gm0 = T.second ? p.mW2 : p.mW1;
            float* gv0 = T.second ? p.vW2 : p.vW1;
            float* gl0 = T.second ? p.W2lo : p.W1lo;
            for (int c = grp; c < T.nch; c += AD_MAX_GROUPS) {
                const int g = g0 + c, slot = g % ADP_RING;
                float gr[AD_R], gl_[AD_R];
                __syncwarp();
                tmem_ld8(taddr + c * AD_R, gr);
                tmem_ld8(taddr + ADAM_TILE + c * AD_R, gl_);
#pragma unroll
                for (int r = 0; r < AD_R; ++r) gr[r] += gl_[r];
                mbar_wait(&wfull[slot], (g / ADP_RING) & 1, 7);
                __syncwarp();
                if (f_ok) {
                    const uint32_t base = smem_u32(ring + (size_t)slot * p.ad_stride) + (uint32_t)fl * 4u;
                    const int64_t off = (T.row_base + T.n0 + (int64_t)c * AD_R) * T.out_dim + T.m0 + fl;
                    float* gl = gl0 ? gl0 + off : nullptr;
                    if (T.wbox == TILE_M && T.out_dim == 256) adam_chunk<TILE_M, 256>(base, gr, adam, gw0 + off, gm0 + off, gv0 + off, gl, 0, 0);
                    else if (T.wbox == TILE_M && T.out_dim == 512) adam_chunk<TILE_M, 512>(base, gr, adam, gw0 + off, gm0 + off, gv0 + off, gl, 0, 0);
                    else adam_chunk<0, 0>(base, gr, adam, gw0 + off, gm0 + off, gv0 + off, gl, T.wbox, T.out_dim);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&wdone[slot]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
            g0 += T.nch; ++i;
        }
        if (n_live == 0 && threadIdx.x == 0 && !p.pdl_early) pdl_release();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512u);
}

// ------------------------------------------------------------------------------------------------ host side
struct TcState {
    // weights / step buffers (fixed for the life of the engine)
    CUtensorMap W1_mn, W2_mn, W2_k;
    CUtensorMap W1_ts, W2_ts;                              // plain [32 k][min(128, out)] weight tiles of the TS kernels
    bool ts = true;                                        // FWD1 / FWD2 take the weights from tensor memory (DEEPIMPUTE_B200_TS=0: from shared memory)
    bool ts_bwd = false;                                   // DEEPIMPUTE_B200_TS_BWD=1: BWD too (written, not yet run on hardware)
    int ts_stages = 0, ts_lo = 0, ts_smem1 = 0, ts_smem2 = 0, ts_tmem = 0, ts_acol0 = 0;
    CUtensorMap H_k, H_mn, DZ2_k, DZ2_mn, DZ1_mn;          // training activations [Bp][...]
    CUtensorMap Hlo_mn, DZ2lo_mn, DZ1lo_mn, Xstep_lo_mn, Xtr_lo_mn;   // residual twins (x3)
    CUtensorMap H_aux;                                     // h tile for the BWD epilogue
    CUtensorMap Xstep_k, Xstep_mn, Ystep_aux;
    CUtensorMap Xchunk_k, Hchunk_k;                        // inference chunk
    CUtensorMap W1_t[3], W2_t[3];                          // {w, m, v} tiles of the ADAM epilogue
    // LT kernels (every operand by TMA): residual twins of the weights and K-major views of the activation twins
    bool lt = false;                                       // DEEPIMPUTE_B200_LT=0 selects the converter-warp kernels
    CUtensorMap W1lo_mn, W2lo_mn, W2lo_k, Hlo_k, DZ2lo_k, Xstep_lo_k, Xtr_lo_k, Xte_lo_k, Xchunk_lo_k, Hchunk_lo_k;
    struct LtCfg { int stages = 0, smem = 0; bool aux = false; };
    LtCfg lt_train, lt_train_noaux, lt_infer;
    bool wlo_stale = true;                                 // converter family: W_lo is not kept by ADAM; inference refreshes it on demand
    int inf_stages = 0, inf_smem = 0;                      // persistent inference kernel
    int lt_ks = 1;                                         // split-K factor of FWD1 / BWD (1: one CTA walks the whole K loop)
    int nacc_ts = 1;                                       // accumulators of the TS kernels (their weight slabs share tensor memory)
    float* kpart[2] = {nullptr, nullptr};                  // scratch tiles of FWD1 / BWD: [S][m_tiles][lt_ks][Bp][128]
    unsigned int* kcount[2] = {nullptr, nullptr};          // arrival counters [S][m_tiles]
    // L2 residency of the optimiser state (DEEPIMPUTE_B200_L2_PERSIST): access-policy window of the training launches
    bool l2_window = false;
    cudaAccessPolicyWindow l2_policy = {};
    // staged train / test matrices (rebuilt by tc_rebind)
    CUtensorMap Xtr_k, Xtr_mn, Xte_k, Ytr_aux;
    bool have_split = false;
    struct Cfg { int stages = 0, lo_stages = 0, smem = 0; bool aux = false; };
    Cfg fwd1_train[3], fwd2_train[3], bwd_train[3], infer;   // [0]: two CTAs per SM where possible, [1]: deepest ring,
                                                             // [2]: "medium": fits beside one ADAM CTA on the same SM
    // epoch graph over sub-network groups
    int n_groups = 1, group_deep = 0;
    static constexpr int MAX_GROUPS = 40;
    int group_s0[MAX_GROUPS + 1] = {0};
    cudaStream_t gstream[MAX_GROUPS][2] = {};
    cudaEvent_t gev[MAX_GROUPS][3] = {};
    cudaEvent_t ev_fork = nullptr;
    cudaGraphExec_t epoch_exec = nullptr;
    int64_t graph_nodes = 0, graph_n_train = -1, lr_capacity = 0;
    uint32_t* d_step_base = nullptr;
    float* d_lr_table = nullptr;
    bool use_graph = true, graph_failed = false;
    int64_t graph_fallbacks = 0;                           // epochs that ran step by step because the graph could not be built
    unsigned long long* d_trace = nullptr;                 // DEEPIMPUTE_B200_TRACE=1: 3 kernels x 256 slots
    int smem_adam = 0;
    bool x3 = false;                                       // forward GEMMs error-compensated (DI_MATH_TF32X3)
    bool x3_bwd = false, simt_adam = false;                // experiments (DEEPIMPUTE_B200_EXPERIMENT bit 0 / bit 1)
    int aux_h = 0, aux_y = 0, wbox1 = 0, wbox2 = 0;
    bool adam_direct = true;                               // DEEPIMPUTE_B200_ADAM_STORE=tma selects the in-place ring + TMA stores
    bool adam_big = false;                                 // one-CTA-per-SM ADAM kernel (DEEPIMPUTE_B200_ADAM=ring disables it)
    bool adam_pers = false;                                // persistent ADAM kernel: several tiles per CTA, overlapped (throughput regime)
    int adam_tpc = 2, smem_adam_pers = 0;                  // tiles per CTA; shared memory
    bool adam_pers2 = false;                               // ... its any-batch form (operands in passes): batches above 64, tf32x3
    int smem_adam_pers2 = 0;
    bool pdl = true;                                       // programmatic dependent launch along a step's kernel chain (DEEPIMPUTE_B200_PDL=0 disables)
    bool pdl_early = false;                                // DEEPIMPUTE_B200_PDL=2: dependents released before the main loop instead of after it
    bool pdl_early_adam = false;                           // ... the ADAM kernel too (its successor is the next step's FWD1)
    bool pdl_prefetch = true;                              // DEEPIMPUTE_B200_PDL_PREFETCH=0: wait for the previous grid before any load
    int pdl_lead = 0;                                      // DEEPIMPUTE_B200_PDL_LEAD=k: release dependents k K blocks before the main loop ends
    int smem_adam_big = 0, ad_nded = 0, ad_stride = 0, ad_groups = 4, ad_kg = 2;
};

void drop_epoch_graph(TcState* st) {
    if (st->epoch_exec) { cudaGraphExecDestroy(st->epoch_exec); st->epoch_exec = nullptr; }
    st->graph_nodes = 0;
}

int pow2_cols(int n) { int c = 32; while (c < n) c <<= 1; return c; }
int smem_for(int n_cols, int stages, int lo_stages, int aux_floats) {
    return (stages + lo_stages) * (int)(A_STAGE_BYTES + n_cols * BLOCK_K * 4) + aux_floats * 4 + 1024;
}
// deepest ring (<= MAX_STAGES) that still leaves room for two CTAs per SM; if even two stages do not fit in half an
// SM, the deepest ring that fits in one
TcState::Cfg pick_cfg(int n_cols, int aux_floats, bool x3, bool deep = false, int only_budget = 0) {
    // (raw slabs, residual slabs) in order of preference; the first that fits the budget wins
    static const int plain[][2] = {{6, 0}, {5, 0}, {4, 0}, {3, 0}, {2, 0}};
    static const int comp[][2] = {{5, 3}, {4, 3}, {4, 2}, {3, 2}, {2, 2}, {2, 1}};
    TcState::Cfg c;
    c.aux = aux_floats > 0;
    for (int budget : {110 * 1024, 224 * 1024}) {
        if (deep && budget < 200 * 1024) continue;
        if (only_budget) budget = only_budget;
        for (int i = 0; i < (x3 ? 6 : 5); ++i) {
            const int hs = x3 ? comp[i][0] : plain[i][0], ls = x3 ? comp[i][1] : 0;
            const int bytes = smem_for(n_cols, hs, ls, aux_floats);
            if (bytes <= budget) { c.stages = hs; c.lo_stages = ls; c.smem = bytes; return c; }
        }
    }
    if (aux_floats > 0) return pick_cfg(n_cols, 0, x3, deep, only_budget);   // large batches: side operand read from global
    return c;      // stages == 0: does not fit
}

TcParams base_params(Engine& e) {
    TcParams p{};
    p.desc = e.d_desc; p.S = e.S; p.H = e.H; p.O = e.O; p.Hp = e.Hp; p.Op = e.Op;
    p.b1 = e.b1; p.mb1 = e.mb1; p.vb1 = e.vb1; p.b2 = e.b2; p.mb2 = e.mb2; p.vb2 = e.vb2;
    p.seed = e.cfg.seed;
    const double r = e.cfg.dropout_rate;
    p.drop_thresh = r > 0.0 ? (uint32_t)(r * 4294967296.0) : 0u;
    p.keep_scale = 1.0f;
    p.nacc = 1; p.lo_acc = 0;
    { static const int teams = [] { const char* v = getenv("DEEPIMPUTE_B200_CONV_TEAMS"); return (v && atoi(v) == 1) ? 1 : 2; }(); p.conv_teams = teams; }
    return p;
}

// kernel launch with or without the programmatic-stream-serialization attribute (see pdl_wait)
const cudaAccessPolicyWindow* g_launch_window = nullptr;   // set around the training launches of an engine that keeps its state in L2

template <typename... KArgs, typename... Args>
void launch_k(void (*kernel)(KArgs...), dim3 grid, int threads, int smem, cudaStream_t stream, bool pdl, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3((unsigned)threads); cfg.dynamicSmemBytes = (size_t)smem; cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (pdl) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    if (g_launch_window) {
        attr[na].id = cudaLaunchAttributeAccessPolicyWindow;
        attr[na].val.accessPolicyWindow = *g_launch_window;
        ++na;
    }
    cfg.attrs = attr; cfg.numAttrs = (unsigned)na;
    cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

template <int OP, bool X3>
void launch(Engine& e, const char* name, const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& c,
            const TcParams& p, dim3 grid, int smem) {
    KernelTimer t(e, name);
    tc_kernel<OP, X3><<<grid, X3 ? NTHREADS_X3 : NTHREADS, smem, e.stream>>>(a, b, c, p);
    count_launch(e, name);
}

inline int cdiv(int a, int b) { return (a + b - 1) / b; }

}  // namespace

bool tc_available() { return true; }

unsigned int tc_take_timeout_word() {
    unsigned int w = 0;
    if (cudaMemcpyFromSymbol(&w, tc::g_mbar_timeout, sizeof w) != cudaSuccess) return 0;
    if (w) { const unsigned int zero = 0; cudaMemcpyToSymbol(tc::g_mbar_timeout, &zero, sizeof zero); }
    return w;
}

bool tc_init(Engine& e) {
    if (e.Bp > 256) { e.err = "DI_MATH_TF32 supports batch sizes up to 256 (use math mode fp32)"; return false; }
    if (!get_encode_fn()) { e.err = "cuTensorMapEncodeTiled is not available from this driver"; return false; }
    auto* st = new TcState();
    e.tc = st;
    const uint64_t SH = (uint64_t)e.S * e.Hp, SO = (uint64_t)e.S * e.Op;
    st->aux_h = (int)std::min<uint64_t>(TILE_M, SH);
    st->aux_y = (int)std::min<uint64_t>(TILE_M, SO);
    st->wbox1 = std::min(TILE_M, e.Hp);
    st->wbox2 = std::min(TILE_M, e.Op);
    bool ok = true;
    ok = ok && make_map_2d(&st->W1_mn, e.W1, e.PT, e.Hp, e.Hp, 32, true);
    ok = ok && make_map_2d(&st->W2_mn, e.W2, SH, e.Op, e.Op, 32, true);
    ok = ok && make_map_2d(&st->W2_k, e.W2, SH, e.Op, e.Op, TILE_M);
    ok = ok && make_map_plain(&st->W1_ts, e.W1, e.PT, e.Hp, e.Hp, st->wbox1, BLOCK_K);
    ok = ok && make_map_plain(&st->W2_ts, e.W2, SH, e.Op, e.Op, st->wbox2, BLOCK_K);
    ok = ok && make_map_2d(&st->H_k, e.Hact, e.Bp, SH, SH, e.Bp);
    ok = ok && make_map_2d(&st->H_mn, e.Hact, e.Bp, SH, SH, 32, true);
    ok = ok && make_map_plain(&st->H_aux, e.Hact, e.Bp, SH, SH, st->aux_h, e.Bp);
    ok = ok && make_map_2d(&st->DZ2_k, e.DZ2, e.Bp, SO, SO, e.Bp);
    ok = ok && make_map_2d(&st->DZ2_mn, e.DZ2, e.Bp, SO, SO, 32, true);
    ok = ok && make_map_2d(&st->DZ1_mn, e.DZ1, e.Bp, SH, SH, 32, true);
    ok = ok && make_map_2d(&st->Xstep_k, e.Xstep, e.Bp, e.PT, e.PT, e.Bp);
    ok = ok && make_map_2d(&st->Xstep_mn, e.Xstep, e.Bp, e.PT, e.PT, 32, true);
    ok = ok && make_map_plain(&st->Ystep_aux, e.Ystep, e.Bp, SO, SO, st->aux_y, e.Bp);
    if (e.cfg.math_mode == DI_MATH_TF32X3) {
        ok = ok && make_map_2d(&st->Hlo_mn, e.Hlo, e.Bp, SH, SH, 32, true);
        ok = ok && make_map_2d(&st->DZ2lo_mn, e.DZ2lo, e.Bp, SO, SO, 32, true);
        ok = ok && make_map_2d(&st->DZ1lo_mn, e.DZ1lo, e.Bp, SH, SH, 32, true);
        ok = ok && make_map_2d(&st->Xstep_lo_mn, e.Xstep_lo, e.Bp, e.PT, e.PT, 32, true);
        // LT kernels: twins of the weights (same views as W1_mn / W2_mn / W2_k) and K-major views of the activation twins
        ok = ok && make_map_2d(&st->W1lo_mn, e.W1lo, e.PT, e.Hp, e.Hp, 32, true);
        ok = ok && make_map_2d(&st->W2lo_mn, e.W2lo, SH, e.Op, e.Op, 32, true);
        ok = ok && make_map_2d(&st->W2lo_k, e.W2lo, SH, e.Op, e.Op, TILE_M);
        ok = ok && make_map_2d(&st->Hlo_k, e.Hlo, e.Bp, SH, SH, e.Bp);
        ok = ok && make_map_2d(&st->DZ2lo_k, e.DZ2lo, e.Bp, SO, SO, e.Bp);
        ok = ok && make_map_2d(&st->Xstep_lo_k, e.Xstep_lo, e.Bp, e.PT, e.PT, e.Bp);
        ok = ok && make_map_2d(&st->Xchunk_lo_k, e.Xchunk_lo, e.chunk_rows, e.PT, e.PT, e.infer_tile);
        ok = ok && make_map_2d(&st->Hchunk_lo_k, e.Hchunk_lo, e.chunk_rows, SH, SH, e.infer_tile);
    }
    ok = ok && make_map_2d(&st->Xchunk_k, e.Xchunk, e.chunk_rows, e.PT, e.PT, e.infer_tile);
    ok = ok && make_map_2d(&st->Hchunk_k, e.Hchunk, e.chunk_rows, SH, SH, e.infer_tile);
    float* w1[3] = {e.W1, e.mW1, e.vW1};
    float* w2[3] = {e.W2, e.mW2, e.vW2};
    for (int i = 0; i < 3; ++i) {
        ok = ok && make_map_plain(&st->W1_t[i], w1[i], e.PT, e.Hp, e.Hp, st->wbox1, AD_R);
        ok = ok && make_map_plain(&st->W2_t[i], w2[i], SH, e.Op, e.Op, st->wbox2, AD_R);

    }
    if (!ok) { e.err = "cuTensorMapEncodeTiled failed"; return false; }
    // shared-memory budgets.  The side-operand tile (aux) is dropped for the compensated FWD2, whose doubled slabs
    // would otherwise push it to one CTA per SM (its 4 x S CTAs are more than one CTA per SM can hold in one wave).
    st->x3 = e.cfg.math_mode == DI_MATH_TF32X3;
    const int aux_floats = e.Bp * TILE_M;
    st->x3_bwd = st->x3;
    if (const char* v = getenv("DEEPIMPUTE_B200_EXPERIMENT")) st->simt_adam = atoi(v) & 2;
    if (const char* v = getenv("DEEPIMPUTE_B200_ADAM_STORE")) st->adam_direct = strcmp(v, "tma") != 0;
    if (const char* v = getenv("DEEPIMPUTE_B200_PDL")) { st->pdl = atoi(v) != 0; st->pdl_early = st->pdl_early_adam = atoi(v) == 2; }
    if (const char* v = getenv("DEEPIMPUTE_B200_PDL_PREFETCH")) st->pdl_prefetch = atoi(v) != 0;
    if (const char* v = getenv("DEEPIMPUTE_B200_PDL_LEAD")) st->pdl_lead = std::max(0, atoi(v));
    for (int deep = 0; deep < 2; ++deep) {
        st->fwd1_train[deep] = pick_cfg(e.Bp, 0, st->x3, deep);
        st->fwd2_train[deep] = pick_cfg(e.Bp, (st->x3 && !deep) ? 0 : aux_floats, st->x3, deep);
        st->bwd_train[deep] = pick_cfg(e.Bp, aux_floats, st->x3, deep);
    }
    // medium: leaves room for one ADAM CTA (its dynamic + static + reserved shared memory) on the same SM; no aux tile
    {
        const int nkb_a = e.Bp / BLOCK_K;
        const int adam_bytes = nkb_a * (int)(A_STAGE_BYTES + ADAM_TILE * BLOCK_K * 4) + AD_STAGES * 3 * AD_R * TILE_M * 4 + 1024;
        const int room = 228 * 1024 - (adam_bytes + 2048) - 2048;
        st->fwd1_train[2] = pick_cfg(e.Bp, 0, st->x3, false, room);
        st->fwd2_train[2] = pick_cfg(e.Bp, 0, st->x3, false, room);
        st->bwd_train[2] = pick_cfg(e.Bp, 0, st->x3, false, room);
        if (!st->fwd1_train[2].stages) { st->fwd1_train[2] = st->fwd1_train[0]; st->fwd2_train[2] = st->fwd2_train[0]; st->bwd_train[2] = st->bwd_train[0]; }
    }
    st->infer = pick_cfg(e.infer_tile, 0, st->x3, true);
    // TS variant of the training FWD1 / FWD2 (X3 only): ring of [A | B] slabs + ring of B residual slabs + Y tile;
    // tensor memory: accumulator (Bp columns) + one 64-column weight slab per residual stage
    if (st->x3) {
        if (const char* v = getenv("DEEPIMPUTE_B200_TS")) st->ts = atoi(v) != 0;
        const int nb = e.Bp * BLOCK_K * 4;
        for (int hs = 6; hs >= 2 && st->ts; --hs) {
            const int ls = std::min(hs, 4);
            const int bytes1 = hs * ((int)A_STAGE_BYTES + nb) + ls * nb + 1024, bytes2 = bytes1 + aux_floats * 4;
            // tensor memory: nacc + 1 accumulators (short chains, acc_sum16) in front of the ls weight slabs
            const int nacc = std::min(4, (512 - ls * TS_COLS) / e.Bp - 1);
            if (bytes2 <= 224 * 1024 && nacc >= 1) {
                st->ts_stages = hs; st->ts_lo = ls; st->ts_smem1 = bytes1; st->ts_smem2 = bytes2;
                st->nacc_ts = nacc;
                st->ts_acol0 = (nacc + 1) * e.Bp;
                st->ts_tmem = pow2_cols(st->ts_acol0 + ls * TS_COLS);
                break;
            }
        }
        if (!st->ts_stages) st->ts = false;
        if (const char* v = getenv("DEEPIMPUTE_B200_TS_BWD")) st->ts_bwd = st->ts && atoi(v) != 0;
    } else {
        st->ts = false;                // the TS kernels are instantiated for the compensated mode only
    }
    // LT kernels: ring of [A | A_lo | B | B_lo] slabs (+ the epilogue's side tile when at least three slabs still fit)
    if (st->x3) {
        // Policy.  The LT kernels read W and W_lo (8 B per weight and GEMM instead of 4) and make the ADAM kernel write
        // W_lo: free while the optimiser state sits in the 126 MB L2, a net loss once it streams from HBM (measured on
        // c3, 40 sub-networks, 180 MB of state: 104 us per step against 94 us with the converter-warp kernels;
        // profiles/r02a_ab_c3.md).  So: LT when the state fits L2 -- few sub-networks per GPU, exactly where the step
        // is bound by the latency of the kernel chain rather than by HBM -- the converter-warp kernels otherwise.
        double max_mb = 100.0;
        if (const char* v = getenv("DEEPIMPUTE_B200_LT_MAX_MB")) max_mb = atof(v);
        st->lt = !st->simt_adam && (double)e.state_bytes <= max_mb * 1048576.0;
        if (const char* v = getenv("DEEPIMPUTE_B200_LT")) st->lt = !st->simt_adam && atoi(v) != 0;
        auto lt_cfg = [&](int n_cols, int aux_fl) {
            TcState::LtCfg c;
            const int stage = 2 * (int)A_STAGE_BYTES + 2 * n_cols * BLOCK_K * 4;
            const int room = 227 * 1024 - 1024 - 2048 - aux_fl * 4;      // alignment slack, static shared memory (cuobjdump: 2048)
            c.stages = std::min(LT_MAX_STAGES, room / stage);
            c.aux = aux_fl > 0;
            c.smem = c.stages * stage + aux_fl * 4 + 1024;
            return c;
        };
        st->lt_train = lt_cfg(e.Bp, aux_floats);
        st->lt_train_noaux = lt_cfg(e.Bp, 0);
        if (st->lt_train.stages < 3) st->lt_train = st->lt_train_noaux;
        st->lt_infer = lt_cfg(e.infer_tile, 0);
        st->inf_stages = 3;                                  // persistent inference kernel: 3 x 64 KB slabs
        st->inf_smem = st->inf_stages * (2 * (int)A_STAGE_BYTES + 2 * INF_TILE * BLOCK_K * 4) + 1024;
        if (st->lt_train.stages < 2 || st->lt_infer.stages < 2) st->lt = false;
        if (st->lt) {
            // split K of FWD1 / BWD: as many CTAs per tile as idle SMs allow (at most 4), never more than K blocks
            const int mh = (e.Hp + TILE_M - 1) / TILE_M;
            int min_kb = e.Op / BLOCK_K;
            for (int s = 0; s < e.S; ++s) min_kb = std::min(min_kb, e.Pp[s] / BLOCK_K);
            st->lt_ks = std::max(1, std::min(std::min(4, min_kb), 148 / std::max(1, mh * e.S)));
            if (const char* v = getenv("DEEPIMPUTE_B200_SPLITK")) st->lt_ks = std::max(1, std::min(std::min(8, min_kb), atoi(v)));
            // dependents released at once (not after the accumulator) when the CTAs of two consecutive kernels of
            // every group fit on the machine together: waiting CTAs then cost nothing (measured at 40 sub-networks
            // per GPU, where they do not fit, early release is a loss: DESIGN.md 3.6)
            const int mo = (e.Op + TILE_M - 1) / TILE_M;
            const int widest = std::max(st->lt_ks * mh + mo, mo + st->lt_ks * mh) * e.S;   // FWD1+FWD2 or FWD2+BWD alive together
            int adam_ctas = 0;
            for (int s = 0; s < e.S; ++s) adam_ctas += ((e.Pp[s] + ADAM_TILE - 1) / ADAM_TILE) * mh + ((e.Hp + ADAM_TILE - 1) / ADAM_TILE) * mo;
            if (!getenv("DEEPIMPUTE_B200_PDL")) {
                st->pdl_early = widest <= 148;
                st->pdl_early_adam = st->pdl_early && adam_ctas + st->lt_ks * mh * e.S <= 148;
            }
            if (st->lt_ks > 1) {
                const size_t tiles = (size_t)e.S * mh;
                for (int k = 0; k < 2 && ok; ++k) {
                    ok = cudaMalloc((void**)&st->kpart[k], tiles * st->lt_ks * e.Bp * TILE_M * sizeof(float)) == cudaSuccess &&
                         cudaMalloc((void**)&st->kcount[k], tiles * sizeof(unsigned int)) == cudaSuccess &&
                         cudaMemset(st->kcount[k], 0, tiles * sizeof(unsigned int)) == cudaSuccess;
                }
                if (!ok) { e.err = "split-K scratch allocation failed"; return false; }
            }
        }
    }
    // L2 residency of the optimiser state: one access-policy window over the state slab.  The persisting share of L2
    // is a device-wide limit; hitRatio = (persisting bytes / window bytes) makes that fraction of the slab's lines
    // stay put while the rest streams, instead of every line evicting another one step before it is needed again.
    {
        // Off by default: measured on c3 it is a large LOSS (2697 ms per bench step against 1561 without): the set-aside
        // takes 79 MB of L2 away from everything that is not the state -- the staged batches, activations and the
        // gather all slow down by more than the state's hits win back (profiles/r02a_ab_c3.md).  Kept as an experiment.
        int want = 0;
        if (const char* v = getenv("DEEPIMPUTE_B200_L2_PERSIST")) want = atoi(v);
        cudaDeviceProp prop{};
        if (want && e.state_slab && cudaGetDeviceProperties(&prop, e.cfg.device) == cudaSuccess &&
            prop.persistingL2CacheMaxSize > 0 && prop.accessPolicyMaxWindowSize > 0) {
            size_t persist = (size_t)prop.persistingL2CacheMaxSize;
            if (want > 1) persist = std::min(persist, (size_t)want << 20);           // value > 1: cap in MiB
            if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, persist) == cudaSuccess) {
                const size_t window = std::min(e.state_bytes, (size_t)prop.accessPolicyMaxWindowSize);
                st->l2_policy.base_ptr = e.state_slab;
                st->l2_policy.num_bytes = window;
                st->l2_policy.hitRatio = (float)std::min(1.0, (double)persist / (double)window);
                st->l2_policy.hitProp = cudaAccessPropertyPersisting;
                st->l2_policy.missProp = cudaAccessPropertyStreaming;
                st->l2_window = true;
            } else cudaGetLastError();
        }
    }
    // sub-network groups of the epoch graph: independent chains on their own streams
    int G = std::min(16, e.S);
    if (const char* v = getenv("DEEPIMPUTE_B200_GROUPS")) G = std::max(1, std::min(std::min((int)TcState::MAX_GROUPS, e.S), atoi(v)));
    if (const char* v = getenv("DEEPIMPUTE_B200_GRAPH")) st->use_graph = atoi(v) != 0;
    st->n_groups = G;
    for (int g = 0; g <= G; ++g) st->group_s0[g] = (int)((int64_t)e.S * g / G);
    // a group's forward / backward grids are small: when all of them fit one CTA per SM, use the deep rings
    const int per_group = (e.S + G - 1) / G;
    st->group_deep = (per_group * cdiv(e.Op, TILE_M) <= 148) ? 1 : 0;
    if (const char* v = getenv("DEEPIMPUTE_B200_DEEP")) st->group_deep = std::max(0, std::min(2, atoi(v)));
    for (int g = 0; g < G; ++g) {
        for (int k = 0; k < 2; ++k)
            if (cudaStreamCreateWithFlags(&st->gstream[g][k], cudaStreamNonBlocking) != cudaSuccess) { e.err = "cudaStreamCreate failed"; return false; }
        for (int k = 0; k < 3; ++k)
            if (cudaEventCreateWithFlags(&st->gev[g][k], cudaEventDisableTiming) != cudaSuccess) { e.err = "cudaEventCreate failed"; return false; }
    }
    if (const char* v = getenv("DEEPIMPUTE_B200_TRACE")) {
        if (atoi(v) && cudaMalloc((void**)&st->d_trace, 5 * 256 * sizeof(unsigned long long)) == cudaSuccess)
            cudaMemset(st->d_trace, 0, 5 * 256 * sizeof(unsigned long long));
    }
    if (cudaEventCreateWithFlags(&st->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaMalloc((void**)&st->d_step_base, sizeof(uint32_t)) != cudaSuccess) { e.err = "epoch-graph set-up failed"; return false; }
    const int nkb = e.Bp / BLOCK_K;
    st->ad_kg = std::min(nkb, 2);         // ring kernel: two K blocks (64 KB of operands) per pass, two CTAs per SM
    st->smem_adam = st->ad_kg * (int)(A_STAGE_BYTES + ADAM_TILE * BLOCK_K * 4) + AD_STAGES * 3 * AD_R * TILE_M * 4 + 1024;
    {   // one-CTA-per-SM variant: all operand sets resident + a stage per chunk
        const int nsets = st->x3 ? 4 : 2;
        const int ops = nsets * nkb * (int)A_STAGE_BYTES;
        st->ad_stride = 3 * AD_R * std::max(st->wbox1, st->wbox2) * 4;
        const int room = 227 * 1024 - 1024 - ops;
        st->ad_nded = room > 0 ? std::min(AD_MAX_CHUNKS, room / st->ad_stride) : 0;
        st->adam_big = st->ad_nded >= 1 && (AD_MAX_CHUNKS - st->ad_nded) * st->ad_stride <= ops;
        st->smem_adam_big = ops + st->ad_nded * st->ad_stride + 1024;
        if (const char* v = getenv("DEEPIMPUTE_B200_ADAM")) if (!strcmp(v, "ring")) st->adam_big = false;
        // persistent form: operand area + 8-stage chunk ring; where throughput counts, i.e. with the converter family
        st->smem_adam_pers = ops + ADP_RING * st->ad_stride + 1024;
        st->adam_pers = st->adam_big && st->x3 && !st->lt && st->smem_adam_pers <= 227 * 1024 - 2048;
        if (const char* v = getenv("DEEPIMPUTE_B200_ADAM")) { if (!strcmp(v, "pers")) st->adam_pers = st->adam_big && st->smem_adam_pers <= 227 * 1024 - 2048; else st->adam_pers = false; }
        if (const char* v = getenv("DEEPIMPUTE_B200_ADAM_TPC")) st->adam_tpc = std::max(1, std::min(8, atoi(v)));
        // batches above 64 (the four operand sets no longer fit): the pass-wise persistent kernel instead of the ring kernel
        st->smem_adam_pers2 = 4 * ADQ_KG * (int)A_STAGE_BYTES + ADP_RING * st->ad_stride + 1024;
        st->adam_pers2 = st->x3 && !st->adam_big && nkb > ADQ_KG && st->smem_adam_pers2 <= 227 * 1024 - 2048;
        if (const char* v = getenv("DEEPIMPUTE_B200_ADAM")) if (!strcmp(v, "ring")) st->adam_pers2 = false;

        if (const char* v = getenv("DEEPIMPUTE_B200_ADAM_GROUPS")) st->ad_groups = std::max(1, std::min(AD_MAX_GROUPS, atoi(v)));
    }
    if (st->smem_adam > 227 * 1024 || !st->fwd1_train[0].stages || !st->fwd2_train[0].stages || !st->bwd_train[0].stages ||
        !st->fwd1_train[1].stages || !st->fwd2_train[1].stages || !st->bwd_train[1].stages || !st->infer.stages) {
        e.err = "tensor-core math modes: this batch size needs more shared memory than one SM has (use math mode fp32)";
        return false;
    }
    const int m1 = std::max(std::max(std::max(st->fwd1_train[0].smem, st->fwd1_train[1].smem), st->fwd1_train[2].smem), st->infer.smem);
    const int m2 = std::max(std::max(std::max(st->fwd2_train[0].smem, st->fwd2_train[1].smem), st->fwd2_train[2].smem), st->infer.smem);
    const int m3 = std::max(std::max(st->bwd_train[0].smem, st->bwd_train[1].smem), st->bwd_train[2].smem);
    cudaError_t ce = cudaSuccess;
    auto set = [&](const void* fn, int bytes) {
        if (ce == cudaSuccess) ce = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    };
    set((const void*)tc_kernel<TC_FWD1, false>, m1); set((const void*)tc_kernel<TC_FWD1, true>, m1);
    set((const void*)tc_kernel<TC_FWD2, false>, m2); set((const void*)tc_kernel<TC_FWD2, true>, m2);
    set((const void*)tc_kernel<TC_BWD, false>, m3);
    set((const void*)tc_kernel<TC_BWD, true>, m3);
    if (st->ts) {
        set((const void*)tc_kernel<TC_FWD1, true, true>, st->ts_smem1);
        set((const void*)tc_kernel<TC_FWD2, true, true>, st->ts_smem2);
        if (st->ts_bwd) set((const void*)tc_kernel<TC_BWD, true, true>, st->ts_smem2);
    }
    if (st->lt) {
        const int ml = std::max(std::max(st->lt_train.smem, st->lt_train_noaux.smem), st->lt_infer.smem);
        set((const void*)tc_lt_kernel<TC_FWD1>, ml);
        set((const void*)tc_lt_kernel<TC_FWD2>, ml);
        set((const void*)tc_lt_kernel<TC_BWD>, ml);
    }
    if (st->x3) {
        set((const void*)tc_lt_infer_kernel<TC_FWD1>, st->inf_smem);
        set((const void*)tc_lt_infer_kernel<TC_FWD2>, st->inf_smem);
    }
    set((const void*)tc_adam_kernel<false>, st->smem_adam);
    set((const void*)tc_adam_kernel<true>, st->smem_adam);
    if (st->adam_big) {
        set((const void*)tc_adam_big_kernel<false>, st->smem_adam_big);
        set((const void*)tc_adam_big_kernel<true>, st->smem_adam_big);
    }
    if (st->adam_pers) {
        set((const void*)tc_adam_pers_kernel<false>, st->smem_adam_pers);
        set((const void*)tc_adam_pers_kernel<true>, st->smem_adam_pers);
    }
    if (st->adam_pers2) set((const void*)tc_adam_pers2_kernel, st->smem_adam_pers2);

    if (ce != cudaSuccess) { e.err = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(ce); return false; }
    return true;
}

void tc_destroy(Engine& e) {
    auto* st = static_cast<TcState*>(e.tc);
    if (st) {
        drop_epoch_graph(st);
        for (int g = 0; g < TcState::MAX_GROUPS; ++g) {
            for (int k = 0; k < 2; ++k) if (st->gstream[g][k]) cudaStreamDestroy(st->gstream[g][k]);
            for (int k = 0; k < 3; ++k) if (st->gev[g][k]) cudaEventDestroy(st->gev[g][k]);
        }
        if (st->ev_fork) cudaEventDestroy(st->ev_fork);
        if (st->d_step_base) cudaFree(st->d_step_base);
        if (st->d_trace) cudaFree(st->d_trace);
        if (st->d_lr_table) cudaFree(st->d_lr_table);
        for (int k = 0; k < 2; ++k) { if (st->kpart[k]) cudaFree(st->kpart[k]); if (st->kcount[k]) cudaFree(st->kcount[k]); }
    }
    delete st;
    e.tc = nullptr;
}

bool tc_rebind(Engine& e) {
    auto* st = static_cast<TcState*>(e.tc);
    if (!st) return true;
    const uint64_t SO = (uint64_t)e.S * e.Op;
    bool ok = true;
    ok = ok && make_map_2d(&st->Xtr_k, e.Xtr, e.n_train_pad, e.PT, e.PT, e.Bp);
    ok = ok && make_map_2d(&st->Xtr_mn, e.Xtr, e.n_train_pad, e.PT, e.PT, 32, true);
    if (e.Xtr_lo) ok = ok && make_map_2d(&st->Xtr_lo_mn, e.Xtr_lo, e.n_train_pad, e.PT, e.PT, 32, true);
    if (e.Xtr_lo) ok = ok && make_map_2d(&st->Xtr_lo_k, e.Xtr_lo, e.n_train_pad, e.PT, e.PT, e.Bp);
    if (e.Xte_lo) ok = ok && make_map_2d(&st->Xte_lo_k, e.Xte_lo, e.n_test_pad, e.PT, e.PT, e.infer_tile);
    ok = ok && make_map_plain(&st->Ytr_aux, e.Ytr, e.n_train_pad, SO, SO, st->aux_y, e.Bp);
    ok = ok && make_map_2d(&st->Xte_k, e.Xte, e.n_test_pad, e.PT, e.PT, e.infer_tile);
    st->have_split = ok;
    drop_epoch_graph(st);             // the graph's nodes hold the old tensor maps
    st->graph_failed = false;
    if (!ok) e.err = "cuTensorMapEncodeTiled failed (staged matrices)";
    return ok;
}

namespace {

// Accumulator plan of one compensated GEMM (see acc_sum16): the cheapest layout whose longest accumulation chain stays
// at or below MAX_CHAIN MMAs.  nkb = K blocks one CTA walks (four a b MMAs and eight small-product MMAs per block);
// budget = tensor-memory columns the kernel may use.  Measured (profiles/r02e_accumulation.md): chains of 192-204 MMAs
// leave a 1e-3 tail in the trained model, chains up to 96 do not.
constexpr int MAX_CHAIN = 64;
void plan_accumulators(int nkb, int n_cols, int budget, int* nacc, int* lo_acc) {
    *nacc = 1; *lo_acc = 0;
    if (nkb * 12 <= MAX_CHAIN || 2 * n_cols > budget) return;                 // one short chain (or no room for a second tile)
    *lo_acc = 1;
    const int want = (nkb * 4 + MAX_CHAIN - 1) / MAX_CHAIN;
    *nacc = std::max(1, std::min(std::min(4, want), budget / n_cols - 1));
}

// Where and how one optimiser step of one sub-network group is launched.
struct StepPlan {
    int s0 = 0, ns = 0;                         // sub-networks [s0, s0 + ns)
    cudaStream_t main = nullptr, side = nullptr; // side != nullptr: ADAM2 runs there, beside ADAM1
    cudaEvent_t ev_bwd = nullptr, ev_adam2 = nullptr;
    bool graph = false;                          // node of the epoch graph: per-epoch values come from device tables
    bool first = false;                          // first step of the capture (nothing to wait for)
    int deep = 0;                                // 1: deep-ring configs (one CTA per SM), for small grids
};

template <int OP, bool X3, bool TS = false>
void launch_on(Engine& e, const StepPlan& pl, const char* name, const CUtensorMap& a, const CUtensorMap& b,
               const CUtensorMap& c, const TcParams& p, dim3 grid, int smem) {
    const bool pdl = static_cast<TcState*>(e.tc)->pdl;
    if (pl.graph) { launch_k(tc_kernel<OP, X3, TS>, grid, X3 ? NTHREADS_X3 : NTHREADS, smem, pl.main, pdl, a, b, c, p); return; }
    KernelTimer t(e, name);
    launch_k(tc_kernel<OP, X3, TS>, grid, X3 ? NTHREADS_X3 : NTHREADS, smem, pl.main, false, a, b, c, p);
    count_launch(e, name);
}

template <int OP>
void launch_lt(Engine& e, const StepPlan& pl, const char* name, const LtMaps& m, const TcParams& p, dim3 grid, int smem) {
    const bool pdl = static_cast<TcState*>(e.tc)->pdl;
    if (pl.graph) { launch_k(tc_lt_kernel<OP>, grid, LT_THREADS, smem, pl.main, pdl, m, p); return; }
    KernelTimer t(e, name);
    launch_k(tc_lt_kernel<OP>, grid, LT_THREADS, smem, pl.main, false, m, p);
    count_launch(e, name);
}

struct WindowScope {    // training launches of an engine whose optimiser state is pinned in L2 carry its access-policy window
    explicit WindowScope(TcState* st) { g_launch_window = st->l2_window ? &st->l2_policy : nullptr; }
    ~WindowScope() { g_launch_window = nullptr; }
};

void launch_step(Engine& e, TcState* st, const StepArgs& a, int which_x, const StepPlan& pl) {
    WindowScope window(st);
    st->wlo_stale = true;
    const CUtensorMap& Xk = which_x == 0 ? st->Xtr_k : st->Xstep_k;
    const CUtensorMap& Yaux = which_x == 0 ? st->Ytr_aux : st->Ystep_aux;
    TcParams p = base_params(e);
    p.n_cols = e.Bp; p.tmem_cols = pow2_cols(e.Bp);
    // short accumulation chains (acc_sum16), planned per kernel below; kernels that may share an SM (pl.deep == 0: two
    // CTAs per SM) stay within half of the 512 tensor-memory columns
    const int acc_budget = (st->lt || pl.deep != 0) ? 512 : 256;
    int max_pp = 0;
    for (int s = pl.s0; s < pl.s0 + pl.ns; ++s) max_pp = std::max(max_pp, e.Pp[s]);
    auto plan = [&](TcParams& q, int nkb, int budget) {
        if (!st->x3) return;
        plan_accumulators(nkb, e.Bp, budget, &q.nacc, &q.lo_acc);
        q.tmem_cols = pow2_cols((q.nacc + q.lo_acc) * e.Bp);
    };
    p.row0 = a.row0; p.rows_per_block_y = 0;
    p.Y = a.Y; p.ldy = a.ldy; p.Hact = e.Hact; p.ldh = (int64_t)e.S * e.Hp; p.DZ2 = e.DZ2; p.DZ1 = e.DZ1;
    p.Hlo = e.Hlo; p.DZ2lo = e.DZ2lo; p.DZ1lo = e.DZ1lo;          // null unless DI_MATH_TF32X3
    p.n_valid = a.n_valid; p.training = 1; p.step = a.step;
    p.keep_scale = p.drop_thresh ? 1.0f / (1.0f - e.cfg.dropout_rate) : 1.0f;
    p.inv_norm = 1.0f / ((float)a.n_valid * (float)e.O);
    p.loss = e.d_loss; p.adam = a.adam;
    p.s_base = pl.s0;
    p.pdl_early = (st->pdl_early && pl.graph) ? 1 : 0;
    p.pdl_prefetch = st->pdl_prefetch ? 1 : 0;
    p.pdl_lead = (st->pdl && pl.graph) ? st->pdl_lead : 0;
    if (pl.graph) { p.step_base = st->d_step_base; p.lr_table = st->d_lr_table; }
    const int mh = cdiv(e.Hp, TILE_M), mo = cdiv(e.Op, TILE_M);
    const TcState::Cfg& c1 = st->fwd1_train[pl.deep];
    const TcState::Cfg& c2 = st->fwd2_train[pl.deep];
    const TcState::Cfg& c3 = st->bwd_train[pl.deep];

    if (st->lt) {
        // every operand by TMA: [W | W_lo | act | act_lo] slabs, twelve MMAs per K block, sixteen epilogue warps
        const TcState::LtCfg& ca = st->lt_train;          // with the epilogue's side tile when it fits
        const TcState::LtCfg& cn = st->lt_train_noaux;
        LtMaps m;
        { TcParams q = p; q.m_tiles = mh; q.Hact = e.Hact - a.row0 * q.ldh; q.Hlo = e.Hlo - a.row0 * q.ldh;
          q.stages = cn.stages; if (!pl.graph && st->d_trace) q.trace = st->d_trace;
          q.kpart = st->kpart[0]; q.kcount = st->kcount[0];
          plan(q, cdiv(max_pp / BLOCK_K, st->lt_ks), 512);
          m.A = st->W1_mn; m.Alo = st->W1lo_mn; m.B = Xk; m.Blo = which_x == 0 ? st->Xtr_lo_k : st->Xstep_lo_k; m.C = Xk;
          launch_lt<TC_FWD1>(e, pl, "fwd1", m, q, dim3(st->lt_ks, mh, pl.ns), cn.smem); }
        { TcParams q = p; q.m_tiles = mo; q.row0 = 0; q.Y = a.Y + a.row0 * a.ldy;
          q.stages = ca.stages; if (!pl.graph && st->d_trace) q.trace = st->d_trace + 256;
          if (ca.aux) { q.aux_cols = st->aux_y; q.aux_row0 = a.row0; }
          plan(q, e.Hp / BLOCK_K, 512);
          m.A = st->W2_mn; m.Alo = st->W2lo_mn; m.B = st->H_k; m.Blo = st->Hlo_k; m.C = Yaux;
          launch_lt<TC_FWD2>(e, pl, "fwd2", m, q, dim3(1, mo, pl.ns), ca.smem); }
        { TcParams q = p; q.m_tiles = mh; q.row0 = 0;
          q.stages = ca.stages; if (!pl.graph && st->d_trace) q.trace = st->d_trace + 512;
          if (ca.aux) { q.aux_cols = st->aux_h; q.aux_row0 = 0; }
          q.kpart = st->kpart[1]; q.kcount = st->kcount[1];
          plan(q, cdiv(e.Op / BLOCK_K, st->lt_ks), 512);
          m.A = st->W2_k; m.Alo = st->W2lo_k; m.B = st->DZ2_k; m.Blo = st->DZ2lo_k; m.C = st->H_aux;
          launch_lt<TC_BWD>(e, pl, "bwd", m, q, dim3(st->lt_ks, mh, pl.ns), ca.smem); }
    } else {
    { TcParams q = p; q.m_tiles = mh; q.Hact = e.Hact - a.row0 * q.ldh;   // kernel indexes h by row0 + b; training h starts at 0
      q.stages = c1.stages; q.lo_stages = c1.lo_stages; if (!pl.graph && st->d_trace) q.trace = st->d_trace;
      plan(q, max_pp / BLOCK_K, st->ts ? st->ts_acol0 : acc_budget);
      if (st->ts) {
          q.stages = st->ts_stages; q.lo_stages = st->ts_lo; q.tmem_cols = st->ts_tmem; q.ts_acol0 = st->ts_acol0; q.ts_wbox = st->wbox1;
          launch_on<TC_FWD1, true, true>(e, pl, "fwd1", st->W1_ts, Xk, Xk, q, dim3(1, mh, pl.ns), st->ts_smem1);
      } else if (st->x3) launch_on<TC_FWD1, true>(e, pl, "fwd1", st->W1_mn, Xk, Xk, q, dim3(1, mh, pl.ns), c1.smem);
      else launch_on<TC_FWD1, false>(e, pl, "fwd1", st->W1_mn, Xk, Xk, q, dim3(1, mh, pl.ns), c1.smem); }
    { TcParams q = p; q.m_tiles = mo; q.row0 = 0; q.Y = a.Y + a.row0 * a.ldy;
      q.stages = c2.stages; q.lo_stages = c2.lo_stages; if (!pl.graph && st->d_trace) q.trace = st->d_trace + 256;
      if (c2.aux) { q.aux_cols = st->aux_y; q.aux_row0 = a.row0; }
      plan(q, e.Hp / BLOCK_K, st->ts ? st->ts_acol0 : acc_budget);
      if (st->ts) {
          q.stages = st->ts_stages; q.lo_stages = st->ts_lo; q.tmem_cols = st->ts_tmem; q.ts_acol0 = st->ts_acol0; q.ts_wbox = st->wbox2;
          q.aux_cols = st->aux_y; q.aux_row0 = a.row0;
          launch_on<TC_FWD2, true, true>(e, pl, "fwd2", st->W2_ts, st->H_k, Yaux, q, dim3(1, mo, pl.ns), st->ts_smem2);
      } else if (st->x3) launch_on<TC_FWD2, true>(e, pl, "fwd2", st->W2_mn, st->H_k, Yaux, q, dim3(1, mo, pl.ns), c2.smem);
      else launch_on<TC_FWD2, false>(e, pl, "fwd2", st->W2_mn, st->H_k, Yaux, q, dim3(1, mo, pl.ns), c2.smem); }
    { TcParams q = p; q.m_tiles = mh; q.row0 = 0;
      if (c3.aux) { q.aux_cols = st->aux_h; q.aux_row0 = 0; }
      q.stages = c3.stages; q.lo_stages = c3.lo_stages; if (!pl.graph && st->d_trace) q.trace = st->d_trace + 512;
      plan(q, e.Op / BLOCK_K, st->ts_bwd ? st->ts_acol0 : acc_budget);
      if (st->ts_bwd) {
          q.stages = st->ts_stages; q.lo_stages = st->ts_lo; q.tmem_cols = st->ts_tmem; q.ts_acol0 = st->ts_acol0; q.ts_wbox = TILE_M;
          q.aux_cols = st->aux_h; q.aux_row0 = 0;
          launch_on<TC_BWD, true, true>(e, pl, "bwd", st->W2_k, st->DZ2_k, st->H_aux, q, dim3(1, mh, pl.ns), st->ts_smem2);
      } else if (st->x3) launch_on<TC_BWD, true>(e, pl, "bwd", st->W2_k, st->DZ2_k, st->H_aux, q, dim3(1, mh, pl.ns), c3.smem);
      else launch_on<TC_BWD, false>(e, pl, "bwd", st->W2_k, st->DZ2_k, st->H_aux, q, dim3(1, mh, pl.ns), c3.smem); }
    }
    if (st->simt_adam && !pl.graph) { simt_adam_only(e, a); return; }
    TcParams q = p;
    q.n_cols = ADAM_TILE; q.tmem_cols = ADAM_TILE; q.nkb_adam = e.Bp / BLOCK_K;
    q.nacc = 1; q.lo_acc = 0;
    if (st->x3 && (e.Bp / BLOCK_K) * 12 > MAX_CHAIN) { q.lo_acc = 1; q.tmem_cols = 2 * ADAM_TILE; }   // long batch: a b and the small products apart
    q.row0 = a.row0; q.wbox = st->wbox1; q.wbox2 = st->wbox2;
    q.W1 = e.W1; q.mW1 = e.mW1; q.vW1 = e.vW1; q.W2 = e.W2; q.mW2 = e.mW2; q.vW2 = e.vW2;
    q.W1lo = st->lt ? e.W1lo : nullptr; q.W2lo = st->lt ? e.W2lo : nullptr;
    q.pdl_early = (st->pdl_early_adam && pl.graph) ? 1 : 0;
    q.adam_direct = st->adam_direct ? 1 : 0;
    if (!pl.graph && st->d_trace) q.trace = st->d_trace + 768;
    int maxPp = 0;
    for (int s = pl.s0; s < pl.s0 + pl.ns; ++s) maxPp = std::max(maxPp, e.Pp[s]);
    q.nx1 = cdiv(maxPp, ADAM_TILE);
    const dim3 grid(q.nx1 + cdiv(e.Hp, ADAM_TILE), std::max(mh, mo), pl.ns);
    AdamMaps m1, m2;
    m1.A = st->DZ1_mn; m1.B = which_x == 0 ? st->Xtr_mn : st->Xstep_mn; m1.W = st->W1_t[0]; m1.M = st->W1_t[1]; m1.V = st->W1_t[2];
    m2.A = st->DZ2_mn; m2.B = st->H_mn; m2.W = st->W2_t[0]; m2.M = st->W2_t[1]; m2.V = st->W2_t[2];
    if (st->x3) {
        m1.Alo = st->DZ1lo_mn; m1.Blo = which_x == 0 ? st->Xtr_lo_mn : st->Xstep_lo_mn;
        m2.Alo = st->DZ2lo_mn; m2.Blo = st->Hlo_mn;
    } else { m1.Alo = m1.A; m1.Blo = m1.B; m2.Alo = m2.A; m2.Blo = m2.B; }
    KernelTimer* t = pl.graph ? nullptr : new KernelTimer(e, "adam");
    q.ad_nded = st->ad_nded; q.ad_stride = st->ad_stride; q.ad_kg = st->ad_kg;
    { static const bool generic = [] { const char* v = getenv("DEEPIMPUTE_B200_ADAM_GENERIC"); return v && atoi(v) != 0; }(); q.ad_generic = generic ? 1 : 0; }
    if (st->adam_pers2) {
        const int nx2 = cdiv(e.Hp, ADAM_TILE);
        const int tiles = pl.ns * (q.nx1 * mh + nx2 * mo);
        q.row_tiles = pl.ns; q.lo_acc = 1; q.tmem_cols = 512;
        // throughput regime: adam_tpc tiles per CTA; latency-bound regime (LT family): one
        const int tpc = st->lt ? 1 : st->adam_tpc;
        launch_k(tc_adam_pers2_kernel, dim3((unsigned)std::min(cdiv(tiles, tpc), 148)), NTHREADS_PERS2, st->smem_adam_pers2, pl.main,
                 st->pdl && pl.graph, m1, m2, q);
    } else if (st->adam_pers) {
        const int nx2 = cdiv(e.Hp, ADAM_TILE);
        const int tiles = pl.ns * (q.nx1 * mh + nx2 * mo);
        q.row_tiles = pl.ns;
        q.tmem_cols = 2 * (q.lo_acc ? 2 * ADAM_TILE : ADAM_TILE);
        const bool pdl = st->pdl && pl.graph;
        // adam_tpc tiles per CTA for the small group launches of the epoch graph; never more CTAs than SMs (a full-width
        // launch is then one persistent CTA per SM walking ~5 tiles)
        const dim3 pgrid((unsigned)std::min(cdiv(tiles, st->adam_tpc), 148));
        if (st->x3) launch_k(tc_adam_pers_kernel<true>, pgrid, NTHREADS_BIG, st->smem_adam_pers, pl.main, pdl, m1, m2, q);
        else launch_k(tc_adam_pers_kernel<false>, pgrid, NTHREADS_BIG, st->smem_adam_pers, pl.main, pdl, m1, m2, q);
    } else if (st->adam_big) {
        const int nthreads = (4 * st->ad_groups + 2) * 32;
        const bool pdl = st->pdl && pl.graph;
        if (st->x3) launch_k(tc_adam_big_kernel<true>, grid, nthreads, st->smem_adam_big, pl.main, pdl, m1, m2, q);
        else launch_k(tc_adam_big_kernel<false>, grid, nthreads, st->smem_adam_big, pl.main, pdl, m1, m2, q);
    } else if (st->x3) launch_k(tc_adam_kernel<true>, grid, NTHREADS, st->smem_adam, pl.main, st->pdl && pl.graph, m1, m2, q);
    else launch_k(tc_adam_kernel<false>, grid, NTHREADS, st->smem_adam, pl.main, st->pdl && pl.graph, m1, m2, q);
    if (t) { delete t; count_launch(e, "adam"); }
}

// Capture one whole epoch -- every optimiser step of every sub-network group -- into a graph that is replayed with a
// single launch per epoch.  Sub-network groups are independent models (reference multinet.py:132-148: the branches
// share nothing), so each group's chain FWD1 -> FWD2 -> BWD -> {ADAM1 | ADAM2} runs on its own pair of streams and
// the latency-bound kernels of one group overlap the bandwidth-bound ones of another.
bool build_epoch_graph(Engine& e, TcState* st) {
    drop_epoch_graph(st);
    const int64_t n_steps = (e.n_train + e.B - 1) / e.B;
    if (n_steps <= 0) return false;
    if (st->lr_capacity < n_steps) {
        if (st->d_lr_table) cudaFree(st->d_lr_table);
        st->d_lr_table = nullptr; st->lr_capacity = 0;
        if (cudaMalloc((void**)&st->d_lr_table, (size_t)n_steps * sizeof(float)) != cudaSuccess) return false;
        st->lr_capacity = n_steps;
    }
    const int G = st->n_groups;
    cudaGraph_t graph = nullptr;
    if (cudaStreamBeginCapture(e.stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); return false; }
    cudaMemsetAsync(e.d_loss, 0, sizeof(double), e.stream);
    cudaEventRecord(st->ev_fork, e.stream);
    for (int g = 0; g < G; ++g) cudaStreamWaitEvent(st->gstream[g][0], st->ev_fork, 0);
    const int64_t ldy = (int64_t)e.S * e.Op;
    for (int64_t i = 0; i < n_steps; ++i) {
        StepArgs a;
        a.X = e.Xtr; a.Y = e.Ytr; a.ldx = e.PT; a.ldy = ldy; a.row0 = i * e.Bp;
        a.n_valid = (int)std::min<int64_t>(e.B, e.n_train - i * e.B);
        a.step = (uint32_t)i;                      // position in the epoch; the kernels add *d_step_base
        a.adam = AdamParams{0.f, 1.0f - e.cfg.beta1, 1.0f - e.cfg.beta2, e.cfg.epsilon};   // lr_t comes from d_lr_table
        for (int g = 0; g < G; ++g) {
            StepPlan pl;
            pl.s0 = st->group_s0[g]; pl.ns = st->group_s0[g + 1] - st->group_s0[g];
            pl.main = st->gstream[g][0];
            pl.graph = true; pl.first = (i == 0); pl.deep = st->group_deep;
            launch_step(e, st, a, 0, pl);
        }
    }
    for (int g = 0; g < G; ++g) {
        cudaEventRecord(st->gev[g][2], st->gstream[g][0]);
        cudaStreamWaitEvent(e.stream, st->gev[g][2], 0);
    }
    cudaError_t ce = cudaStreamEndCapture(e.stream, &graph);
    if (ce != cudaSuccess || !graph) { cudaGetLastError(); return false; }
    ce = cudaGraphInstantiate(&st->epoch_exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess) { cudaGetLastError(); st->epoch_exec = nullptr; return false; }
    st->graph_nodes = n_steps * G * 4;
    st->graph_n_train = e.n_train;
    return true;
}

}  // namespace

void tc_train_step(Engine& e, const StepArgs& a, int which_x) {
    auto* st = static_cast<TcState*>(e.tc);
    StepPlan pl;
    pl.s0 = 0; pl.ns = e.S; pl.main = e.stream;
    if (const char* v = getenv("DEEPIMPUTE_B200_DEEP")) pl.deep = std::max(0, std::min(2, atoi(v)));
    launch_step(e, st, a, which_x, pl);
    if (st->d_trace && which_x == 1) {          // explicit-batch step: dump the pipeline trace of CTA (0,0,0)
        static unsigned long long h[5 * 256];
        cudaStreamSynchronize(e.stream);
        cudaMemcpy(h, st->d_trace, sizeof h, cudaMemcpyDeviceToHost);
        const char* names[3] = {"fwd1", "fwd2", "bwd"};
        for (int k = 0; k < 3; ++k) {
            const unsigned long long* t = h + 256 * k; const unsigned long long t0 = t[0];
            fprintf(stderr, "[trace %s] prologue %llu | aux/tmem wait begins %llu | accumulator ready %llu | epilogue done %llu | exit %llu (cycles from CTA start)\n",
                    names[k], t[1] - t0, t[2] - t0, t[3] - t0, t[4] - t0, t[5] - t0);
            fprintf(stderr, "   kb: tma-issue  mma-slab-ready  mma-lo-ready  conv-done | converter: start  slab-ok  lo-free  converted  fenced\n");
            for (int kb = 0; kb < 40 && t[8 + kb]; ++kb) {
                fprintf(stderr, "   %2d: %9llu %9llu %9llu %9llu", kb, t[8 + kb] - t0, t[48 + kb] - t0, t[88 + kb] - t0,
                        t[128 + kb] ? t[128 + kb] - t0 : 0ull);
                if (kb < 16 && t[168 + 5 * kb])
                    fprintf(stderr, " | %9llu %9llu %9llu %9llu %9llu", t[168 + 5 * kb] - t0, t[169 + 5 * kb] - t0,
                            t[170 + 5 * kb] - t0, t[171 + 5 * kb] - t0, t[172 + 5 * kb] - t0);
                fprintf(stderr, "\n");
            }
        }
        for (int k = 3; k < 4; ++k) {
            const unsigned long long* t = h + 256 * k; const unsigned long long t0 = t[0];
            fprintf(stderr, "[trace %s] epilogue waits for accumulator from %llu | accumulator ready %llu | last chunk done %llu | exit %llu\n",
                    "adam", t[2] - t0, t[3] - t0, t[4] - t0, t[5] - t0);
            fprintf(stderr, "   chunk: tile-in-smem  updated(store issued)\n");
            for (int c = 0; c < 40 && t[8 + c]; ++c)
                fprintf(stderr, "   %2d: %9llu %9llu\n", c, t[8 + c] - t0, t[48 + c] ? t[48 + c] - t0 : 0ull);
        }
        cudaMemset(st->d_trace, 0, sizeof h);
    }
}

bool tc_train_epoch_graph(Engine& e, int64_t first_step, const float* lr_t, int64_t n_steps) {
    auto* st = static_cast<TcState*>(e.tc);
    if (!st || !st->use_graph || st->simt_adam) return false;
    if (!st->epoch_exec || st->graph_n_train != e.n_train) {
        if (st->graph_failed || !build_epoch_graph(e, st)) {
            // not silent: the first failure is reported on stderr, every epoch that runs step by step is counted
            // (di_graph_fallbacks, the bench line's "graph_fallbacks")
            if (!st->graph_failed)
                fprintf(stderr, "deepimpute_b200: the epoch graph could not be built (%s); epochs run step by step from now on\n",
                        cudaGetErrorString(cudaPeekAtLastError()));
            st->graph_failed = true; ++st->graph_fallbacks; cudaGetLastError();
            return false;
        }
    }
    const uint32_t base = (uint32_t)first_step;
    if (cudaMemcpyAsync(st->d_step_base, &base, sizeof base, cudaMemcpyHostToDevice, e.stream) != cudaSuccess) return false;
    if (cudaMemcpyAsync(st->d_lr_table, lr_t, (size_t)n_steps * sizeof(float), cudaMemcpyHostToDevice, e.stream) != cudaSuccess) return false;
    if (cudaGraphLaunch(st->epoch_exec, e.stream) != cudaSuccess) { cudaGetLastError(); return false; }
    st->wlo_stale = true;
    e.launches += st->graph_nodes;
    return true;
}

namespace {
__global__ void residual_kernel(const float* __restrict__ w, float* __restrict__ lo, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        lo[i] = tf32_residual(w[i]);
}
}  // namespace

void tc_weights_changed(Engine& e, int s) {
    if (!e.W1lo || !e.W2lo) return;
    // (this refreshes one sub-network; the others may be stale in the converter family: leave the flag alone)
    const int64_t n1 = (int64_t)e.Pp[s] * e.Hp, o1 = e.coff[s] * e.Hp, n2 = (int64_t)e.Hp * e.Op, o2 = (int64_t)s * e.Hp * e.Op;
    residual_kernel<<<(unsigned)std::min<int64_t>((n1 + 255) / 256, 1184), 256, 0, e.stream>>>(e.W1 + o1, e.W1lo + o1, n1);
    residual_kernel<<<(unsigned)std::min<int64_t>((n2 + 255) / 256, 1184), 256, 0, e.stream>>>(e.W2 + o2, e.W2lo + o2, n2);
}

const char* tc_describe(Engine& e) {
    static thread_local char buf[512];
    auto* st = static_cast<TcState*>(e.tc);
    if (!st) return "fp32 CUDA-core kernels";
    snprintf(buf, sizeof buf,
             "fwd/bwd=%s splitk=%d stages=%d/%d adam=%s groups=%d graph=%d pdl=%d%s l2_window=%d (%.1f MB of %.1f MB state, hitRatio %.2f) graph_fallbacks=%lld",
             st->lt ? "lt" : (st->ts ? "ts" : (st->x3 ? "x3-smem" : "tf32")), st->lt ? st->lt_ks : 1,
             st->lt ? st->lt_train.stages : st->fwd1_train[1].stages, st->lt ? st->lt_infer.stages : st->infer.stages,
             st->adam_pers2 ? "persistent-passes" : (st->adam_pers ? "persistent" : (st->adam_big ? "resident" : "ring")), st->n_groups, st->use_graph ? 1 : 0, st->pdl ? 1 : 0,
             st->pdl_early_adam ? "(early release, all four kernels)" : (st->pdl_early ? "(early release, fwd/bwd)" : ""), st->l2_window ? 1 : 0,
             st->l2_window ? st->l2_policy.num_bytes * (double)st->l2_policy.hitRatio / 1048576.0 : 0.0,
             e.state_bytes / 1048576.0, st->l2_window ? (double)st->l2_policy.hitRatio : 0.0, (long long)st->graph_fallbacks);
    return buf;
}

int64_t tc_fallbacks(Engine& e) {
    auto* st = static_cast<TcState*>(e.tc);
    return st ? st->graph_fallbacks : 0;
}

void tc_forward(Engine& e, int which_x, int64_t row0, int64_t rows, int64_t n_valid, bool with_loss,
                float* out, int64_t ld_out) {
    auto* st = static_cast<TcState*>(e.tc);
    const CUtensorMap& Xk = which_x == 2 ? st->Xte_k : st->Xchunk_k;
    TcParams p = base_params(e);
    p.n_cols = e.infer_tile; p.tmem_cols = e.infer_tile; p.stages = st->infer.stages; p.lo_stages = st->infer.lo_stages;
    if (st->x3) {        // accumulators as in training (one CTA per SM: all 512 columns); longest K loop: FWD1
        plan_accumulators(e.maxPp / BLOCK_K, e.infer_tile, 512, &p.nacc, &p.lo_acc);
        p.tmem_cols = pow2_cols((p.nacc + p.lo_acc) * e.infer_tile);
    }
    p.rows_per_block_y = e.infer_tile;
    p.ldh = (int64_t)e.S * e.Hp;
    p.n_valid = (int)n_valid; p.training = 0; p.drop_thresh = 0;
    const int row_tiles = (int)(rows / e.infer_tile);
    const int mh = cdiv(e.Hp, TILE_M), mo = cdiv(e.Op, TILE_M);
    if (st->x3) {
        // persistent kernel, every operand by TMA.  The converter family does not keep W_lo during training: bring it
        // up to date first (one element-wise pass over the weights, 43 MB at c3)
        if (!st->lt && st->wlo_stale) {
            const int64_t n1 = e.PT * e.Hp, n2 = (int64_t)e.S * e.Hp * e.Op;
            residual_kernel<<<1184, 256, 0, e.stream>>>(e.W1, e.W1lo, n1);
            residual_kernel<<<1184, 256, 0, e.stream>>>(e.W2, e.W2lo, n2);
            st->wlo_stale = false;
        }
        LtMaps m;
        p.stages = st->inf_stages; p.row_tiles = row_tiles;
        { TcParams q = p; q.m_tiles = mh; q.row0 = row0; q.Hact = e.Hchunk - row0 * q.ldh; q.Hlo = e.Hchunk_lo - row0 * q.ldh;
          m.A = st->W1_mn; m.Alo = st->W1lo_mn; m.B = Xk; m.Blo = which_x == 2 ? st->Xte_lo_k : st->Xchunk_lo_k; m.C = Xk;
          const int tiles = e.S * mh * row_tiles;
          KernelTimer t(e, "infer1");
          tc_lt_infer_kernel<TC_FWD1><<<dim3(std::min(tiles, 148)), LT_THREADS, st->inf_smem, e.stream>>>(m, q);
          count_launch(e, "infer1"); }
        { TcParams q = p; q.m_tiles = mo; q.row0 = 0;
          if (with_loss) { q.Y = e.Yte + row0 * (int64_t)e.S * e.Op; q.ldy = (int64_t)e.S * e.Op; q.loss = e.d_loss + 1; }
          q.out = out; q.ld_out = ld_out;
          m.A = st->W2_mn; m.Alo = st->W2lo_mn; m.B = st->Hchunk_k; m.Blo = st->Hchunk_lo_k; m.C = st->Hchunk_k;
          const int tiles = e.S * mo * row_tiles;
          KernelTimer t(e, "infer2");
          tc_lt_infer_kernel<TC_FWD2><<<dim3(std::min(tiles, 148)), LT_THREADS, st->inf_smem, e.stream>>>(m, q);
          count_launch(e, "infer2"); }
        return;
    }
    // hidden activations of this pass live in Hchunk rows [0, rows)
    { TcParams q = p; q.m_tiles = mh; q.row0 = row0; q.Hact = e.Hchunk - row0 * q.ldh;
      if (st->x3) launch<TC_FWD1, true>(e, "infer1", st->W1_mn, Xk, Xk, q, dim3(1, mh * row_tiles, e.S), st->infer.smem);
      else launch<TC_FWD1, false>(e, "infer1", st->W1_mn, Xk, Xk, q, dim3(1, mh * row_tiles, e.S), st->infer.smem); }
    { TcParams q = p; q.m_tiles = mo; q.row0 = 0;
      if (with_loss) { q.Y = e.Yte + row0 * (int64_t)e.S * e.Op; q.ldy = (int64_t)e.S * e.Op; q.loss = e.d_loss + 1; }
      q.out = out; q.ld_out = ld_out;
      if (st->x3) launch<TC_FWD2, true>(e, "infer2", st->W2_mn, st->Hchunk_k, st->Hchunk_k, q, dim3(1, mo * row_tiles, e.S), st->infer.smem);
      else launch<TC_FWD2, false>(e, "infer2", st->W2_mn, st->Hchunk_k, st->Hchunk_k, q, dim3(1, mo * row_tiles, e.S), st->infer.smem); }
}

}  // namespace di
