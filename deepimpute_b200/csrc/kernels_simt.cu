// fp32 CUDA-core path (DI_MATH_FP32): every product is an fp32 FFMA, so results track the fp32 oracle to
// summation-order noise.  One tiled, batched-over-sub-networks GEMM template with the layer-specific work
// fused into its epilogue:
//   FWD1  Hact = dropout(relu(X W1 + b1))                         (reference ops F1+F2)
//   FWD2  yhat = softplus(Hact W2 + b2); wMSE partial; dz2         (F3+F4+B1)
//   BWD   dz1  = (dz2 W2^T) * dropout/relu mask                    (B3)
//   ADAM2 W2  <- Adam(Hact^T dz2)   dW never stored                (B2+U)
//   ADAM1 W1  <- Adam(X^T dz1)                                     (B4+U)
// It also hosts the staging gather and the bias-gradient/Adam kernel used by both math modes.
#include "engine.h"

namespace di {

namespace {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

enum { OP_FWD1 = 0, OP_FWD2 = 1, OP_BWD = 2, OP_ADAM2 = 3, OP_ADAM1 = 4 };

struct GemmParams {
    const SubnetDesc* desc;
    int S, H, O, Hp, Op;
    // operands
    const float* X; int64_t ldx; int64_t row0;      // packed predictors
    const float* Y; int64_t ldy;                    // packed targets (FWD2)
    float* Hact; float* DZ2; float* DZ1;            // [rows][S*Hp], [rows][S*Op], [rows][S*Hp]
    float *W1, *mW1, *vW1, *W2, *mW2, *vW2;
    const float *b1, *b2;
    int rows;                                       // batch rows covered by the grid (multiple of 64)
    int n_valid;
    // epilogue controls
    int training;                                   // dropout on / gradients produced
    uint32_t step; uint64_t seed; uint32_t drop_thresh; float keep_scale;
    float inv_norm;                                 // 1 / (n_valid * O)
    double* loss;                                   // raw sum accumulator (nullptr: none)
    float* out; int64_t ld_out;                     // prediction output (FWD2 inference)
    AdamParams adam;
};

template <int OP>
__global__ void __launch_bounds__(NT) simt_gemm(GemmParams p) {
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    __shared__ double red[NT / 32];

    const int s = blockIdx.z;
    const SubnetDesc d = p.desc[s];
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;

    int M, N, K;
    const float *A, *Bm;
    int64_t a_ms, a_ks, b_ks, b_ns;
    if constexpr (OP == OP_FWD1) {
        M = p.rows; N = p.Hp; K = d.Pp;
        A = p.X + p.row0 * p.ldx + d.coff; a_ms = p.ldx; a_ks = 1;
        Bm = p.W1 + d.coff * p.Hp; b_ks = p.Hp; b_ns = 1;
    } else if constexpr (OP == OP_FWD2) {
        M = p.rows; N = p.Op; K = p.Hp;
        A = p.Hact + (int64_t)s * p.Hp; a_ms = (int64_t)p.S * p.Hp; a_ks = 1;
        Bm = p.W2 + (int64_t)s * p.Hp * p.Op; b_ks = p.Op; b_ns = 1;
    } else if constexpr (OP == OP_BWD) {
        M = p.rows; N = p.Hp; K = p.Op;
        A = p.DZ2 + (int64_t)s * p.Op; a_ms = (int64_t)p.S * p.Op; a_ks = 1;
        Bm = p.W2 + (int64_t)s * p.Hp * p.Op; b_ks = 1; b_ns = p.Op;
    } else if constexpr (OP == OP_ADAM2) {
        M = p.Hp; N = p.Op; K = p.rows;
        A = p.Hact + (int64_t)s * p.Hp; a_ms = 1; a_ks = (int64_t)p.S * p.Hp;
        Bm = p.DZ2 + (int64_t)s * p.Op; b_ks = (int64_t)p.S * p.Op; b_ns = 1;
    } else {
        M = d.Pp; N = p.Hp; K = p.rows;
        A = p.X + p.row0 * p.ldx + d.coff; a_ms = 1; a_ks = p.ldx;
        Bm = p.DZ1 + (int64_t)s * p.Hp; b_ks = (int64_t)p.S * p.Hp; b_ns = 1;
    }
    if (m0 >= M || n0 >= N) return;

    const int tid = threadIdx.x;
    const int tm = (tid / 16) * 4, tn = (tid % 16) * 4;   // 4x4 micro-tile
    float acc[4][4] = {};

    constexpr bool A_K_CONTIG = (OP == OP_FWD1 || OP == OP_FWD2 || OP == OP_BWD);
    constexpr bool B_K_CONTIG = (OP == OP_BWD);

    for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
        for (int i = 0; i < (BM * BK) / NT; ++i) {
            int idx = tid + i * NT, m, k;
            if (A_K_CONTIG) { k = idx % BK; m = idx / BK; } else { m = idx % BM; k = idx / BM; }
            float v = 0.f;
            if (m0 + m < M && k0 + k < K) v = A[(int64_t)(m0 + m) * a_ms + (int64_t)(k0 + k) * a_ks];
            As[k][m] = v;
        }
#pragma unroll
        for (int i = 0; i < (BN * BK) / NT; ++i) {
            int idx = tid + i * NT, n, k;
            if (B_K_CONTIG) { k = idx % BK; n = idx / BK; } else { n = idx % BN; k = idx / BN; }
            float v = 0.f;
            if (n0 + n < N && k0 + k < K) v = Bm[(int64_t)(k0 + k) * b_ks + (int64_t)(n0 + n) * b_ns];
            Bs[k][n] = v;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(&As[k][tm]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tn]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }

    // ------------------------------------------------------------------------------------------ epilogues
    if constexpr (OP == OP_FWD1) {
        // rows m = batch row b (4 consecutive, 4-aligned), cols n = hidden unit
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int hcol = n0 + tn + j;
            if (hcol >= N) continue;
            const float bias = p.b1[(int64_t)s * p.Hp + hcol];
            uint32_t w[4] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
            if (p.training && p.drop_thresh)
                dropout_words((uint32_t)hcol, (uint32_t)((m0 + tm) >> 2), (uint32_t)d.gid, p.step, p.seed, w);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int b = m0 + tm + i;
                if (b >= M) continue;
                float a = fmaxf(acc[i][j] + bias, 0.f);
                if (p.training && p.drop_thresh) a = (w[i] >= p.drop_thresh) ? a * p.keep_scale : 0.f;
                p.Hact[(int64_t)b * p.S * p.Hp + (int64_t)s * p.Hp + hcol] = a;
            }
        }
    } else if constexpr (OP == OP_FWD2) {
        double part = 0.0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int b = m0 + tm + i;
            if (b >= M) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int o = n0 + tn + j;
                if (o >= p.O) continue;
                const float z = acc[i][j] + p.b2[(int64_t)s * p.Op + o];
                const float yhat = softplus_f(z);
                if (p.out) {
                    if (b < p.n_valid) p.out[(int64_t)b * p.ld_out + (int64_t)s * p.O + o] = yhat;
                }
                if (p.Y) {
                    const float y = p.Y[(p.row0 + b) * p.ldy + (int64_t)s * p.Op + o];
                    const float diff = y - yhat;
                    part += (double)(y * diff * diff);
                    if (p.training)
                        p.DZ2[(int64_t)b * p.S * p.Op + (int64_t)s * p.Op + o] =
                            2.0f * y * (yhat - y) * sigmoid_f(z) * p.inv_norm;
                }
            }
        }
        if (p.loss) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
            if ((tid & 31) == 0) red[tid >> 5] = part;
            __syncthreads();
            if (tid == 0) {
                double t = 0.0;
                for (int i = 0; i < NT / 32; ++i) t += red[i];
                atomicAdd(p.loss, t);
            }
        }
    } else if constexpr (OP == OP_BWD) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int b = m0 + tm + i;
            if (b >= M) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int hcol = n0 + tn + j;
                if (hcol >= N) continue;
                const int64_t idx = (int64_t)b * p.S * p.Hp + (int64_t)s * p.Hp + hcol;
                p.DZ1[idx] = (p.Hact[idx] > 0.f) ? acc[i][j] * p.keep_scale : 0.f;
            }
        }
    } else {
        float *W, *mW, *vW;
        int64_t base; int ldw;
        if constexpr (OP == OP_ADAM2) { W = p.W2; mW = p.mW2; vW = p.vW2; base = (int64_t)s * p.Hp * p.Op; ldw = p.Op; }
        else { W = p.W1; mW = p.mW1; vW = p.vW1; base = d.coff * p.Hp; ldw = p.Hp; }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = m0 + tm + i;
            if (r >= M) continue;
            const int c = n0 + tn;
            if (c + 3 < N) {
                const int64_t idx = base + (int64_t)r * ldw + c;
                float4 w = *reinterpret_cast<float4*>(W + idx);
                float4 m = *reinterpret_cast<float4*>(mW + idx);
                float4 v = *reinterpret_cast<float4*>(vW + idx);
                adam_update(acc[i][0], w.x, m.x, v.x, p.adam);
                adam_update(acc[i][1], w.y, m.y, v.y, p.adam);
                adam_update(acc[i][2], w.z, m.z, v.z, p.adam);
                adam_update(acc[i][3], w.w, m.w, v.w, p.adam);
                *reinterpret_cast<float4*>(W + idx) = w;
                *reinterpret_cast<float4*>(mW + idx) = m;
                *reinterpret_cast<float4*>(vW + idx) = v;
            } else {
                for (int j = 0; j < 4 && c + j < N; ++j) {
                    const int64_t idx = base + (int64_t)r * ldw + c + j;
                    adam_update(acc[i][j], W[idx], mW[idx], vW[idx], p.adam);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ gather
__global__ void gather_kernel(const float* __restrict__ norm, int64_t G, const int32_t* __restrict__ rows,
                              const int32_t* __restrict__ perm, int64_t first_row, int64_t n_out,
                              int64_t n_valid, const int32_t* __restrict__ cols, int64_t width,
                              float* __restrict__ out, int batch, int batch_pitch, float* __restrict__ out_lo) {
    for (int64_t io = blockIdx.y; io < n_out; io += gridDim.y) {
        float* dst = out + io * width;
        float* dst_lo = out_lo ? out_lo + io * width : nullptr;
        int64_t i = io;
        bool valid = io < n_valid;
        if (batch > 0) {
            const int64_t r = io % batch_pitch;
            i = (io / batch_pitch) * batch + r;
            valid = r < batch && i < n_valid;
        }
        if (!valid) {
            for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < width; j += (int64_t)gridDim.x * blockDim.x) {
                dst[j] = 0.f;
                if (dst_lo) dst_lo[j] = 0.f;
            }
            continue;
        }
        const int64_t r = rows ? (int64_t)rows[perm ? perm[i] : i] : first_row + i;
        const float* src = norm + r * G;
        for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < width; j += (int64_t)gridDim.x * blockDim.x) {
            const int32_t c = cols[j];
            const float v = (c >= 0) ? __ldg(src + c) : 0.f;
            dst[j] = v;
            if (dst_lo) dst_lo[j] = v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
        }
    }
}

// One block per output row (grid-stride): the whole source row of norm goes through shared memory once.
__global__ void __launch_bounds__(512) gather_xy_kernel(const float* __restrict__ norm, int64_t G, const int32_t* __restrict__ rows,
                                                        const int32_t* __restrict__ perm, int64_t n_out, int64_t n_valid,
                                                        int batch, int batch_pitch,
                                                        const int32_t* __restrict__ xcols, int64_t xw, float* __restrict__ X,
                                                        float* __restrict__ X_lo,
                                                        const int32_t* __restrict__ ycols, int64_t yw, float* __restrict__ Y) {
    extern __shared__ float srow[];
    for (int64_t io = blockIdx.x; io < n_out; io += gridDim.x) {
        int64_t i = io;
        bool valid = io < n_valid;
        if (batch > 0) {
            const int64_t r = io % batch_pitch;
            i = (io / batch_pitch) * batch + r;
            valid = r < batch && i < n_valid;
        }
        float* xd = X + io * xw;
        float* xl = X_lo ? X_lo + io * xw : nullptr;
        float* yd = Y + io * yw;
        if (!valid) {                                   // padding row (uniform per block)
            for (int64_t j = threadIdx.x; j < xw; j += blockDim.x) { xd[j] = 0.f; if (xl) xl[j] = 0.f; }
            for (int64_t j = threadIdx.x; j < yw; j += blockDim.x) yd[j] = 0.f;
            continue;
        }
        const int64_t r = (int64_t)rows[perm ? perm[i] : i];
        const float* src = norm + r * G;
        __syncthreads();                                // previous row fully consumed
        if ((G & 3) == 0) {
            const float4* s4 = reinterpret_cast<const float4*>(src);
            float4* d4 = reinterpret_cast<float4*>(srow);
            for (int64_t j = threadIdx.x; j < (G >> 2); j += blockDim.x) d4[j] = __ldg(s4 + j);
        } else {
            for (int64_t j = threadIdx.x; j < G; j += blockDim.x) srow[j] = __ldg(src + j);
        }
        __syncthreads();
        // four columns per thread: one 16-byte index load and one 16-byte store per array instead of four of each
        // (xw and yw are multiples of 32 and every row starts 128-byte aligned)
        auto pick = [&](int32_t c) { return (c >= 0) ? srow[c] : 0.f; };
        auto lo = [](float v) { return v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); };
        const int4* xc4 = reinterpret_cast<const int4*>(xcols);
        float4* xd4 = reinterpret_cast<float4*>(xd);
        float4* xl4 = reinterpret_cast<float4*>(xl);
        for (int64_t j = threadIdx.x; j < (xw >> 2); j += blockDim.x) {
            const int4 c = __ldg(xc4 + j);
            const float4 v = make_float4(pick(c.x), pick(c.y), pick(c.z), pick(c.w));
            __stcs(xd4 + j, v);
            if (xl) __stcs(xl4 + j, make_float4(lo(v.x), lo(v.y), lo(v.z), lo(v.w)));
        }
        const int4* yc4 = reinterpret_cast<const int4*>(ycols);
        float4* yd4 = reinterpret_cast<float4*>(yd);
        for (int64_t j = threadIdx.x; j < (yw >> 2); j += blockDim.x) {
            const int4 c = __ldg(yc4 + j);
            __stcs(yd4 + j, make_float4(pick(c.x), pick(c.y), pick(c.z), pick(c.w)));
        }
    }
}

// ------------------------------------------------------------------------------------------ bias + Adam
// db2[s][o] = sum_b DZ2[b][s*Op+o],  db1[s][h] = sum_b DZ1[b][s*Hp+h]; one thread per bias element.
__global__ void bias_adam_kernel(const float* __restrict__ DZ2, const float* __restrict__ DZ1, int rows,
                                 int64_t n2, int64_t n1, float* b2, float* mb2, float* vb2,
                                 float* b1, float* mb1, float* vb1, AdamParams adam) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n2) {
        float g = 0.f;
        for (int b = 0; b < rows; ++b) g += DZ2[(int64_t)b * n2 + i];
        adam_update(g, b2[i], mb2[i], vb2[i], adam);
    } else if (i < n2 + n1) {
        const int64_t k = i - n2;
        float g = 0.f;
        for (int b = 0; b < rows; ++b) g += DZ1[(int64_t)b * n1 + k];
        adam_update(g, b1[k], mb1[k], vb1[k], adam);
    }
}

GemmParams base_params(Engine& e) {
    GemmParams p{};
    p.desc = e.d_desc; p.S = e.S; p.H = e.H; p.O = e.O; p.Hp = e.Hp; p.Op = e.Op;
    p.W1 = e.W1; p.mW1 = e.mW1; p.vW1 = e.vW1; p.W2 = e.W2; p.mW2 = e.mW2; p.vW2 = e.vW2;
    p.b1 = e.b1; p.b2 = e.b2;
    p.seed = e.cfg.seed;
    const double r = e.cfg.dropout_rate;
    p.drop_thresh = r > 0.0 ? (uint32_t)(r * 4294967296.0) : 0u;
    p.keep_scale = 1.0f;
    return p;
}

inline int cdiv(int a, int b) { return (a + b - 1) / b; }

}  // namespace

void launch_gather(Engine& e, const int32_t* rows, const int32_t* perm, int64_t first_row, int64_t n_out,
                   int64_t n_valid, const int32_t* cols, int64_t width, float* out, int batch, int batch_pitch,
                   float* out_lo) {
    if (n_out <= 0 || width <= 0) return;
    KernelTimer t(e, "gather");
    dim3 grid((unsigned)std::min<int64_t>((width + 255) / 256, 64), (unsigned)std::min<int64_t>(n_out, 16384));
    gather_kernel<<<grid, 256, 0, e.stream>>>(e.d_norm, e.G, rows, perm, first_row, n_out, n_valid, cols, width, out,
                                              batch, batch_pitch, out_lo);
    count_launch(e, "gather");
}

void launch_gather_xy(Engine& e, const int32_t* rows, const int32_t* perm, int64_t n_out, int64_t n_valid,
                      int batch, int batch_pitch, float* X, float* X_lo, float* Y) {
    if (n_out <= 0) return;
    const int64_t ldy = (int64_t)e.S * e.Op;
    const size_t smem = (size_t)e.G * sizeof(float);
    static int max_smem = -1;
    if (max_smem < 0) {
        int dev = 0; cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    }
    if (!rows || smem > (size_t)max_smem) {             // row does not fit in shared memory: two plain gathers
        launch_gather(e, rows, perm, 0, n_out, n_valid, e.d_pred_cols, e.PT, X, batch, batch_pitch, X_lo);
        launch_gather(e, rows, perm, 0, n_out, n_valid, e.d_targ_cols, ldy, Y, batch, batch_pitch);
        return;
    }
    KernelTimer t(e, "gather");
    cudaFuncSetAttribute(gather_xy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int per_sm = std::max(1, std::min(4, (int)((size_t)max_smem / std::max<size_t>(smem, 1))));
    const unsigned grid = (unsigned)std::min<int64_t>(n_out, (int64_t)148 * per_sm);
    gather_xy_kernel<<<grid, 512, smem, e.stream>>>(e.d_norm, e.G, rows, perm, n_out, n_valid, batch, batch_pitch,
                                                    e.d_pred_cols, e.PT, X, X_lo, e.d_targ_cols, ldy, Y);
    count_launch(e, "gather");
}

void launch_bias_adam(Engine& e, const AdamParams& adam) {
    KernelTimer t(e, "bias");
    const int64_t n2 = (int64_t)e.S * e.Op, n1 = (int64_t)e.S * e.Hp;
    const int64_t n = n1 + n2;
    bias_adam_kernel<<<(unsigned)((n + 255) / 256), 256, 0, e.stream>>>(e.DZ2, e.DZ1, e.Bp, n2, n1, e.b2, e.mb2, e.vb2,
                                                                        e.b1, e.mb1, e.vb1, adam);
    count_launch(e, "bias");
}

void simt_train_step(Engine& e, const StepArgs& a) {
    GemmParams p = base_params(e);
    p.X = a.X; p.ldx = a.ldx; p.row0 = a.row0; p.Y = a.Y; p.ldy = a.ldy;
    p.Hact = e.Hact; p.DZ2 = e.DZ2; p.DZ1 = e.DZ1;
    p.rows = e.Bp; p.n_valid = a.n_valid; p.training = 1; p.step = a.step;
    p.keep_scale = p.drop_thresh ? 1.0f / (1.0f - e.cfg.dropout_rate) : 1.0f;
    p.inv_norm = 1.0f / ((float)a.n_valid * (float)e.O);
    p.loss = e.d_loss; p.adam = a.adam;
    const int mb = cdiv(e.Bp, BM);
    { KernelTimer t(e, "fwd1");
      simt_gemm<OP_FWD1><<<dim3(cdiv(e.Hp, BN), mb, e.S), NT, 0, e.stream>>>(p); count_launch(e, "fwd1"); }
    { KernelTimer t(e, "fwd2");
      simt_gemm<OP_FWD2><<<dim3(cdiv(e.Op, BN), mb, e.S), NT, 0, e.stream>>>(p); count_launch(e, "fwd2"); }
    { KernelTimer t(e, "bwd");
      simt_gemm<OP_BWD><<<dim3(cdiv(e.Hp, BN), mb, e.S), NT, 0, e.stream>>>(p); count_launch(e, "bwd"); }
    { KernelTimer t(e, "adam2");
      simt_gemm<OP_ADAM2><<<dim3(cdiv(e.Op, BN), cdiv(e.Hp, BM), e.S), NT, 0, e.stream>>>(p); count_launch(e, "adam2"); }
    { KernelTimer t(e, "adam1");
      simt_gemm<OP_ADAM1><<<dim3(cdiv(e.Hp, BN), cdiv(e.maxPp, BM), e.S), NT, 0, e.stream>>>(p); count_launch(e, "adam1"); }
    launch_bias_adam(e, a.adam);
}

// exact-fp32 weight updates only (dW on the CUDA cores + Adam); biases are left to the caller.  Used by the
// tensor-core path as the reference implementation of its ADAM kernel in experiments.
void simt_adam_only(Engine& e, const StepArgs& a) {
    GemmParams p = base_params(e);
    p.X = a.X; p.ldx = a.ldx; p.row0 = a.row0; p.Y = a.Y; p.ldy = a.ldy;
    p.Hact = e.Hact; p.DZ2 = e.DZ2; p.DZ1 = e.DZ1;
    p.rows = e.Bp; p.n_valid = a.n_valid; p.training = 1; p.step = a.step; p.adam = a.adam;
    { KernelTimer t(e, "adam2");
      simt_gemm<OP_ADAM2><<<dim3(cdiv(e.Op, BN), cdiv(e.Hp, BM), e.S), NT, 0, e.stream>>>(p); count_launch(e, "adam2"); }
    { KernelTimer t(e, "adam1");
      simt_gemm<OP_ADAM1><<<dim3(cdiv(e.Hp, BN), cdiv(e.maxPp, BM), e.S), NT, 0, e.stream>>>(p); count_launch(e, "adam1"); }
}

void simt_forward(Engine& e, const float* X, int64_t ldx, int64_t rows, int64_t n_valid, float* Hbuf,
                  const float* Y, int64_t ldy, float* out, int64_t ld_out) {
    GemmParams p = base_params(e);
    p.X = X; p.ldx = ldx; p.row0 = 0; p.Y = Y; p.ldy = ldy;
    p.Hact = Hbuf; p.rows = (int)rows; p.n_valid = (int)n_valid; p.training = 0; p.drop_thresh = 0;
    p.loss = Y ? e.d_loss + 1 : nullptr; p.out = out; p.ld_out = ld_out;
    const int mb = cdiv((int)rows, BM);
    { KernelTimer t(e, "infer1");
      simt_gemm<OP_FWD1><<<dim3(cdiv(e.Hp, BN), mb, e.S), NT, 0, e.stream>>>(p); count_launch(e, "infer1"); }
    { KernelTimer t(e, "infer2");
      simt_gemm<OP_FWD2><<<dim3(cdiv(e.Op, BN), mb, e.S), NT, 0, e.stream>>>(p); count_launch(e, "infer2"); }
}

}  // namespace di
