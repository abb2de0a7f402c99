// Predictor selection on the GPU: |Pearson r| between genes on RAW counts
// (reference get_distance_matrix, multinet.py:20-34: np.abs(np.corrcoef(raw.T)) with NaN -> 0) and, per target gene,
// the ntop most correlated candidate genes outside the target's own sub-network (setPredictors, multinet.py:356-360:
// argsort(-|r|)[:, :ntop]).  This is the O(G^2 N) step of fit: ~4e13 flops at 50k cells x 20k genes, minutes in
// float64 numpy on the host, about two seconds here.
//
//   1. column statistics in double (two passes: mean, then sum of squared deviations)
//   2. Z = (x - mean) / sqrt(ss) in place, fp32  (|r_ij| = |sum_n Z_ni Z_nj|)
//   3. C = Z^T Z with an fp32 register-blocked kernel (128 x 128 tiles, 8 x 8 per thread); both operand tiles are
//      row slices of Z, so every global read is a contiguous 512-byte segment and no transpose is needed
//   4. per target row: masked top-ntop scan of |C| in the caller's candidate order (ties go to the earlier candidate)
//
// fp32 accumulation over N terms carries ~1e-6 relative error against the reference's float64; selections can differ
// from the host path only where the ntop-th and (ntop+1)-th correlations of a target are closer than that.
#include <cuda_runtime.h>
#include <stdint.h>
#include <cmath>
#include <cstdio>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "../../include/deepimpute_b200.h"

namespace {

thread_local std::string g_err;

#define CK(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t err__ = (call);                                                                       \
        if (err__ != cudaSuccess) {                                                                       \
            char buf__[512];                                                                              \
            snprintf(buf__, sizeof buf__, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(err__),      \
                     __FILE__, __LINE__);                                                                 \
            g_err = buf__;                                                                                \
            rc = err__ == cudaErrorMemoryAllocation ? DI_ERR_OOM : DI_ERR_CUDA;                           \
            goto done;                                                                                    \
        }                                                                                                 \
    } while (0)

constexpr int ROWS_PER_BLOCK = 512;

// acc[g] += sum over this block's rows of (x - shift[g])^power ; power 1 with shift = 0 gives the column sums
__global__ void colsum_kernel(const float* __restrict__ x, int64_t N, int64_t G, const double* __restrict__ mean,
                              double* __restrict__ acc) {
    const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (g >= G) return;
    const int64_t r0 = (int64_t)blockIdx.y * ROWS_PER_BLOCK, r1 = min(N, r0 + ROWS_PER_BLOCK);
    double s = 0.0;
    if (mean) {
        const double m = mean[g];
        for (int64_t r = r0; r < r1; ++r) { const double d = (double)x[r * G + g] - m; s += d * d; }
    } else {
        for (int64_t r = r0; r < r1; ++r) s += (double)x[r * G + g];
    }
    atomicAdd(acc + g, s);
}

__global__ void finish_mean_kernel(double* mean, int64_t G, int64_t N) {
    const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (g < G) mean[g] /= (double)N;
}

__global__ void standardise_kernel(float* __restrict__ x, int64_t N, int64_t G, const double* __restrict__ mean,
                                   const double* __restrict__ ss) {
    const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (g >= G) return;
    const double m = mean[g];
    const double inv = ss[g] > 0.0 ? 1.0 / sqrt(ss[g]) : 0.0;       // zero-variance gene: correlation NaN -> 0
    const int64_t r0 = (int64_t)blockIdx.y * ROWS_PER_BLOCK, r1 = min(N, r0 + ROWS_PER_BLOCK);
    for (int64_t r = r0; r < r1; ++r) x[r * G + g] = (float)(((double)x[r * G + g] - m) * inv);
}

// C[i][j] = sum_n Z[n][i] * Z[n][j];  one 128 x 128 tile per block, 256 threads, 8 x 8 per thread, K slabs of 8 rows
constexpr int TM = 128, TK = 8;
__global__ void __launch_bounds__(256) gram_kernel(const float* __restrict__ Z, int64_t N, int64_t G, float* __restrict__ C) {
    __shared__ __align__(16) float As[2][TK][TM];
    __shared__ __align__(16) float Bs[2][TK][TM];
    const int64_t i0 = (int64_t)blockIdx.y * TM, j0 = (int64_t)blockIdx.x * TM;
    const int tid = threadIdx.x;
    const int ty = tid / 16, tx = tid % 16;                  // thread owns rows ty*8.., cols tx*8..
    // loader: 256 threads x float4 = one [8][128] slab per operand
    const int lk = tid / 32, lc = (tid % 32) * 4;
    const bool a_ok = i0 + lc < G, b_ok = j0 + lc < G;       // G is padded to a multiple of 4 by the caller
    float acc[8][8] = {};
    auto load = [&](int buf, int64_t n0) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        const int64_t n = n0 + lk;
        if (n < N) {
            if (a_ok) a = __ldg(reinterpret_cast<const float4*>(Z + n * G + i0 + lc));
            if (b_ok) b = __ldg(reinterpret_cast<const float4*>(Z + n * G + j0 + lc));
        }
        *reinterpret_cast<float4*>(&As[buf][lk][lc]) = a;
        *reinterpret_cast<float4*>(&Bs[buf][lk][lc]) = b;
    };
    load(0, 0);
    __syncthreads();
    int buf = 0;
    for (int64_t n0 = 0; n0 < N; n0 += TK) {
        if (n0 + TK < N) load(buf ^ 1, n0 + TK);
#pragma unroll
        for (int k = 0; k < TK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8 + 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 8]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 8 + 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
        buf ^= 1;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t r = i0 + ty * 8 + i;
        if (r >= G) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int64_t c = j0 + tx * 8 + j;
            if (c < G) C[r * G + c] = acc[i][j];
        }
    }
}

// excl[s][g / 32] bit g % 32 = gene g is a target of sub-network s
__global__ void mark_targets_kernel(const int32_t* __restrict__ targ, int S, int O, int64_t words, uint32_t* __restrict__ excl) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= (int64_t)S * O) return;
    const int s = (int)(i / O);
    const int32_t g = targ[i];
    atomicOr(excl + (int64_t)s * words + (g >> 5), 1u << (g & 31));
}

// one warp per target (s, o): the ntop largest |C[target][cand[j]]| over candidates j not excluded for s
constexpr int MAX_TOP = 8;
__global__ void topk_kernel(const float* __restrict__ C, int64_t G, const int32_t* __restrict__ cand, int64_t n_cand,
                            const int32_t* __restrict__ targ, int S, int O, const uint32_t* __restrict__ excl, int64_t words,
                            int ntop, int32_t* __restrict__ top_out, float* __restrict__ val_out) {
    const int64_t row = blockIdx.x * (int64_t)(blockDim.x / 32) + threadIdx.x / 32;
    if (row >= (int64_t)S * O) return;
    const int lane = threadIdx.x & 31;
    const int s = (int)(row / O);
    const float* crow = C + (int64_t)targ[row] * G;
    const uint32_t* ex = excl + (int64_t)s * words;
    float bv[MAX_TOP]; int64_t bj[MAX_TOP];
#pragma unroll
    for (int k = 0; k < MAX_TOP; ++k) { bv[k] = -1.f; bj[k] = INT64_MAX; }
    for (int64_t j = lane; j < n_cand; j += 32) {
        const int32_t g = cand[j];
        if ((ex[g >> 5] >> (g & 31)) & 1u) continue;
        float v = fabsf(crow[g]);
        if (!(v == v)) v = 0.f;                               // NaN -> 0 (reference fillna(0))
        if (v > bv[ntop - 1]) {                               // strict: a later equal value never displaces an earlier one
            int k = ntop - 1;
            while (k > 0 && v > bv[k - 1]) { bv[k] = bv[k - 1]; bj[k] = bj[k - 1]; --k; }
            bv[k] = v; bj[k] = j;
        }
    }
    // merge the 32 sorted lists: ntop rounds of (max value, then min position)
    int head = 0;
    for (int k = 0; k < ntop; ++k) {
        float v = head < ntop ? bv[head] : -1.f;
        int64_t j = head < ntop ? bj[head] : INT64_MAX;
        float wv = v; int64_t wj = j;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, wv, off);
            const int64_t oj = __shfl_xor_sync(0xffffffffu, wj, off);
            if (ov > wv || (ov == wv && oj < wj)) { wv = ov; wj = oj; }
        }
        if (wj == j && wv == v && j != INT64_MAX) ++head;     // this lane's head was taken
        if (lane == 0) {
            top_out[row * ntop + k] = wj == INT64_MAX ? -1 : (int32_t)wj;
            if (val_out) val_out[row * ntop + k] = wv;
        }
    }
}

}  // namespace

extern "C" {

const char* di_corr_last_error(void) { return g_err.c_str(); }

int di_corr_topk(int32_t device, const float* raw, int64_t n_cells, int64_t n_genes, const int32_t* cand, int64_t n_cand,
                 const int32_t* targ, int32_t n_subnets, int32_t sub_outputdim, int32_t ntop, int32_t* top_out,
                 float* val_out, float* device_ms_out) {
    struct R { R() { nvtxRangePushA("di_corr_topk"); } ~R() { nvtxRangePop(); } } nvtx_range;
    int rc = DI_OK;
    if (!raw || !cand || !targ || !top_out || n_cells <= 1 || n_genes <= 0 || n_cand <= 0 || n_subnets <= 0 ||
        sub_outputdim <= 0 || ntop <= 0 || ntop > MAX_TOP) { g_err = "di_corr_topk: bad arguments"; return DI_ERR_ARG; }
    for (int64_t j = 0; j < n_cand; ++j)
        if (cand[j] < 0 || cand[j] >= n_genes) { g_err = "di_corr_topk: candidate column out of range"; return DI_ERR_ARG; }
    const int64_t rows = (int64_t)n_subnets * sub_outputdim;
    for (int64_t i = 0; i < rows; ++i)
        if (targ[i] < 0 || targ[i] >= n_genes) { g_err = "di_corr_topk: target column out of range"; return DI_ERR_ARG; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
        g_err = "di_corr_topk: no such CUDA device (there is no CPU fallback)"; return DI_ERR_CUDA;
    }
    const int64_t N = n_cells, G = (n_genes + 3) / 4 * 4;    // pad genes to a multiple of 4 (float4 loads)
    const int64_t words = (G + 31) / 32;
    float *dZ = nullptr, *dC = nullptr, *dval = nullptr;
    double *dmean = nullptr, *dss = nullptr;
    int32_t *dcand = nullptr, *dtarg = nullptr, *dtop = nullptr;
    uint32_t* dexcl = nullptr;
    cudaStream_t st = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    {
        CK(cudaSetDevice(device));
        CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        CK(cudaMalloc((void**)&dZ, (size_t)N * G * sizeof(float)));
        CK(cudaMalloc((void**)&dC, (size_t)G * G * sizeof(float)));
        CK(cudaMalloc((void**)&dmean, (size_t)G * sizeof(double)));
        CK(cudaMalloc((void**)&dss, (size_t)G * sizeof(double)));
        CK(cudaMalloc((void**)&dcand, (size_t)n_cand * sizeof(int32_t)));
        CK(cudaMalloc((void**)&dtarg, (size_t)rows * sizeof(int32_t)));
        CK(cudaMalloc((void**)&dtop, (size_t)rows * ntop * sizeof(int32_t)));
        CK(cudaMalloc((void**)&dval, (size_t)rows * ntop * sizeof(float)));
        CK(cudaMalloc((void**)&dexcl, (size_t)n_subnets * words * sizeof(uint32_t)));
        if (G != n_genes) CK(cudaMemsetAsync(dZ, 0, (size_t)N * G * sizeof(float), st));
        CK(cudaMemcpy2DAsync(dZ, (size_t)G * sizeof(float), raw, (size_t)n_genes * sizeof(float), (size_t)n_genes * sizeof(float),
                             (size_t)N, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(dcand, cand, (size_t)n_cand * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(dtarg, targ, (size_t)rows * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        CK(cudaMemsetAsync(dmean, 0, (size_t)G * sizeof(double), st));
        CK(cudaMemsetAsync(dss, 0, (size_t)G * sizeof(double), st));
        CK(cudaMemsetAsync(dexcl, 0, (size_t)n_subnets * words * sizeof(uint32_t), st));
        CK(cudaEventRecord(e0, st));
        const dim3 cgrid((unsigned)((G + 127) / 128), (unsigned)((N + ROWS_PER_BLOCK - 1) / ROWS_PER_BLOCK));
        colsum_kernel<<<cgrid, 128, 0, st>>>(dZ, N, G, nullptr, dmean);
        finish_mean_kernel<<<(unsigned)((G + 255) / 256), 256, 0, st>>>(dmean, G, N);
        colsum_kernel<<<cgrid, 128, 0, st>>>(dZ, N, G, dmean, dss);
        standardise_kernel<<<cgrid, 128, 0, st>>>(dZ, N, G, dmean, dss);
        const unsigned tiles = (unsigned)((G + TM - 1) / TM);
        gram_kernel<<<dim3(tiles, tiles), 256, 0, st>>>(dZ, N, G, dC);
        mark_targets_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(dtarg, n_subnets, sub_outputdim, words, dexcl);
        topk_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(dC, G, dcand, n_cand, dtarg, n_subnets, sub_outputdim, dexcl,
                                                                words, ntop, dtop, dval);
        CK(cudaEventRecord(e1, st));
        CK(cudaMemcpyAsync(top_out, dtop, (size_t)rows * ntop * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        if (val_out) CK(cudaMemcpyAsync(val_out, dval, (size_t)rows * ntop * sizeof(float), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        CK(cudaGetLastError());
        if (device_ms_out) CK(cudaEventElapsedTime(device_ms_out, e0, e1));
    }
done:
    cudaFree(dZ); cudaFree(dC); cudaFree(dmean); cudaFree(dss); cudaFree(dcand); cudaFree(dtarg); cudaFree(dtop);
    cudaFree(dval); cudaFree(dexcl);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (st) cudaStreamDestroy(st);
    return rc;
}

}  // extern "C"
