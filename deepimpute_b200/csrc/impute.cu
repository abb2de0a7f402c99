// Element-wise ends of the path, both bound by HBM bandwidth (no reuse, no tensor-core work):
//
//   counts_to_norm   raw counts -> log1p as fp32          (reference multinet.py:217 fit, :271 predict)
//   impute_kernel    the tail of MultiNet.predict          (reference multinet.py:282-303)
//
// The reference does both in pandas/numpy through several N x G float64 temporaries.  Here each is ONE pass:
// counts_to_norm reads 4 (8) bytes and writes 4 per value; impute_kernel reads the count (4/8 B) and -- for target
// genes only -- the predicted slots (4 B each, an in-row gather served by L2), and writes the imputed count once
// (8 B as float64 like the reference's DataFrame, or 4 B).  Arithmetic is float64 exactly where numpy's is.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <algorithm>

#include "engine.h"

namespace di {

namespace {

constexpr int kThreads = 256;

// ---------------------------------------------------------------------------------------------- counts -> norm
template <typename T> struct Vec;
template <> struct Vec<float>  { using type = float4;  static constexpr int n = 4; };
template <> struct Vec<double> { using type = double2; static constexpr int n = 2; };

// float64 log1p costs ~100 fp64 instructions; counts are overwhelmingly small integers, so each block first tabulates
// (float)log1p((double)k) for k < kTab with the very same expression and looks those values up (bit-identical
// results, the kernel becomes bandwidth-bound); anything else -- large, fractional, negative, NaN -- is computed.
constexpr int kTab = 1024;
__device__ __forceinline__ float norm_direct(double c) { return (float)log1p(c); }
__device__ __forceinline__ float norm_of(double c, const float* tab) {
    const int k = (int)c;                                  // saturating / NaN -> 0 conversion, checked below
    return ((unsigned)k < (unsigned)kTab && (double)k == c) ? tab[k] : norm_direct(c);
}

template <typename T>
__global__ void __launch_bounds__(kThreads) counts_to_norm_kernel(const T* __restrict__ raw, float* __restrict__ norm,
                                                                  int64_t n, unsigned long long* max_bits) {
    using V = typename Vec<T>::type;
    constexpr int VN = Vec<T>::n;
    __shared__ float tab[kTab];
    for (int k = threadIdx.x; k < kTab; k += kThreads) tab[k] = norm_direct((double)k);
    __syncthreads();
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    const int64_t nvec = n / VN;
    double best = 0.0;
    const V* rv = reinterpret_cast<const V*>(raw);
    for (int64_t i = tid; i < nvec; i += nthreads) {
        const V v = __ldcs(rv + i);                      // streamed: every value is read exactly once
        if constexpr (VN == 4) {
            const double a = v.x, b = v.y, c = v.z, d = v.w;
            best = fmax(best, fmax(fmax(a, b), fmax(c, d)));
            __stcs(reinterpret_cast<float4*>(norm) + i,
                   make_float4(norm_of(a, tab), norm_of(b, tab), norm_of(c, tab), norm_of(d, tab)));
        } else {
            best = fmax(best, fmax((double)v.x, (double)v.y));
            __stcs(reinterpret_cast<float2*>(norm) + i, make_float2(norm_of(v.x, tab), norm_of(v.y, tab)));
        }
    }
    for (int64_t i = nvec * VN + tid; i < n; i += nthreads) {
        const double c = (double)raw[i];
        best = fmax(best, c);
        norm[i] = norm_of(c, tab);
    }
    // non-negative doubles order like their bit patterns; fmax drops NaN and best starts at 0
    unsigned long long bits = (unsigned long long)__double_as_longlong(best);
    for (int off = 16; off; off >>= 1) {
        const unsigned long long o = __shfl_xor_sync(0xffffffffu, bits, off);
        bits = o > bits ? o : bits;
    }
    if ((threadIdx.x & 31) == 0 && bits) atomicMax(max_bits, bits);
}

// ------------------------------------------------------------------------------------------------------ impute
// One thread per (cell, gene).  gene_ent[g] = (first, count): the prediction columns gene_slots[first .. first+count)
// all predict gene g (targets are padded with repeated genes, multinet.py:334-342); count == 0: g was not imputed.
template <typename TRaw, typename TOut>
__global__ void __launch_bounds__(kThreads) impute_kernel(const float* __restrict__ pred, int64_t ld_pred,
                                                          const int2* __restrict__ gene_ent,
                                                          const int32_t* __restrict__ gene_slots,
                                                          const TRaw* __restrict__ raw, int64_t G, int64_t rows,
                                                          double clamp, int policy, TOut* __restrict__ out) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    const int2 ent = gene_ent[g];
    for (int64_t r = blockIdx.y; r < rows; r += gridDim.y) {
        const TRaw c = __ldcs(raw + r * G + g);
        if (policy == DI_POLICY_RESTORE && c > (TRaw)0) {      // observed counts are kept whatever was predicted (:295-298)
            __stcs(out + r * G + g, (TOut)c);
            continue;
        }
        double v;
        if (ent.y == 0) {
            // not a target: the reference carries float64 log1p(raw) through expm1 (multinet.py:286-287, :293);
            // a zero count stays exactly zero through log1p, the clamp test and expm1
            v = (c == (TRaw)0) ? 0.0 : log1p((double)c);
        } else {
            // groupby-mean of float32 columns (:284) as pandas computes it: NaN entries are skipped, the sum is a
            // Kahan sum in float32 over the columns in ascending order, divided by the count in float32
            const float* p = pred + r * ld_pred;
            float sum = 0.f, comp = 0.f;
            int n = 0;
            for (int k = 0; k < ent.y; ++k) {
                const float val = p[gene_slots[ent.x + k]];
                if (val == val) {
                    ++n;
                    const float y = val - comp;
                    const float t = sum + y;
                    comp = (t - sum) - y;
                    if (comp != comp) comp = 0.f;
                    sum = t;
                }
            }
            v = n ? (double)(sum / (float)n) : (double)NAN;
        }
        if (v > clamp || isnan(v)) v = 0.0;              // "to prevent overflow" (:291)
        if (v != 0.0) v = expm1(v);                      // back to counts (:293); expm1(0) = 0
        if (policy == DI_POLICY_RESTORE) { if (c > (TRaw)0) v = (double)c; }            // :295-298
        else if (policy == DI_POLICY_MAX) { if ((double)c > v) v = (double)c; }         // :299-302
        __stcs(out + r * G + g, (TOut)v);
    }
}

template <typename TRaw>
void impute_dispatch(Engine& e, const float* pred, int64_t ld_pred, int64_t row0, int64_t rows, double clamp,
                     int policy, int out_dtype, void* out) {
    const int2* ent = reinterpret_cast<const int2*>(e.d_gene_off);
    const TRaw* raw = static_cast<const TRaw*>(e.d_raw) + row0 * e.G;
    // gene tiles x row lanes: about 16 resident blocks per SM in total, every thread walks rows with stride grid.y
    const int64_t gx = (e.G + kThreads - 1) / kThreads;
    dim3 grid((unsigned)gx, (unsigned)std::max<int64_t>(1, std::min<int64_t>(rows, (148 * 16 + gx - 1) / gx)));
    if (out_dtype == DI_DTYPE_F64)
        impute_kernel<TRaw, double><<<grid, kThreads, 0, e.stream>>>(pred, ld_pred, ent, e.d_gene_slots, raw, e.G, rows,
                                                                     clamp, policy, static_cast<double*>(out));
    else
        impute_kernel<TRaw, float><<<grid, kThreads, 0, e.stream>>>(pred, ld_pred, ent, e.d_gene_slots, raw, e.G, rows,
                                                                    clamp, policy, static_cast<float*>(out));
}

}  // namespace

void launch_counts_to_norm(Engine& e, const void* raw, int dtype, float* norm, int64_t n, unsigned long long* max_bits) {
    if (n <= 0) return;
    KernelTimer t(e, "log1p");
    const int64_t per_thread = 16;
    const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>((n + kThreads * per_thread - 1) / (kThreads * per_thread),
                                                                           148 * 16));
    if (dtype == DI_DTYPE_F64)
        counts_to_norm_kernel<double><<<grid, kThreads, 0, e.stream>>>(static_cast<const double*>(raw), norm, n, max_bits);
    else
        counts_to_norm_kernel<float><<<grid, kThreads, 0, e.stream>>>(static_cast<const float*>(raw), norm, n, max_bits);
    count_launch(e, "log1p");
}

void launch_impute(Engine& e, const float* pred, int64_t ld_pred, int64_t row0, int64_t rows, double clamp,
                   int policy, int out_dtype, void* out) {
    if (rows <= 0) return;
    KernelTimer t(e, "impute");
    if (e.raw_dtype == DI_DTYPE_F64) impute_dispatch<double>(e, pred, ld_pred, row0, rows, clamp, policy, out_dtype, out);
    else impute_dispatch<float>(e, pred, ld_pred, row0, rows, clamp, policy, out_dtype, out);
    count_launch(e, "impute");
}

}  // namespace di
