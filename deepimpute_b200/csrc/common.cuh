// Shared device helpers of the DeepImpute B200 engine: Philox dropout masks, activations, Adam update.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace di {

// ---------------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011).  Defines the dropout mask that Keras' Dropout(rate, seed) layer
// (reference multinet.py:139-141) would draw from TensorFlow's RNG; same definition as oracle/philox.py:
//   element (b, j) of sub-network s at optimiser step t:
//   words = philox(counter = (j, b >> 2, s, t), key = (seed_lo, seed_hi));  keep = words[b & 3] >= thresh
// ---------------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ void philox_round(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3,
                                                      uint32_t k0, uint32_t k1) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
}

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                       uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        philox_round(c0, c1, c2, c3, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// keep-bits of rows 4*bq .. 4*bq+3 for hidden unit j
__device__ __forceinline__ void dropout_words(uint32_t j, uint32_t bq, uint32_t s, uint32_t step,
                                              uint64_t seed, uint32_t w[4]) {
    philox4x32_10(j, bq, s, step, (uint32_t)seed, (uint32_t)(seed >> 32), w);
}

// ---------------------------------------------------------------------------------------------------------
// softplus / sigmoid of the output layer (Dense(O, activation="softplus"), multinet.py:145) and the wMSE
// gradient through it (multinet.py:36-41):  L = mean(y (y - yhat)^2)  ->  dL/dz2 = 2 y (yhat - y) sigmoid(z2) / (n O)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float softplus_f(float z) {
    return fmaxf(z, 0.0f) + log1pf(expf(-fabsf(z)));
}
__device__ __forceinline__ float sigmoid_f(float z) {
    return 1.0f / (1.0f + expf(-z));
}

// Keras/TensorFlow Adam (ResourceApplyAdam): lr_t carries the bias correction, epsilon is added to sqrt(v).
struct AdamParams {
    float lr_t;            // lr * sqrt(1 - b2^t) / (1 - b1^t)
    float one_minus_b1;
    float one_minus_b2;
    float eps;
};

__device__ __forceinline__ void adam_update(float g, float& w, float& m, float& v, const AdamParams& a) {
    m = m + (g - m) * a.one_minus_b1;
    v = v + (g * g - v) * a.one_minus_b2;
    w = w - a.lr_t * m / (sqrtf(v) + a.eps);
}

// Same update with approximate reciprocal / square root (2 MUFU ops, ~1e-7 relative): used by the tensor-core
// path, whose TF32 operands already carry a 1e-3 relative error, where the IEEE sequences would make the weight
// epilogue instruction-bound instead of HBM-bound.
__device__ __forceinline__ void adam_update_fast(float g, float& w, float& m, float& v, const AdamParams& a) {
    m = m + (g - m) * a.one_minus_b1;
    v = v + (g * g - v) * a.one_minus_b2;
    float r;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(v));
    w = w - a.lr_t * __fdividef(m, r + a.eps);
}

// Per-sub-network geometry, resident in device memory.  Every row pitch is a multiple of 32 floats (128 B)
// so that TMA tiles and 128-byte swizzle atoms never straddle sub-networks.
struct SubnetDesc {
    int64_t coff;    // first column of this sub-network in the packed X matrix == first row in packed W1
    int32_t P;       // predictors (K of layer 1)
    int32_t Pp;      // padded to a multiple of 32
    int32_t gid;     // global sub-network number (keys the dropout stream; differs from the local index when the
                     // sub-networks of one model are sharded over several GPUs)
    int32_t pad_;
};

}  // namespace di
