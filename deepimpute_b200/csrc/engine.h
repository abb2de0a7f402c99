// Engine state and kernel-launch interface shared by di_api.cu, kernels_simt.cu and kernels_tc.cu.
//
// Data layout in HBM (all fp32, row-major, every pitch a multiple of 32 floats):
//   norm   [N][G]            log1p(counts), uploaded once (reference multinet.py:217)
//   X*     [rows][PT]        packed predictor values: sub-network s owns columns coff_s .. coff_s+Pp_s
//   Y*     [rows][S*Op]      packed target values:    sub-network s owns columns s*Op .. s*Op+O
//   W1     [PT][Hp]          sub-network s owns rows  coff_s .. coff_s+Pp_s   (Keras layout W[in][out])
//   W2     [S*Hp][Op]        sub-network s owns rows  s*Hp .. s*Hp+Hp
//   b1     [S][Hp],  b2 [S][Op];  Adam moments m*, v* mirror the weights
//   Hact   [rows][S*Hp]      hidden activations after relu (+dropout when training)
//   DZ2    [B][S*Op], DZ1 [B][S*Hp]   gradients w.r.t. the pre-activations
// Padding columns/rows hold zeros and stay zero under Adam (g = 0 -> m = v = 0 -> update 0).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include <map>

#include "../../include/deepimpute_b200.h"
#include "common.cuh"

namespace di {

struct StepArgs {
    const float* X;        // [rows][ldx] packed predictors of the batch
    const float* Y;        // [rows][ldy] packed targets of the batch
    int64_t ldx, ldy;
    int64_t row0;          // first row of the batch inside X / Y
    int32_t n_valid;       // rows of the batch that are real cells (<= B); the rest are zero padding
    uint32_t step;         // global optimiser step (dropout counter)
    AdamParams adam;
};

struct Engine {
    di_config cfg{};
    int S = 0, H = 0, O = 0, B = 0, Bp = 0, Hp = 0, Op = 0;
    std::vector<int> P, Pp, gid;
    std::vector<int64_t> coff;
    int64_t PT = 0;                 // sum of padded predictor counts
    int maxPp = 0;
    SubnetDesc* d_desc = nullptr;   // [S]

    int64_t N = 0, G = 0;
    float* d_norm = nullptr;
    int32_t* d_pred_cols = nullptr; // [PT]   gene column of every packed X column, -1 = padding
    int32_t* d_targ_cols = nullptr; // [S*Op] gene column of every packed Y column, -1 = padding
    bool have_partition = false;
    std::vector<int32_t> h_targ;    // host copy of the target columns [S*O] (slot -> gene table of di_impute)
    // raw counts as uploaded by di_upload_counts (fp32 or fp64), kept for the restore / max policies of di_impute
    void* d_raw = nullptr;
    int raw_dtype = -1;             // DI_DTYPE_F32 / DI_DTYPE_F64, -1 = the matrix came through di_upload_matrix
    double raw_max = 0.0;           // largest count of the matrix (the clamp of multinet.py:291 is 2*log1p of it)
    unsigned long long* d_raw_max = nullptr;
    int32_t *d_gene_off = nullptr, *d_gene_slots = nullptr;   // CSR gene -> prediction columns (di_impute)
    void* d_imp[2] = {nullptr, nullptr};                      // [imp_rows][G] imputed chunk, double-buffered
    int64_t imp_rows = 0; size_t imp_bytes = 0;
    bool split_stale = false;       // partition changed since di_set_split: staged matrices must be refilled

    // weights, Adam moments and (DI_MATH_TF32X3) the residual twins of the weights live in ONE allocation, so that a
    // single L2 access-policy window can cover the whole optimiser state (kernels_tc.cu: l2_window)
    float* state_slab = nullptr;
    size_t state_bytes = 0;
    float *W1 = nullptr, *mW1 = nullptr, *vW1 = nullptr;
    float *b1 = nullptr, *mb1 = nullptr, *vb1 = nullptr;
    float *W2 = nullptr, *mW2 = nullptr, *vW2 = nullptr;
    float *b2 = nullptr, *mb2 = nullptr, *vb2 = nullptr;
    float *W1lo = nullptr, *W2lo = nullptr;       // W - trunc_tf32(W): rewritten by the ADAM kernel with every update
    int64_t adam_t = 0;

    int32_t *d_train_rows = nullptr, *d_test_rows = nullptr, *d_perm = nullptr;
    int64_t n_train = 0, n_test = 0, n_train_pad = 0, n_test_pad = 0;
    float *Xtr = nullptr, *Ytr = nullptr, *Xte = nullptr, *Yte = nullptr;

    float *Xstep = nullptr, *Ystep = nullptr;     // [B][PT], [B][S*Op]   explicit-batch step
    int32_t* d_step_rows = nullptr;               // [B]
    float *Hact = nullptr, *DZ2 = nullptr, *DZ1 = nullptr;
    // DI_MATH_TF32X3 only: residual twins a_lo = a - trunc_tf32(a) of the operands of the weight-gradient GEMMs,
    // written by the epilogue (h, dz2, dz1) or the staging gather (X) that produces the value itself
    float *Hlo = nullptr, *DZ2lo = nullptr, *DZ1lo = nullptr, *Xtr_lo = nullptr, *Xstep_lo = nullptr;
    float *Xte_lo = nullptr, *Xchunk_lo = nullptr, *Hchunk_lo = nullptr;   // twins of the inference operands

    int infer_tile = 128;                         // cells per CTA of the inference forward (UMMA N): 128 or 256
    int64_t chunk_rows = 0;                       // inference chunk (multiple of infer_tile)
    float *Xchunk = nullptr, *Hchunk = nullptr, *Ochunk = nullptr, *OchunkB = nullptr;
    float* Ochunk2[2] = {nullptr, nullptr};       // {Ochunk, OchunkB}: double buffer of the direct D2H path
    cudaEvent_t ev_fwd[2] = {nullptr, nullptr};
    int32_t* d_chunk_rows = nullptr;
    float* h_pinned[2] = {nullptr, nullptr};      // D2H staging for di_predict
    cudaEvent_t ev_pinned[2] = {nullptr, nullptr};

    double* d_loss = nullptr;                     // [2]: 0 = training raw sum, 1 = validation raw sum

    cudaStream_t stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;
    float last_ms = 0.f;
    int64_t launches = 0;
    bool profiling = false;
    std::map<std::string, std::pair<double, int64_t>> kernel_ms;   // name -> (total ms, count)
    struct PendingTimer { const char* name; cudaEvent_t a, b; };
    std::vector<PendingTimer> pending_timers;                      // recorded, not yet read back
    std::vector<cudaEvent_t> event_pool;
    std::string err;

    void* tc = nullptr;                           // tensor-core path state (tensor maps), kernels_tc.cu
};

// ---- staging ------------------------------------------------------------------------------------------------
// out[i][j] = norm[row(src(i))][cols[j]] when src(i) is valid and cols[j] >= 0, else 0, for i in [0, n_out);
//   row(k) = rows[perm ? perm[k] : k]   (rows == nullptr: row(k) = first_row + k)
//   batch == 0: src(i) = i, valid iff i < n_valid
//   batch  > 0: output rows come in groups of batch_pitch holding `batch` source rows each (the per-step batches
//               of the training matrices): src(i) = (i / batch_pitch) * batch + i % batch_pitch, valid iff
//               i % batch_pitch < batch and src(i) < n_valid
//   out_lo != nullptr: also writes the TF32 residual x - trunc_tf32(x) of every value to out_lo (same shape)
void launch_gather(Engine& e, const int32_t* rows, const int32_t* perm, int64_t first_row, int64_t n_out,
                   int64_t n_valid, const int32_t* cols, int64_t width, float* out, int batch = 0, int batch_pitch = 0,
                   float* out_lo = nullptr);

// Predictor AND target matrices of the same rows in one pass: every source row of norm is staged once in shared
// memory (coalesced read), both packed outputs are gathered from there (coalesced writes).  X -> [n_out][PT] (+ its
// TF32 residual twin when X_lo != nullptr), Y -> [n_out][S*Op].  Row mapping as in launch_gather.
void launch_gather_xy(Engine& e, const int32_t* rows, const int32_t* perm, int64_t n_out, int64_t n_valid,
                      int batch, int batch_pitch, float* X, float* X_lo, float* Y);

// ---- counts -> log1p and the fused tail of predict (impute.cu) -----------------------------------------------
// norm[i] = (float)log1p((double)raw[i]) for n values (multinet.py:217, :271) and *max_bits = bits of the largest
// non-negative count as a double (atomicMax on the ordered bit pattern); dtype = DI_DTYPE_*
void launch_counts_to_norm(Engine& e, const void* raw, int dtype, float* norm, int64_t n, unsigned long long* max_bits);
// out[r][g] for r in [0, rows), g in [0, G): the imputed count of multinet.py:282-303 (duplicate-slot mean, overflow
// clamp, expm1, restore / max policy).  pred: [rows][ld_pred] predictions of these rows; raw: first of the rows in
// the resident count matrix; out_dtype = DI_DTYPE_*
void launch_impute(Engine& e, const float* pred, int64_t ld_pred, int64_t row0, int64_t rows, double clamp,
                   int policy, int out_dtype, void* out);

// ---- fp32 CUDA-core path (kernels_simt.cu) -------------------------------------------------------------------
void simt_train_step(Engine& e, const StepArgs& a);
void simt_adam_only(Engine& e, const StepArgs& a);   // W1/W2 <- Adam(exact fp32 dW); no bias update
// forward for `rows` rows of X (multiple of 64): Hbuf [rows][S*Hp] scratch; if Y != nullptr accumulates the raw
// validation sum into d_loss[1]; if out != nullptr writes yhat to out[rows][ld_out] (column s*O + o).
void simt_forward(Engine& e, const float* X, int64_t ldx, int64_t rows, int64_t n_valid, float* Hbuf,
                  const float* Y, int64_t ldy, float* out, int64_t ld_out);

// ---- tcgen05 / TMA path (kernels_tc.cu) ----------------------------------------------------------------------
unsigned int tc_take_timeout_word();   // kernels_tc.cu: non-zero if an mbarrier wait timed out since the last call
bool tc_available();                // false while the tensor-core kernels are not part of the build
bool tc_init(Engine& e);            // builds tensor maps; false + e.err on failure
void tc_destroy(Engine& e);
bool tc_rebind(Engine& e);          // after (re)allocation of X/Y buffers
void tc_weights_changed(Engine& e, int s);   // di_set_weights wrote sub-network s: refresh the residual twins of W1 / W2
const char* tc_describe(Engine& e);  // one line: which kernels / knobs this handle runs with (di_describe)
int64_t tc_fallbacks(Engine& e);     // times the epoch graph was unavailable and the steps were issued one by one
void tc_train_step(Engine& e, const StepArgs& a, int which_x);   // which_x: 0 = Xtr/Ytr, 1 = Xstep/Ystep
// All optimiser steps of one epoch over the staged training matrices as ONE graph launch on e.stream (built on first
// use, rebuilt when the split changes).  lr_t[i] = Adam's bias-corrected rate of step first_step + i.  Returns false
// when the graph path is unavailable (then the caller issues the steps one by one with tc_train_step).
bool tc_train_epoch_graph(Engine& e, int64_t first_step, const float* lr_t, int64_t n_steps);
void tc_forward(Engine& e, int which_x, int64_t row0, int64_t rows, int64_t n_valid, bool with_loss,
                float* out, int64_t ld_out);                     // which_x: 2 = Xte/Yte, 3 = Xchunk

// bias gradients + Adam for b1 and b2 (shared by both paths): db2 = sum_b DZ2, db1 = sum_b DZ1
void launch_bias_adam(Engine& e, const AdamParams& adam);

void count_launch(Engine& e, const char* name);
void resolve_timers(Engine& e);   // read back all recorded kernel timings (call after a stream sync)
struct KernelTimer {   // CUDA-event timing of one launch on e.stream when profiling is on (deferred read-back)
    Engine& e; const char* name; cudaEvent_t a = nullptr, b = nullptr;
    KernelTimer(Engine& e_, const char* n);
    ~KernelTimer();
};

}  // namespace di
