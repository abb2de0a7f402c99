/*
 * deepimpute_b200.h -- C ABI of the B200-native DeepImpute hot path.
 *
 * The reference (lanagarmire/deepimpute) has no native boundary: `MultiNet.fit/predict` call Keras directly
 * (reference deepimpute/multinet.py:226-253, :276-280).  This header is the boundary a maintainer would bind
 * instead (ctypes stub in INTEGRATION.md).  Every entry point names the reference call it replaces.
 *
 * Conventions
 *  - plain C types only; every pointer is a HOST pointer to C-contiguous memory owned by the caller unless the
 *    name starts with `d_` (device pointer owned by the caller, e.g. a torch tensor's data_ptr());
 *  - the library copies during the call and never keeps a host pointer;
 *  - every function returns 0 on success or a DI_ERR_* code; di_last_error() gives the message;
 *  - there is no CPU fallback: without a CUDA device di_create() fails with DI_ERR_CUDA;
 *  - a handle is used by one thread at a time; it owns one device, one stream and all device memory.
 *
 * Shapes: S sub-networks, sub-network s has P[s] predictor genes, H hidden units, O output genes
 * (reference multinet.py:132-146).  Weights use the Keras layout W[in][out], row-major, fp32.
 */
#ifndef DEEPIMPUTE_B200_H
#define DEEPIMPUTE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DI_OK            0
#define DI_ERR_ARG       1   /* bad argument / wrong call order            */
#define DI_ERR_CUDA      2   /* CUDA runtime or driver error, or no device */
#define DI_ERR_OOM       3   /* device allocation failed                   */
#define DI_ERR_NUMERIC   4   /* non-finite loss                            */

#define DI_MATH_FP32     0   /* CUDA-core fp32 FFMA kernels (bit-for-bit fp32 products)          */
#define DI_MATH_TF32     1   /* tcgen05 kind::tf32 tensor-core kernels, fp32 accumulate in TMEM; operands are
                                TRUNCATED to TF32 by the tensor core (about 1e-3 relative per product)        */
#define DI_MATH_TF32X3   2   /* same tensor cores, EVERY GEMM of the step error-compensated (a b ~ a_hi b_hi + a_hi b_lo +
                                a_lo b_hi with a_hi = trunc_tf32(a), a_lo = a - a_hi): the three forward / backward GEMMs
                                and both weight-gradient GEMMs; fp32-level products (dropped a_lo b_lo term: 2^-22).
                                The default of the Python layer; DESIGN.md section 4 states the measured tolerance */

#define DI_DTYPE_F32     0   /* element type of a count matrix / of the imputed output */
#define DI_DTYPE_F64     1

#define DI_POLICY_NONE     0 /* MultiNet.predict(policy=...) (multinet.py:295-302): keep the network's value        */
#define DI_POLICY_RESTORE  1 /* "restore": observed counts > 0 are kept, only zeros are imputed (the default)        */
#define DI_POLICY_MAX      2 /* "max": the larger of the observed count and the imputed one                          */

typedef struct di_handle di_handle;

/* Hyper-parameters fixed at build time.  Replaces MultiNet.build + model.compile(Adam(lr), wMSE)
 * (reference multinet.py:126-167; defaults :67-79, :99-103; Keras Adam defaults beta1 .9 beta2 .999 eps 1e-7). */
typedef struct di_config {
    int32_t  n_subnets;       /* S                                                       */
    int32_t  hidden;          /* H: neurons of the hidden Dense(relu)                    */
    int32_t  sub_outputdim;   /* O: neurons of the output Dense(softplus), 512           */
    int32_t  batch_size;      /* B, 64                                                   */
    float    learning_rate;
    float    beta1, beta2, epsilon;
    float    dropout_rate;    /* r of Dropout(r) between the two Dense layers            */
    uint64_t seed;            /* keys the dropout masks (Philox4x32-10)   */
    int32_t  math_mode;       /* DI_MATH_*                                               */
    int32_t  device;          /* CUDA ordinal                                            */
} di_config;

/* Create an engine for S sub-networks with n_pred[s] predictors each.  Weights start at zero; call
 * di_set_weights.  Replaces MultiNet.build(inputdims) (multinet.py:126, called at :226). */
int di_create(di_handle** out, const di_config* cfg, const int32_t* n_pred);
void di_destroy(di_handle* h);
const char* di_last_error(const di_handle* h);   /* h may be NULL: error of the last failed di_create */

/* Global numbers of the S sub-networks of this handle (default 0..S-1).  When one model is sharded over several
 * GPUs each handle owns a subset of the branches of multinet.py:132-148; the global number keys the dropout
 * stream so that results do not depend on the sharding. */
int di_set_subnet_ids(di_handle* h, const int32_t* ids);

/* Normalised expression matrix norm[N][G] (= log1p(raw) as fp32, multinet.py:217 / :271), copied to the device.
 * pred_idx: concatenated predictor gene columns, sub-network s owns pred_idx[pred_off[s] .. pred_off[s+1]);
 * targ_idx[S][O]: target gene columns.  Replaces the S x 4 pandas gathers X_train.. Y_test = norm.loc[cells, genes]
 * (multinet.py:231-235, :273-274): the gathers run on the device from these index tables. */
int di_upload_matrix(di_handle* h, const float* norm, int64_t n_cells, int64_t n_genes);
int di_set_partition(di_handle* h, const int32_t* pred_idx, const int64_t* pred_off, const int32_t* targ_idx);

/* Same as di_upload_matrix but from RAW counts raw[N][G] (dtype DI_DTYPE_F32 or DI_DTYPE_F64): the device computes
 * norm = (float)log1p((double)raw) -- np.log1p(raw).astype(np.float32), multinet.py:217 -- in one pass and keeps the
 * counts resident for di_impute.  Replaces the host-side float64 log1p of fit (:217) and predict (:271). */
int di_upload_counts(di_handle* h, const void* raw, int32_t dtype, int64_t n_cells, int64_t n_genes);

/* Row numbers (into norm) of training and held-out cells (multinet.py:228-229). */
int di_set_split(di_handle* h, const int32_t* train_rows, int64_t n_train,
                 const int32_t* test_rows, int64_t n_test);

/* Weights of sub-network s: W1[P][H], b1[H], W2[H][O], b2[O].  set also zeroes its Adam moments.
 * Replace Keras initialisation / model.get_weights / load_weights (multinet.py:105-124). */
int di_set_weights(di_handle* h, int32_t s, const float* W1, const float* b1, const float* W2, const float* b2);
int di_get_weights(di_handle* h, int32_t s, float* W1, float* b1, float* W2, float* b2);
/* Adam state (first/second moments in the same shapes, and the shared step counter t). */
int di_get_adam_state(di_handle* h, int32_t s, float* mW1, float* vW1, float* mb1, float* vb1,
                      float* mW2, float* vW2, float* mb2, float* vb2, int64_t* t);

/* One epoch of model.fit (multinet.py:238-244): visits train_rows[perm[i]] in order, batches of B with the
 * partial last batch kept, dropout on, one Adam step per batch for all S sub-networks; then the validation
 * pass over the held-out cells in inference mode.  loss_out = sample-weighted mean over batches of the summed
 * per-sub-network wMSE (what Keras logs as `loss`), val_loss_out = sum_s mean_{cells,genes} y (y - yhat)^2
 * (`val_loss`).  `first_step` is the global index of the epoch's first optimiser step (keys dropout masks and,
 * +1, is Adam's t).  Early stopping stays with the caller (Keras EarlyStopping callback, multinet.py:242). */
int di_train_epoch(di_handle* h, const int32_t* perm, int64_t first_step, float* loss_out, float* val_loss_out);

/* One optimiser step on an explicit batch of rows (positions into norm), nrows <= B.  Same arithmetic as one
 * iteration of di_train_epoch; exists so parity tests can check a single step.  loss_out: summed wMSE. */
int di_train_step(di_handle* h, const int32_t* rows, int32_t nrows, int64_t step, float* loss_out);

/* Validation loss alone (second half of di_train_epoch). */
int di_validation_loss(di_handle* h, float* val_loss_out);

/* Inference forward, dropout off: out[n][S*O] (column s*O+o = target gene targ_idx[s][o]) for rows[0..n), or for
 * all cells in order when rows == NULL.  Replaces model.predict(X_list) + np.hstack (multinet.py:253, :278-280).
 * di_predict_device writes to a device buffer with leading dimension ld_out floats (so that a rank can write
 * its column block straight into the all-gather buffer of a multi-GPU run). */
int di_predict(di_handle* h, const int32_t* rows, int64_t n, float* out);
int di_predict_device(di_handle* h, const int32_t* rows, int64_t n, float* d_out, int64_t ld_out);

/* The whole of MultiNet.predict after the gathers (multinet.py:278-303), fused: inference forward of every sub-network,
 * mean over prediction columns that target the same gene (:284), imputed = log1p(raw) with the target genes replaced
 * (:286-287), values above 2*max(log1p(raw)) or NaN set to 0 (:291), expm1 (:293), then the policy (:295-302).
 * out[N][G] (DI_DTYPE_F64 like the reference's DataFrame, or DI_DTYPE_F32) may be pageable or page-locked host memory
 * (or device memory); chunks of cells are imputed while the previous chunk is copied out.  Needs di_upload_counts.
 *   slot_gene[n_slots]: gene column predicted by column k of the prediction matrix (negative = ignore the column);
 *                       NULL = the handle's own targets (targ_idx of di_set_partition, n_slots = S*O).
 *   d_pred:             NULL = run the forward pass of this handle; otherwise a DEVICE matrix [N][ld_pred] holding
 *                       n_slots prediction columns (e.g. the all-gathered blocks of a sub-network-sharded model),
 *                       and slot_gene is required.
 * Arithmetic is float64 where numpy's is (log1p/expm1 of CUDA's libm: within 2 ulp of glibc's). */
int di_impute(di_handle* h, int32_t policy, const int32_t* slot_gene, int64_t n_slots, const float* d_pred,
              int64_t ld_pred, int32_t out_dtype, void* out);

/* Per-gene mean and variance (ddof = 1, like pandas .var()) of raw[n_cells][n_genes] (dtype DI_DTYPE_F32 / _F64), in
 * float64: the inputs of both gene filters of fit -- the imputation ranking var / (1 + mean), multinet.py:191-192, and
 * the predictor-candidate filter std / mean > 0, multinet.py:22-24 -- which pandas computes in three single-core passes
 * over the frame.  Stand-alone like di_corr_topk (no handle; allocates and frees its own device memory).  Two passes
 * (mean, then squared deviations): values agree with pandas to ~1e-15 relative. */
int di_gene_stats(int32_t device, const void* raw, int32_t dtype, int64_t n_cells, int64_t n_genes, double* mean_out,
                  double* var_out, float* device_ms_out);
const char* di_gene_stats_last_error(void);

/* Predictor selection (the O(G^2 N) step of fit).  |Pearson r| between genes on RAW
 * counts raw[n_cells][n_genes] -- get_distance_matrix, multinet.py:20-34: abs(np.corrcoef(raw.T)), NaN -> 0 -- and for
 * every target gene targ[s][o] the ntop (<= 8) most correlated candidates that are not targets of sub-network s --
 * setPredictors, multinet.py:349-360: argsort(-|r|)[:, :ntop].  cand lists the candidate gene columns in the order
 * ties are broken (the reference visits them label-sorted); top_out[s][o][k] is a POSITION into cand (or -1 when
 * fewer than ntop candidates remain), val_out (optional) the |r| found.  Stand-alone: needs no handle, allocates and
 * frees its own device memory on `device`; device_ms_out (optional) receives the CUDA-event time of the kernels.
 * fp32 on the device against float64 in the reference: selections differ only at ties closer than ~1e-6. */
int di_corr_topk(int32_t device, const float* raw, int64_t n_cells, int64_t n_genes, const int32_t* cand, int64_t n_cand,
                 const int32_t* targ, int32_t n_subnets, int32_t sub_outputdim, int32_t ntop, int32_t* top_out,
                 float* val_out, float* device_ms_out);
const char* di_corr_last_error(void);

/* Introspection for tests and benchmarks. */
int di_device_sync(di_handle* h);
/* CUDA-event stopwatch on the handle's own stream (the stream every kernel of this handle is launched on):
 * di_timer_start records an event, di_timer_stop records a second one, waits for it and returns the elapsed
 * device time in ms -- idle gaps between calls included, host work after the last kernel excluded. */
int di_timer_start(di_handle* h);
int di_timer_stop(di_handle* h, float* ms_out);
/* kernel launches issued by this handle since creation (bench.py reports them as gpu_launches) */
int64_t di_launch_count(const di_handle* h);
/* CUDA-event time in ms of the device work of the last di_train_epoch / di_predict* / di_train_step call */
float di_last_device_ms(const di_handle* h);
/* Average CUDA-event duration (ms) of the kernel named `which` over the last di_train_epoch, measured on the
 * handle's stream when profiling is enabled with di_set_profiling(h, 1); -1 if unknown.
 * names: "gather", "log1p", "impute", "fwd1", "fwd2", "bwd", "adam" (tensor-core modes; "adam2", "adam1", "bias" in fp32 mode), "infer1", "infer2".
 * di_kernel_launches: how many launches of that kernel the average covers. */
int di_set_profiling(di_handle* h, int32_t on);
float di_kernel_ms(const di_handle* h, const char* which);
int64_t di_kernel_launches(const di_handle* h, const char* which);
/* Copies an internal activation buffer of the last training step to the host (parity debugging):
 * which = "h" [B][S*Hp], "dz2" [B][S*Op], "dz1" [B][S*Hp], or "norm" [N][G] (the resident normalised matrix);
 * *ld receives the row pitch in floats. */
int di_debug_read(di_handle* h, const char* which, float* out, int64_t capacity_floats, int64_t* ld);
/* One line naming the kernels and knobs this handle runs with (kernel family, ring depths, epoch graph, L2 window),
 * valid until the next call on this thread; and how many epochs could NOT be replayed as a CUDA graph and were issued
 * step by step instead (0 in normal operation; also reported on stderr the first time it happens). */
const char* di_describe(di_handle* h);
int64_t di_graph_fallbacks(di_handle* h);
int di_version(void);
/* 1 if this build carries kernels for the DI_MATH_* mode, else 0 (di_create then fails with DI_ERR_ARG). */
int di_math_mode_available(int32_t math_mode);

#ifdef __cplusplus
}
#endif
#endif /* DEEPIMPUTE_B200_H */
