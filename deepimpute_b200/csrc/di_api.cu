// C-ABI of the DeepImpute B200 engine (include/deepimpute_b200.h): handle, device memory, staging and the
// epoch / step / inference drivers.  All arithmetic is in kernels_simt.cu (DI_MATH_FP32) and kernels_tc.cu
// (DI_MATH_TF32); there is no host arithmetic and no CPU fallback.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include <nvtx3/nvToolsExt.h>   // header-only: ranges show up under a profiler, cost nothing without one

#include "engine.h"

using namespace di;

namespace {
struct NvtxRange {   // one range per C-ABI call of the path: upload / split / epoch / predict / impute
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};
}  // namespace

struct di_handle { Engine e; };

static thread_local std::string g_create_err;

namespace di {

void count_launch(Engine& e, const char* name) {
    ++e.launches;
    // DEEPIMPUTE_B200_DEBUG_SYNC=1: wait for every kernel and name the one that failed (asynchronous errors are
    // otherwise reported by whatever runtime call comes next)
    static const bool debug_sync = [] { const char* v = getenv("DEEPIMPUTE_B200_DEBUG_SYNC"); return v && *v == '1'; }();
    if (debug_sync) {
        cudaError_t err = cudaStreamSynchronize(e.stream);
        if (err == cudaSuccess) err = cudaGetLastError();
        if (err != cudaSuccess) fprintf(stderr, "deepimpute_b200: kernel '%s' (launch %lld) failed: %s\n", name,
                                        (long long)e.launches, cudaGetErrorString(err));
    }
}

// Per-kernel timing: an event pair around every launch on e.stream, recorded without synchronising and resolved
// at the next host sync (resolve_timers), so the kernels still run back to back while they are being timed.
static cudaEvent_t pooled_event(Engine& e) {
    if (!e.event_pool.empty()) { cudaEvent_t ev = e.event_pool.back(); e.event_pool.pop_back(); return ev; }
    cudaEvent_t ev = nullptr;
    cudaEventCreate(&ev);
    return ev;
}
KernelTimer::KernelTimer(Engine& e_, const char* n) : e(e_), name(n) {
    if (!e.profiling) return;
    a = pooled_event(e); b = pooled_event(e);
    cudaEventRecord(a, e.stream);
}
KernelTimer::~KernelTimer() {
    if (!a) return;
    cudaEventRecord(b, e.stream);
    e.pending_timers.push_back(Engine::PendingTimer{name, a, b});
}
void resolve_timers(Engine& e) {
    for (auto& t : e.pending_timers) {
        float ms = 0.f;
        if (cudaEventSynchronize(t.b) == cudaSuccess && cudaEventElapsedTime(&ms, t.a, t.b) == cudaSuccess) {
            auto& acc = e.kernel_ms[t.name];
            acc.first += ms; acc.second += 1;
        }
        e.event_pool.push_back(t.a); e.event_pool.push_back(t.b);
    }
    e.pending_timers.clear();
}

}  // namespace di

namespace {

#define DI_CUDA(call)                                                                                     \
    do {                                                                                                  \
        cudaError_t err__ = (call);                                                                       \
        if (err__ != cudaSuccess) {                                                                       \
            char buf__[512];                                                                              \
            snprintf(buf__, sizeof buf__, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(err__),      \
                     __FILE__, __LINE__);                                                                 \
            e.err = buf__;                                                                                \
            return err__ == cudaErrorMemoryAllocation ? DI_ERR_OOM : DI_ERR_CUDA;                         \
        }                                                                                                 \
    } while (0)

int round_up(int x, int m) { return (x + m - 1) / m * m; }
int64_t round_up64(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

template <typename T>
int dev_alloc(Engine& e, T** p, int64_t count, bool zero = true) {
    *p = nullptr;
    if (count <= 0) return DI_OK;
    DI_CUDA(cudaMalloc((void**)p, (size_t)count * sizeof(T)));
    if (zero) DI_CUDA(cudaMemsetAsync(*p, 0, (size_t)count * sizeof(T), e.stream));
    return DI_OK;
}
template <typename T>
void dev_free(T*& p) { if (p) cudaFree(p); p = nullptr; }

int fail(Engine& e, int code, const char* msg) { e.err = msg; return code; }

AdamParams adam_for_step(const Engine& e, int64_t step) {
    // Keras/TF: lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t), t = 1 for the first step
    const double t = (double)(step + 1);
    const double b1 = (double)e.cfg.beta1;   // cfg carries float32 values
    const double b2 = (double)e.cfg.beta2;
    AdamParams a;
    a.lr_t = (float)((double)e.cfg.learning_rate * std::sqrt(1.0 - std::pow(b2, t)) / (1.0 - std::pow(b1, t)));
    a.one_minus_b1 = 1.0f - e.cfg.beta1;
    a.one_minus_b2 = 1.0f - e.cfg.beta2;
    a.eps = e.cfg.epsilon;
    return a;
}

int sync_check(Engine& e) {
    DI_CUDA(cudaStreamSynchronize(e.stream));
    DI_CUDA(cudaGetLastError());
    if (!e.pending_timers.empty()) resolve_timers(e);
    if (e.cfg.math_mode != DI_MATH_FP32) {
        const unsigned int w = tc_take_timeout_word();
        if (w) {
            char buf[160];
            snprintf(buf, sizeof buf, "tensor-core kernel gave up waiting on an mbarrier (tag %u, sub-network %u, "
                     "block y %u, warp %u): results of this call are invalid", (w >> 24) & 0x7Fu, (w >> 12) & 0xFFFu,
                     (w >> 4) & 0xFFu, w & 0xFu);
            e.err = buf;
            return DI_ERR_CUDA;
        }
    }
    return DI_OK;
}

void free_split(Engine& e) {
    dev_free(e.d_train_rows); dev_free(e.d_test_rows); dev_free(e.d_perm);
    dev_free(e.Xtr); dev_free(e.Ytr); dev_free(e.Xte); dev_free(e.Yte); dev_free(e.Xtr_lo); dev_free(e.Xte_lo);
    e.n_train = e.n_test = e.n_train_pad = e.n_test_pad = 0;
}

int run_step(Engine& e, const float* X, const float* Y, int64_t row0, int n_valid, int64_t step, int which_x) {
    StepArgs a;
    a.X = X; a.Y = Y; a.ldx = e.PT; a.ldy = (int64_t)e.S * e.Op; a.row0 = row0; a.n_valid = n_valid;
    a.step = (uint32_t)step; a.adam = adam_for_step(e, step);
    if (e.cfg.math_mode != DI_MATH_FP32) tc_train_step(e, a, which_x);
    else simt_train_step(e, a);
    e.adam_t = step + 1;
    return DI_OK;
}

// forward over rows [row0, row0+rows) of a resident packed matrix; rows is a multiple of the inference tile
void run_forward(Engine& e, int which_x, const float* X, const float* Y, int64_t row0, int64_t rows,
                 int64_t n_valid, float* out, int64_t ld_out) {
    if (e.cfg.math_mode != DI_MATH_FP32) {
        tc_forward(e, which_x, row0, rows, n_valid, Y != nullptr, out, ld_out);
    } else {
        simt_forward(e, X + row0 * e.PT, e.PT, rows, n_valid, e.Hchunk,
                     Y ? Y + row0 * (int64_t)e.S * e.Op : nullptr, (int64_t)e.S * e.Op, out, ld_out);
    }
}

int validation_pass(Engine& e) {
    DI_CUDA(cudaMemsetAsync(e.d_loss + 1, 0, sizeof(double), e.stream));
    for (int64_t r0 = 0; r0 < e.n_test_pad; r0 += e.chunk_rows) {
        const int64_t rows = std::min(e.chunk_rows, e.n_test_pad - r0);
        const int64_t valid = std::max<int64_t>(0, std::min(rows, e.n_test - r0));
        run_forward(e, 2, e.Xte, e.Yte, r0, rows, valid, nullptr, 0);
    }
    return DI_OK;
}

int read_losses(Engine& e, double out[2]) {
    DI_CUDA(cudaMemcpyAsync(out, e.d_loss, 2 * sizeof(double), cudaMemcpyDeviceToHost, e.stream));
    return sync_check(e);
}

}  // namespace

extern "C" {

int di_version(void) { return 100; }

int di_math_mode_available(int32_t math_mode) {
    if (math_mode == DI_MATH_FP32) return 1;
    if (math_mode == DI_MATH_TF32 || math_mode == DI_MATH_TF32X3) return tc_available() ? 1 : 0;
    return 0;
}

const char* di_last_error(const di_handle* h) { return h ? h->e.err.c_str() : g_create_err.c_str(); }

int di_create(di_handle** out, const di_config* cfg, const int32_t* n_pred) {
    if (!out || !cfg || !n_pred) { g_create_err = "null argument"; return DI_ERR_ARG; }
    *out = nullptr;
    if (cfg->n_subnets <= 0 || cfg->hidden <= 0 || cfg->sub_outputdim <= 0 || cfg->batch_size <= 0 ||
        cfg->dropout_rate < 0.f || cfg->dropout_rate >= 1.f ||
        (cfg->math_mode != DI_MATH_FP32 && cfg->math_mode != DI_MATH_TF32 && cfg->math_mode != DI_MATH_TF32X3)) {
        g_create_err = "invalid di_config"; return DI_ERR_ARG;
    }
    for (int s = 0; s < cfg->n_subnets; ++s)
        if (n_pred[s] <= 0) { g_create_err = "n_pred must be positive"; return DI_ERR_ARG; }
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0) {
        g_create_err = std::string("no CUDA device (there is no CPU fallback): ") + cudaGetErrorString(ce);
        return DI_ERR_CUDA;
    }
    if (cfg->device < 0 || cfg->device >= ndev) { g_create_err = "device ordinal out of range"; return DI_ERR_ARG; }
    di_handle* h = new (std::nothrow) di_handle();
    if (!h) { g_create_err = "host allocation failed"; return DI_ERR_OOM; }
    Engine& e = h->e;
    e.cfg = *cfg;
    e.S = cfg->n_subnets; e.H = cfg->hidden; e.O = cfg->sub_outputdim; e.B = cfg->batch_size;
    e.Bp = round_up(e.B, 32);          // rows of every per-step buffer; batch i is staged at rows [i*Bp, i*Bp + B)
    e.Hp = round_up(e.H, 32); e.Op = round_up(e.O, 32);
    e.P.assign(n_pred, n_pred + e.S);
    e.Pp.resize(e.S); e.coff.resize(e.S);
    int64_t off = 0;
    for (int s = 0; s < e.S; ++s) {
        e.Pp[s] = round_up(e.P[s], 32);
        e.coff[s] = off; off += e.Pp[s];
        e.maxPp = std::max(e.maxPp, e.Pp[s]);
    }
    e.PT = off;

    auto body = [&]() -> int {
        DI_CUDA(cudaSetDevice(cfg->device));
        DI_CUDA(cudaStreamCreateWithFlags(&e.stream, cudaStreamNonBlocking));
        DI_CUDA(cudaStreamCreateWithFlags(&e.copy_stream, cudaStreamNonBlocking));
        DI_CUDA(cudaEventCreate(&e.ev0)); DI_CUDA(cudaEventCreate(&e.ev1));
        DI_CUDA(cudaEventCreate(&e.ev_t0)); DI_CUDA(cudaEventCreate(&e.ev_t1));
        std::vector<SubnetDesc> desc(e.S);
        for (int s = 0; s < e.S; ++s) desc[s] = SubnetDesc{e.coff[s], e.P[s], e.Pp[s], s, 0};
        e.gid.resize(e.S);
        for (int s = 0; s < e.S; ++s) e.gid[s] = s;
        int rc;
        if ((rc = dev_alloc(e, &e.d_desc, e.S))) return rc;
        DI_CUDA(cudaMemcpyAsync(e.d_desc, desc.data(), sizeof(SubnetDesc) * e.S, cudaMemcpyHostToDevice, e.stream));
        const int64_t n1 = e.PT * e.Hp, nb1 = (int64_t)e.S * e.Hp, n2 = (int64_t)e.S * e.Hp * e.Op, nb2 = (int64_t)e.S * e.Op;
        {   // one slab: [W1 mW1 vW1 W2 mW2 vW2 | W1lo W2lo | biases and their moments]; every block a multiple of 128 B
            const bool x3 = cfg->math_mode == DI_MATH_TF32X3;
            const int64_t total = 3 * n1 + 3 * n2 + (x3 ? n1 + n2 : 0) + 3 * nb1 + 3 * nb2;
            if ((rc = dev_alloc(e, &e.state_slab, total))) return rc;
            e.state_bytes = (size_t)total * sizeof(float);
            float* q = e.state_slab;
            auto take = [&](float** dst, int64_t n) { *dst = q; q += n; };
            take(&e.W1, n1); take(&e.mW1, n1); take(&e.vW1, n1);
            take(&e.W2, n2); take(&e.mW2, n2); take(&e.vW2, n2);
            if (x3) { take(&e.W1lo, n1); take(&e.W2lo, n2); }
            take(&e.b1, nb1); take(&e.mb1, nb1); take(&e.vb1, nb1);
            take(&e.b2, nb2); take(&e.mb2, nb2); take(&e.vb2, nb2);
        }
        if ((rc = dev_alloc(e, &e.d_pred_cols, e.PT))) return rc;
        if ((rc = dev_alloc(e, &e.d_targ_cols, nb2))) return rc;
        if ((rc = dev_alloc(e, &e.Xstep, (int64_t)e.Bp * e.PT))) return rc;
        if ((rc = dev_alloc(e, &e.Ystep, (int64_t)e.Bp * nb2))) return rc;
        if ((rc = dev_alloc(e, &e.d_step_rows, e.Bp))) return rc;
        if ((rc = dev_alloc(e, &e.Hact, (int64_t)e.Bp * nb1))) return rc;
        if ((rc = dev_alloc(e, &e.DZ2, (int64_t)e.Bp * nb2))) return rc;
        if ((rc = dev_alloc(e, &e.DZ1, (int64_t)e.Bp * nb1))) return rc;
        if (cfg->math_mode == DI_MATH_TF32X3) {
            if ((rc = dev_alloc(e, &e.Hlo, (int64_t)e.Bp * nb1))) return rc;
            if ((rc = dev_alloc(e, &e.DZ2lo, (int64_t)e.Bp * nb2))) return rc;
            if ((rc = dev_alloc(e, &e.DZ1lo, (int64_t)e.Bp * nb1))) return rc;
            if ((rc = dev_alloc(e, &e.Xstep_lo, (int64_t)e.Bp * e.PT))) return rc;
        }
        if ((rc = dev_alloc(e, &e.d_loss, 2))) return rc;
        // inference chunk: keep Xchunk + Hchunk + Ochunk around 1.5 GB
        const int64_t per_row = (e.PT + nb1 + 2 * nb2) * (int64_t)sizeof(float);
        int64_t cr = (int64_t)(1536ll << 20) / std::max<int64_t>(per_row, 1);
        // inference tile: 256 cells per CTA on the tensor-core path (N = 256 MMAs read 12 KB of operands per 128 cycles
        // instead of 8 KB per 64: measured 5.09 -> 4.50 ms for 16k cells x 40 sub-networks), 128 otherwise
        // (tf32x3: the persistent inference kernel works on 128-cell tiles with a double-buffered accumulator)
        e.infer_tile = cfg->math_mode == DI_MATH_TF32 ? 256 : 128;
        if (const char* v = getenv("DEEPIMPUTE_B200_INFER_TILE")) if (atoi(v) == 128) e.infer_tile = 128;
        cr = std::max<int64_t>(e.infer_tile, std::min<int64_t>(cr / e.infer_tile * e.infer_tile, 16384));
        e.chunk_rows = cr;
        if ((rc = dev_alloc(e, &e.Xchunk, cr * e.PT))) return rc;
        if ((rc = dev_alloc(e, &e.Hchunk, cr * nb1))) return rc;
        if ((rc = dev_alloc(e, &e.Ochunk, cr * nb2))) return rc;
        if ((rc = dev_alloc(e, &e.OchunkB, cr * nb2))) return rc;
        if (cfg->math_mode == DI_MATH_TF32X3) {
            if ((rc = dev_alloc(e, &e.Xchunk_lo, cr * e.PT))) return rc;
            if ((rc = dev_alloc(e, &e.Hchunk_lo, cr * nb1))) return rc;
        }
        e.Ochunk2[0] = e.Ochunk; e.Ochunk2[1] = e.OchunkB;
        if ((rc = dev_alloc(e, &e.d_chunk_rows, cr))) return rc;
        for (int i = 0; i < 2; ++i) {
            DI_CUDA(cudaMallocHost((void**)&e.h_pinned[i], (size_t)cr * e.S * e.O * sizeof(float)));
            DI_CUDA(cudaEventCreateWithFlags(&e.ev_pinned[i], cudaEventDisableTiming));
            DI_CUDA(cudaEventCreateWithFlags(&e.ev_fwd[i], cudaEventDisableTiming));
        }
        if (cfg->math_mode != DI_MATH_FP32 && !tc_init(e)) return DI_ERR_CUDA;
        return sync_check(e);
    };
    int rc = body();
    if (rc != DI_OK) { g_create_err = e.err; di_destroy(h); return rc; }
    *out = h;
    return DI_OK;
}

void di_destroy(di_handle* h) {
    if (!h) return;
    Engine& e = h->e;
    cudaSetDevice(e.cfg.device);
    if (e.stream) cudaStreamSynchronize(e.stream);
    tc_destroy(e);
    free_split(e);
    dev_free(e.d_desc); dev_free(e.d_norm); dev_free(e.d_pred_cols); dev_free(e.d_targ_cols);
    float** all[] = {&e.state_slab, &e.Xstep, &e.Ystep, &e.Hact, &e.DZ2, &e.DZ1, &e.Xchunk, &e.Hchunk, &e.Ochunk, &e.OchunkB,
                     &e.Hlo, &e.DZ2lo, &e.DZ1lo, &e.Xstep_lo, &e.Xchunk_lo, &e.Hchunk_lo};
    for (float** p : all) dev_free(*p);
    dev_free(e.d_step_rows); dev_free(e.d_chunk_rows); dev_free(e.d_loss);
    dev_free(e.d_raw_max); dev_free(e.d_gene_off); dev_free(e.d_gene_slots);
    if (e.d_raw) cudaFree(e.d_raw);
    for (int i = 0; i < 2; ++i) if (e.d_imp[i]) cudaFree(e.d_imp[i]);
    resolve_timers(e);
    for (cudaEvent_t ev : e.event_pool) cudaEventDestroy(ev);
    for (int i = 0; i < 2; ++i) {
        if (e.h_pinned[i]) cudaFreeHost(e.h_pinned[i]);
        if (e.ev_pinned[i]) cudaEventDestroy(e.ev_pinned[i]);
        if (e.ev_fwd[i]) cudaEventDestroy(e.ev_fwd[i]);
    }
    if (e.ev0) cudaEventDestroy(e.ev0);
    if (e.ev1) cudaEventDestroy(e.ev1);
    if (e.ev_t0) cudaEventDestroy(e.ev_t0);
    if (e.ev_t1) cudaEventDestroy(e.ev_t1);
    if (e.stream) cudaStreamDestroy(e.stream);
    if (e.copy_stream) cudaStreamDestroy(e.copy_stream);
    delete h;
}

int di_set_subnet_ids(di_handle* h, const int32_t* ids) {
    if (!h) return DI_ERR_ARG;
    Engine& e = h->e;
    if (!ids) return fail(e, DI_ERR_ARG, "di_set_subnet_ids: null argument");
    for (int s = 0; s < e.S; ++s) if (ids[s] < 0) return fail(e, DI_ERR_ARG, "di_set_subnet_ids: negative id");
    DI_CUDA(cudaSetDevice(e.cfg.device));
    std::vector<SubnetDesc> desc(e.S);
    for (int s = 0; s < e.S; ++s) { e.gid[s] = ids[s]; desc[s] = SubnetDesc{e.coff[s], e.P[s], e.Pp[s], ids[s], 0}; }
    DI_CUDA(cudaMemcpyAsync(e.d_desc, desc.data(), sizeof(SubnetDesc) * e.S, cudaMemcpyHostToDevice, e.stream));
    return sync_check(e);
}

// (re)allocates the resident normalised matrix for an n_cells x n_genes input; everything staged from the previous
// matrix is dropped when the shape changes
static int ensure_matrix(Engine& e, int64_t n_cells, int64_t n_genes) {
    if (n_cells != e.N || n_genes != e.G) {
        DI_CUDA(cudaStreamSynchronize(e.stream));
        dev_free(e.d_norm);
        free_split(e);
        e.N = e.G = 0;
        int rc = dev_alloc(e, &e.d_norm, n_cells * n_genes, false);
        if (rc) return rc;
        e.N = n_cells; e.G = n_genes;
    }
    return DI_OK;
}

static void drop_counts(Engine& e) {
    if (e.d_raw) cudaFree(e.d_raw);
    e.d_raw = nullptr; e.raw_dtype = -1; e.raw_max = 0.0;
}

int di_upload_matrix(di_handle* h, const float* norm, int64_t n_cells, int64_t n_genes) {
    if (!h) return DI_ERR_ARG;
    Engine& e = h->e;
    NvtxRange nvtx_range("di_upload_matrix");
    if (!norm || n_cells <= 0 || n_genes <= 0) return fail(e, DI_ERR_ARG, "di_upload_matrix: bad arguments");
    DI_CUDA(cudaSetDevice(e.cfg.device));
    int rc = ensure_matrix(e, n_cells, n_genes);
    if (rc) return rc;
    DI_CUDA(cudaStreamSynchronize(e.stream));
    drop_counts(e);                 // the counts of a previous di_upload_counts no longer describe this matrix
    DI_CUDA(cudaMemcpyAsync(e.d_norm, norm, (size_t)n_cells * n_genes * sizeof(float), cudaMemcpyHostToDevice, e.stream));
    e.split_stale = true;           // staged train / test matrices (if any) hold values of the previous matrix
    return sync_check(e);
}

int di_upload_counts(di_handle* h, const void* raw, int32_t dtype, int64_t n_cells, int64_t n_genes) {
    if (!h) return DI_ERR_ARG;
    Engine& e = h->e;
    NvtxRange nvtx_range("di_upload_counts");
    if (!raw || n_cells <= 0 || n_genes <= 0 || (dtype != DI_DTYPE_F32 && dtype != DI_DTYPE_F64))
        return fail(e, DI_ERR_ARG, "di_upload_counts: bad arguments");
    DI_CUDA(cudaSetDevice(e.cfg.device));
    const bool same = e.d_raw && dtype == e.raw_dtype && n_cells == e.N && n_genes == e.G;
    int rc = ensure_matrix(e, n_cells, n_genes);
    if (rc) return rc;
    const size_t bytes = (size_t)n_cells * n_genes * (dtype == DI_DTYPE_F64 ? sizeof(double) : sizeof(float));
    if (!same) {
        DI_CUDA(cudaStreamSynchronize(e.stream));
        drop_counts(e);
        DI_CUDA(cudaMalloc(&e.d_raw, bytes));
        e.raw_dtype = dtype;
    }
    if (!e.d_raw_max && (rc = dev_alloc(e, &e.d_raw_max, 1))) return rc;
    DI_CUDA(cudaMemcpyAsync(e.d_raw, raw, bytes, cudaMemcpyHostToDevice, e.stream));
    DI_CUDA(cudaMemsetAsync(e.d_raw_max, 0, sizeof(unsigned long long), e.stream));
    DI_CUDA(cudaEventRecord(e.ev0, e.stream));
    launch_counts_to_norm(e, e.d_raw, dtype, e.d_norm, n_cells * n_genes, e.d_raw_max);
    DI_CUDA(cudaEventRecord(e.ev1, e.stream));
    unsigned long long bits = 0;
    DI_CUDA(cudaMemcpyAsync(&bits, e.d_raw_max, sizeof bits, cudaMemcpyDeviceToHost, e.stream));
    e.split_stale = true;
    rc = sync_check(e);
    if (rc) return rc;
    memcpy(&e.raw_max, &bits, sizeof bits);
    DI_CUDA(cudaEventElapsedTime(&e.last_ms, e.ev0, e.ev1));
    return DI_OK;
}

int di_set_partition(di_handle* h, const int32_t* pred_idx, const int64_t* pred_off, const int32_t* targ_idx) {
    if (!h) return DI_ERR_ARG;
    Engine& e = h->e;
    if (!pred_idx || !pred_off || !targ_idx) return fail(e, DI_ERR_ARG, "di_set_partition: null argument");
    if (!e.d_norm) return fail(e, DI_ERR_ARG, "di_set_partition: call di_upload_matrix first");
    DI_CUDA(cudaSetDevice(e.cfg.device));
    std::vector<int32_t> pc((size_t)e.PT, -1), tc((size_t)e.S * e.Op, -1);
    for (int s = 0; s < e.S; ++s) {
        if (pred_off[s + 1] - pred_off[s] != e.P[s]) return fail(e, DI_ERR_ARG, "di_set_partition: pred_off does not match n_pred");
        for (int j = 0; j < e.P[s]; ++j) {
            const int32_t c = pred_idx[pred_off[s] + j];
            if (c < 0 || c >= e.G) return fail(e, DI_ERR_ARG, "di_set_partition: predictor column out of range");
            pc[(size_t)e.coff[s] + j] = c;
        }
        for (int o = 0; o < e.O; ++o) {
            const int32_t c = targ_idx[(size_t)s * e.O + o];
            if (c < 0 || c >= e.G) return fail(e, DI_ERR_ARG, "di_set_partition: target column out of range");
            tc[(size_t)s * e.Op + o] = c;
        }
    }
    DI_CUDA(cudaMemcpyAsync(e.d_pred_cols, pc.data(), pc.size() * sizeof(int32_t), cudaMemcpyHostToDevice, e.stream));
    DI_CUDA(cudaMemcpyAsync(e.d_targ_cols, tc.data(), tc.size() * sizeof(int32_t), cudaMemcpyHostToDevice, e.stream));
    int rc = sync_check(e);
    e.have_partition = rc == DI_OK;
    e.h_targ.assign(targ_idx, targ_idx + (size_t)e.S * e.O);
    // a new partition invalidates the CONTENTS of the staged train/test matrices; the buffers (and everything bound
    // to their addresses: tensor maps, the epoch graph) are kept and refilled by the next di_set_split
    e.split_stale = true;
    return rc;
}

int di_set_split(di_handle* h, const int32_t* train_rows, int64_t n_train, const int32_t* test_rows, int64_t n_test) {
    if (!h) return DI_ERR_ARG;
    Engine& e = h->e;
    NvtxRange nvtx_range("di_set_split");
    if (!e.have_partition) return fail(e, DI_ERR_ARG, "di_set_split: call di_set_partition first");
    if (n_train < 0 || n_test < 0 || (n_train && !train_rows) || (n_test && !test_rows))
        return fail(e, DI_ERR_ARG, "di_set_split: bad arguments");
    for (int64_t i = 0; i < n_train; ++i) if (train_rows[i] < 0 || train_rows[i] >= e.N) return fail(e, DI_ERR_ARG, "di_set_split: train row out of range");
    for (int64_t i = 0; i < n_test; ++i) if (test_rows[i] < 0 || test_rows[i] >= e.N) return fail(e, DI_ERR_ARG, "di_set_split: test row out of range");
    DI_CUDA(cudaSetDevice(e.cfg.device));
    DI_CUDA(cudaStreamSynchronize(e.stream));
    const int64_t ldy = (int64_t)e.S * e.Op;
    // same geometry as the staged split: keep the buffers, so tensor maps and the captured epoch graph stay valid
    const bool reuse = e.Xtr && e.Xte && n_train == e.n_train && n_test == e.n_test;
    int rc;
    if (!reuse) {
        free_split(e);
        e.n_train = n_train; e.n_test = n_test;
        e.n_train_pad = std::max<int64_t>((n_train + e.B - 1) / e.B, 1) * e.Bp;
        e.n_test_pad = round_up64(std::max<int64_t>(n_test, 1), e.infer_tile);
        if ((rc = dev_alloc(e, &e.d_train_rows, std::max<int64_t>(n_train, 1)))) return rc;
        if ((rc = dev_alloc(e, &e.d_perm, std::max<int64_t>(n_train, 1)))) return rc;
        if ((rc = dev_alloc(e, &e.d_test_rows, std::max<int64_t>(n_test, 1)))) return rc;
        if ((rc = dev_alloc(e, &e.Xtr, e.n_train_pad * e.PT, false))) return rc;
        if (e.cfg.math_mode == DI_MATH_TF32X3 && (rc = dev_alloc(e, &e.Xtr_lo, e.n_train_pad * e.PT, false))) return rc;
        if ((rc = dev_alloc(e, &e.Ytr, e.n_train_pad * ldy, false))) return rc;
        if ((rc = dev_alloc(e, &e.Xte, e.n_test_pad * e.PT, false))) return rc;
        if (e.cfg.math_mode == DI_MATH_TF32X3 && (rc = dev_alloc(e, &e.Xte_lo, e.n_test_pad * e.PT, false))) return rc;
        if ((rc = dev_alloc(e, &e.Yte, e.n_test_pad * ldy, false))) return rc;
    }
    e.split_stale = false;
    if (n_train) DI_CUDA(cudaMemcpyAsync(e.d_train_rows, train_rows, n_train * sizeof(int32_t), cudaMemcpyHostToDevice, e.stream));
    if (n_test) DI_CUDA(cudaMemcpyAsync(e.d_test_rows, test_rows, n_test * sizeof(int32_t), cudaMemcpyHostToDevice, e.stream));
    // held-out matrices are staged once; training matrices are re-gathered in shuffled order every epoch
    launch_gather_xy(e, e.d_test_rows, nullptr, e.n_test_pad, n_test, 0, 0, e.Xte, e.Xte_lo, e.Yte);
    if (!reuse && e.cfg.math_mode != DI_MATH_FP32 && !tc_rebind(e)) return DI_ERR_CUDA;
    return sync_check(e);
}

static int weights_xfer(di_handle* h, int32_t s, float* W1, float* b1, float* W2, float* b2, int which, bool to_device) {
    // which: 0 = weights, 1 = first moments, 2 = second moments
    Engine& e = h->e;
    if (s < 0 || s >= e.S) return fail(e, DI_ERR_ARG, "sub-network index out of range");
    DI_CUDA(cudaSetDevice(e.cfg.device));
    float* dW1 = (which == 0 ? e.W1 : which == 1 ? e.mW1 : e.vW1) + e.coff[s] * e.Hp;
    float* db1 = (which == 0 ? e.b1 : which == 1 ? e.mb1 : e.vb1) + (int64_t)s * e.Hp;
    float* dW2 = (which == 0 ? e.W2 : which == 1 ? e.mW2 : e.vW2) + (int64_t)s * e.Hp * e.Op;
    float* db2 = (which == 0 ? e.b2 : which == 1 ? e.mb2 : e.vb2) + (int64_t)s * e.Op;
    const cudaMemcpyKind kind = to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
    auto copy2d = [&](float* dev, int64_t dpitch, float* host, int64_t rows, int64_t cols) -> cudaError_t {
        if (!host) return cudaSuccess;
        return to_device
            ? cudaMemcpy2DAsync(dev, dpitch * sizeof(float), host, cols * sizeof(float), cols * sizeof(float), rows, kind, e.stream)
            : cudaMemcpy2DAsync(host, cols * sizeof(float), dev, dpitch * sizeof(float), cols * sizeof(float), rows, kind, e.stream);
    };
    DI_CUDA(copy2d(dW1, e.Hp, W1, e.P[s], e.H));
    DI_CUDA(copy2d(db1, e.Hp, b1, 1, e.H));
    DI_CUDA(copy2d(dW2, e.Op, W2, e.H, e.O));
    DI_CUDA(copy2d(db2, e.Op, b2, 1, e.O));
    return sync_check(e);
}

int di_set_weights(di_handle* h, int32_t s, const float* W1, const float* b1, const float* W2, const float* b2) {
    if (!h) return DI_ERR_ARG;
    Engine& e = h->e;
    if (!W1 || !b1 || !W2 || !b2) return fail(e, DI_ERR_ARG, "di_set_weights: null argument");
    if (s < 0 || s >= e.S) return fail(e, DI_ERR_ARG, "sub-network index out of range");
    DI_CUDA(cudaSetDevice(e.cfg.device));
    // zero the padded blocks (weights and both moments) before copying the real entries in
    float* w1[] = {e.W1, e.mW1, e.vW1}; float* bb1[] = {e.b1, e.mb1, e.vb1};
    float* w2[] = {e.W2, e.mW2, e.vW2}; float* bb2[] = {e.b2, e.mb2, e.vb2};
    for (int i = 0; i < 3; ++i) {
        DI_CUDA(cudaMemsetAsync(w1[i] + e.coff[s] * e.Hp, 0, (size_t)e.Pp[s] * e.Hp * sizeof(float), e.stream));
        DI_CUDA(cudaMemsetAsync(bb1[i] + (int64_t)s * e.Hp, 0, (size_t)e.Hp * sizeof(float), e.stream));
        DI_CUDA(cudaMemsetAsync(w2[i] + (int64_t)s * e.Hp * e.Op, 0, (size_t)e.Hp * e.Op * sizeof(float), e.stream));
        DI_CUDA(cudaMemsetAsync(bb2[i] + (int64_t)s * e.Op, 0, (size_t)e.Op * sizeof(float), e.stream));
    }
    e.adam_t = 0;
    int rc = weights_xfer(h, s, const_cast<float*>(W1), const_cast<float*>(b1), const_cast<float*>(W2),
                          const_cast<float*>(b2), 0, true);
    if (rc) return rc;
    if (e.cfg.math_mode != DI_MATH_FP32) { tc_weights_changed(e, s); rc = sync_check(e); }
    return rc;
}

int di_get_weights(di_handle* h, int32_t s, float* W1, float* b1, float* W2, float* b2) {
    if (!h) return DI_ERR_ARG;
    return weights_xfer(h, s, W1, b1, W2, b2, 0, false);
}

int di_get_adam_state(di_handle* h, int32_t s, float* mW1, float* vW1, float* mb1, float* vb1,
                      float* mW2, float* vW2, float* mb2, float* vb2, int64_t* t) {
    if (!h) return DI_ERR_ARG;
    int rc = weights_xfer(h, s, mW1, mb1, mW2, mb2, 1, false);
    if (rc) return rc;
    rc = weights_xfer(h, s, vW1, vb1, vW2, vb2, 2, false);
    if (t) *t = h->e.adam_t;
    return rc;
}

int di_train_step(di_handle* h, const int32_t* rows, int32_t nrows, int64_t step, float* loss_out) {
    if (!h) return DI_ERR_ARG;
    Engine& e = h->e;
    NvtxRange nvtx_range("di_train_step");
    if (!e.have_partition) return fail(e, DI_ERR_ARG, "di_train_step: no data (di_upload_matrix + di_set_partition)");
    if (!rows || nrows <= 0 || nrows > e.B || step < 0) return fail(e, DI_ERR_ARG, "di_train_step: need 1..B rows");
    for (int i = 0; i < nrows; ++i) if (rows[i] < 0 || rows[i] >= e.N) return fail(e, DI_ERR_ARG, "di_train_step: row out of range");
    DI_CUDA(cudaSetDevice(e.cfg.device));
    DI_CUDA(cudaMemcpyAsync(e.d_step_rows, rows, nrows * sizeof(int32_t), cudaMemcpyHostToDevice, e.stream));
    DI_CUDA(cudaEventRecord(e.ev0, e.stream));
    DI_CUDA(cudaMemsetAsync(e.d_loss, 0, sizeof(double), e.stream));
    launch_gather_xy(e, e.d_step_rows, nullptr, e.Bp, nrows, 0, 0, e.Xstep, e.Xstep_lo, e.Ystep);
    int rc = run_step(e, e.Xstep, e.Ystep, 0, nrows, step, 1);
    if (rc) return rc;
    DI_CUDA(cudaEventRecord(e.ev1, e.stream));
    double l[2];
    if ((rc = read_losses(e, l))) return rc;
    DI_CUDA(cudaEventElapsedTime(&e.last_ms, e.ev0, e.ev1));
    if (loss_out) *loss_out = (float)(l[0] / ((double)nrows * e.O));
    return std::isfinite(l[0]) ? DI_OK : fail(e, DI_ERR_NUMERIC, "non-finite training loss");
}

int di_validation_loss(di_handle* h, float* val_loss_out) {
    if (!h) return DI_ERR_ARG;
    Engine& e = h->e;
    if (!e.Xte || e.split_stale) return fail(e, DI_ERR_ARG, "di_validation_loss: call di_set_split first");
    DI_CUDA(cudaSetDevice(e.cfg.device));
    int rc = validation_pass(e);
    if (rc) return rc;
    double l[2];
    if ((rc = read_losses(e, l))) return rc;
    if (val_loss_out) *val_loss_out = e.n_test ? (float)(l[1] / ((double)e.n_test * e.O)) : 0.f;
    return DI_OK;
}

int di_train_epoch(di_handle* h, const int32_t* perm, int64_t first_step, float* loss_out, float* val_loss_out) {
    if (!h) return DI_ERR_ARG;
    Engine& e = h->e;
    NvtxRange nvtx_range("di_train_epoch");
    if (!e.Xtr || e.n_train <= 0 || e.split_stale) return fail(e, DI_ERR_ARG, "di_train_epoch: call di_set_split first");
    if (!perm || first_step < 0) return fail(e, DI_ERR_ARG, "di_train_epoch: bad arguments");
    for (int64_t i = 0; i < e.n_train; ++i) if (perm[i] < 0 || perm[i] >= e.n_train) return fail(e, DI_ERR_ARG, "di_train_epoch: perm out of range");
    DI_CUDA(cudaSetDevice(e.cfg.device));
    DI_CUDA(cudaMemcpyAsync(e.d_perm, perm, e.n_train * sizeof(int32_t), cudaMemcpyHostToDevice, e.stream));
    DI_CUDA(cudaEventRecord(e.ev0, e.stream));
    DI_CUDA(cudaMemsetAsync(e.d_loss, 0, 2 * sizeof(double), e.stream));
    // stage this epoch's visiting order: batch i is rows [i*B, (i+1)*B) of Xtr / Ytr
    launch_gather_xy(e, e.d_train_rows, e.d_perm, e.n_train_pad, e.n_train, e.B, e.Bp, e.Xtr, e.Xtr_lo, e.Ytr);
    const int64_t n_steps = (e.n_train + e.B - 1) / e.B;
    bool graphed = false;
    if (e.cfg.math_mode != DI_MATH_FP32 && !e.profiling) {
        // one graph launch for the whole epoch (kernels_tc.cu); the per-step Adam rates go along as a table
        std::vector<float> lr_t((size_t)n_steps);
        for (int64_t i = 0; i < n_steps; ++i) lr_t[(size_t)i] = adam_for_step(e, first_step + i).lr_t;
        graphed = tc_train_epoch_graph(e, first_step, lr_t.data(), n_steps);
        if (graphed) {
            DI_CUDA(cudaStreamSynchronize(e.stream));       // lr_t is a stack-lifetime host buffer
            e.adam_t = first_step + n_steps;
        }
    }
    int64_t step = first_step;
    for (int64_t i0 = 0, r0 = 0; !graphed && i0 < e.n_train; i0 += e.B, r0 += e.Bp, ++step) {
        const int n_valid = (int)std::min<int64_t>(e.B, e.n_train - i0);
        int rc = run_step(e, e.Xtr, e.Ytr, r0, n_valid, step, 0);
        if (rc) return rc;
    }
    int rc = validation_pass(e);
    if (rc) return rc;
    DI_CUDA(cudaEventRecord(e.ev1, e.stream));
    double l[2];
    if ((rc = read_losses(e, l))) return rc;
    DI_CUDA(cudaEventElapsedTime(&e.last_ms, e.ev0, e.ev1));
    // Keras `loss`: sum_batches(L_batch * n_batch) / n_train with L_batch = raw / (n_batch * O)
    if (loss_out) *loss_out = (float)(l[0] / ((double)e.n_train * e.O));
    if (val_loss_out) *val_loss_out = e.n_test ? (float)(l[1] / ((double)e.n_test * e.O)) : 0.f;
    return std::isfinite(l[0]) ? DI_OK : fail(e, DI_ERR_NUMERIC, "non-finite training loss");
}

static int predict_impl(di_handle* h, const int32_t* rows, int64_t n, float* host_out, float* d_out, int64_t ld_out) {
    Engine& e = h->e;
    NvtxRange nvtx_range("di_predict");
    if (!e.have_partition) return fail(e, DI_ERR_ARG, "di_predict: no data (di_upload_matrix + di_set_partition)");
    if (n < 0 || (!rows && n > e.N)) return fail(e, DI_ERR_ARG, "di_predict: bad row count");
    if (n == 0) return DI_OK;
    if (rows) for (int64_t i = 0; i < n; ++i) if (rows[i] < 0 || rows[i] >= e.N) return fail(e, DI_ERR_ARG, "di_predict: row out of range");
    DI_CUDA(cudaSetDevice(e.cfg.device));
    const int64_t SO = (int64_t)e.S * e.O;
    // a page-locked caller buffer takes the DMA directly; a pageable one is fed through two pinned staging buffers
    bool direct = false;
    if (host_out) {
        cudaPointerAttributes attr{};
        if (cudaPointerGetAttributes(&attr, host_out) == cudaSuccess && attr.type == cudaMemoryTypeHost) direct = true;
        cudaGetLastError();
    }
    DI_CUDA(cudaEventRecord(e.ev0, e.stream));
    int buf = 0;
    for (int64_t r0 = 0; r0 < n; r0 += e.chunk_rows, buf ^= 1) {
        const int64_t valid = std::min(e.chunk_rows, n - r0);
        const int64_t rows_pad = round_up64(valid, e.infer_tile);
        if (rows) DI_CUDA(cudaMemcpyAsync(e.d_chunk_rows, rows + r0, valid * sizeof(int32_t), cudaMemcpyHostToDevice, e.stream));
        launch_gather(e, rows ? e.d_chunk_rows : nullptr, nullptr, r0, rows_pad, valid, e.d_pred_cols, e.PT, e.Xchunk, 0, 0, e.Xchunk_lo);
        if (d_out) {
            run_forward(e, 3, e.Xchunk, nullptr, 0, rows_pad, valid, d_out + r0 * ld_out, ld_out);
        } else if (direct) {
            float* o = e.Ochunk2[buf];
            // the copy of chunk i-2 out of this buffer must have finished before chunk i overwrites it
            DI_CUDA(cudaStreamWaitEvent(e.stream, e.ev_pinned[buf], 0));
            run_forward(e, 3, e.Xchunk, nullptr, 0, rows_pad, valid, o, SO);
            DI_CUDA(cudaEventRecord(e.ev_fwd[buf], e.stream));
            DI_CUDA(cudaStreamWaitEvent(e.copy_stream, e.ev_fwd[buf], 0));
            DI_CUDA(cudaMemcpyAsync(host_out + r0 * SO, o, (size_t)valid * SO * sizeof(float), cudaMemcpyDeviceToHost, e.copy_stream));
            DI_CUDA(cudaEventRecord(e.ev_pinned[buf], e.copy_stream));
        } else {
            run_forward(e, 3, e.Xchunk, nullptr, 0, rows_pad, valid, e.Ochunk, SO);
            // D2H through two pinned buffers so the next chunk's kernels overlap the host-side copy-out
            DI_CUDA(cudaEventSynchronize(e.ev_pinned[buf]));
            DI_CUDA(cudaMemcpyAsync(e.h_pinned[buf], e.Ochunk, (size_t)valid * SO * sizeof(float), cudaMemcpyDeviceToHost, e.stream));
            DI_CUDA(cudaEventRecord(e.ev_pinned[buf], e.stream));
            if (r0 > 0) {
                const int prev = buf ^ 1;
                const int64_t pr0 = r0 - e.chunk_rows;
                DI_CUDA(cudaEventSynchronize(e.ev_pinned[prev]));
                memcpy(host_out + pr0 * SO, e.h_pinned[prev], (size_t)e.chunk_rows * SO * sizeof(float));
            }
        }
    }
    if (direct) {   // the timed region ends when the last copy has landed
        DI_CUDA(cudaStreamWaitEvent(e.stream, e.ev_pinned[0], 0));
        DI_CUDA(cudaStreamWaitEvent(e.stream, e.ev_pinned[1], 0));
    }
    DI_CUDA(cudaEventRecord(e.ev1, e.stream));
    int rc = sync_check(e);
    if (rc) return rc;
    if (!d_out && !direct) {
        const int64_t last0 = (n - 1) / e.chunk_rows * e.chunk_rows;
        const int last_buf = (int)((last0 / e.chunk_rows) & 1);
        memcpy(host_out + last0 * SO, e.h_pinned[last_buf], (size_t)(n - last0) * SO * sizeof(float));
    }
    DI_CUDA(cudaEventElapsedTime(&e.last_ms, e.ev0, e.ev1));
    return DI_OK;
}

int di_predict(di_handle* h, const int32_t* rows, int64_t n, float* out) {
    if (!h) return DI_ERR_ARG;
    if (!out && n > 0) return fail(h->e, DI_ERR_ARG, "di_predict: null output");
    return predict_impl(h, rows, n, out, nullptr, 0);
}

int di_predict_device(di_handle* h, const int32_t* rows, int64_t n, float* d_out, int64_t ld_out) {
    if (!h) return DI_ERR_ARG;
    if ((!d_out && n > 0) || ld_out < (int64_t)h->e.S * h->e.O) return fail(h->e, DI_ERR_ARG, "di_predict_device: bad output");
    return predict_impl(h, rows, n, nullptr, d_out, ld_out);
}

int di_impute(di_handle* h, int32_t policy, const int32_t* slot_gene, int64_t n_slots, const float* d_pred,
              int64_t ld_pred, int32_t out_dtype, void* out) {
    if (!h) return DI_ERR_ARG;
    Engine& e = h->e;
    NvtxRange nvtx_range("di_impute");
    if (!e.d_raw) return fail(e, DI_ERR_ARG, "di_impute: no counts on the device (di_upload_counts)");
    if (!out || (out_dtype != DI_DTYPE_F32 && out_dtype != DI_DTYPE_F64) ||
        (policy != DI_POLICY_NONE && policy != DI_POLICY_RESTORE && policy != DI_POLICY_MAX))
        return fail(e, DI_ERR_ARG, "di_impute: bad arguments");
    const int64_t SO = (int64_t)e.S * e.O;
    if (!d_pred) {
        if (!e.have_partition) return fail(e, DI_ERR_ARG, "di_impute: no partition (di_set_partition)");
        if (slot_gene && n_slots != SO) return fail(e, DI_ERR_ARG, "di_impute: n_slots must be S*O when the handle predicts");
        ld_pred = SO;
    } else if (!slot_gene || ld_pred < n_slots) {
        return fail(e, DI_ERR_ARG, "di_impute: a prediction matrix needs its slot table and ld_pred >= n_slots");
    }
    if (!slot_gene) {
        if (!e.have_partition) return fail(e, DI_ERR_ARG, "di_impute: no partition (di_set_partition)");
        slot_gene = e.h_targ.data(); n_slots = SO;
    }
    if (n_slots < 0) return fail(e, DI_ERR_ARG, "di_impute: negative n_slots");
    DI_CUDA(cudaSetDevice(e.cfg.device));

    // gene -> prediction columns (CSR, columns ascending inside a gene); a negative slot_gene entry is ignored
    std::vector<int32_t> ent((size_t)2 * e.G, 0), slots((size_t)std::max<int64_t>(n_slots, 1));
    for (int64_t k = 0; k < n_slots; ++k) {
        if (slot_gene[k] >= e.G) return fail(e, DI_ERR_ARG, "di_impute: slot gene out of range");
        if (slot_gene[k] >= 0) ++ent[2 * (size_t)slot_gene[k] + 1];
    }
    int32_t run = 0;
    for (int64_t g = 0; g < e.G; ++g) { ent[2 * g] = run; run += ent[2 * g + 1]; ent[2 * g + 1] = 0; }
    for (int64_t k = 0; k < n_slots; ++k) {
        if (slot_gene[k] < 0) continue;
        const size_t g = (size_t)slot_gene[k];
        slots[(size_t)ent[2 * g] + ent[2 * g + 1]++] = (int32_t)k;
    }
    int rc;
    DI_CUDA(cudaStreamSynchronize(e.stream));
    dev_free(e.d_gene_off); dev_free(e.d_gene_slots);
    if ((rc = dev_alloc(e, &e.d_gene_off, 2 * e.G, false))) return rc;
    if ((rc = dev_alloc(e, &e.d_gene_slots, (int64_t)slots.size(), false))) return rc;
    DI_CUDA(cudaMemcpyAsync(e.d_gene_off, ent.data(), ent.size() * sizeof(int32_t), cudaMemcpyHostToDevice, e.stream));
    DI_CUDA(cudaMemcpyAsync(e.d_gene_slots, slots.data(), slots.size() * sizeof(int32_t), cudaMemcpyHostToDevice, e.stream));

    const size_t esz = out_dtype == DI_DTYPE_F64 ? sizeof(double) : sizeof(float);
    const int64_t rows_max = std::min<int64_t>(e.chunk_rows, round_up64(e.N, e.infer_tile));
    const size_t need = (size_t)rows_max * e.G * esz;
    if (e.imp_bytes < need) {
        for (int i = 0; i < 2; ++i) { if (e.d_imp[i]) cudaFree(e.d_imp[i]); e.d_imp[i] = nullptr; }
        e.imp_bytes = 0;
        for (int i = 0; i < 2; ++i) DI_CUDA(cudaMalloc(&e.d_imp[i], need));
        e.imp_bytes = need;
    }
    const double clamp = 2.0 * std::log1p(e.raw_max);     // 2 * norm.max() of multinet.py:291, float64 like numpy

    DI_CUDA(cudaEventRecord(e.ev0, e.stream));
    int buf = 0;
    for (int64_t r0 = 0; r0 < e.N; r0 += e.chunk_rows, buf ^= 1) {
        const int64_t valid = std::min(e.chunk_rows, e.N - r0);
        // the copy of chunk i-2 out of this buffer pair must have finished before chunk i overwrites it
        DI_CUDA(cudaStreamWaitEvent(e.stream, e.ev_pinned[buf], 0));
        const float* pred;
        if (d_pred) {
            pred = d_pred + r0 * ld_pred;
        } else {
            const int64_t rows_pad = round_up64(valid, e.infer_tile);
            launch_gather(e, nullptr, nullptr, r0, rows_pad, valid, e.d_pred_cols, e.PT, e.Xchunk, 0, 0, e.Xchunk_lo);
            run_forward(e, 3, e.Xchunk, nullptr, 0, rows_pad, valid, e.Ochunk2[buf], SO);
            pred = e.Ochunk2[buf];
        }
        launch_impute(e, pred, ld_pred, r0, valid, clamp, policy, out_dtype, e.d_imp[buf]);
        DI_CUDA(cudaEventRecord(e.ev_fwd[buf], e.stream));
        DI_CUDA(cudaStreamWaitEvent(e.copy_stream, e.ev_fwd[buf], 0));
        DI_CUDA(cudaMemcpyAsync(static_cast<char*>(out) + (size_t)r0 * e.G * esz, e.d_imp[buf], (size_t)valid * e.G * esz,
                                cudaMemcpyDefault, e.copy_stream));
        DI_CUDA(cudaEventRecord(e.ev_pinned[buf], e.copy_stream));
    }
    DI_CUDA(cudaStreamWaitEvent(e.stream, e.ev_pinned[0], 0));
    DI_CUDA(cudaStreamWaitEvent(e.stream, e.ev_pinned[1], 0));
    DI_CUDA(cudaEventRecord(e.ev1, e.stream));
    rc = sync_check(e);
    if (rc) return rc;
    DI_CUDA(cudaEventElapsedTime(&e.last_ms, e.ev0, e.ev1));
    return DI_OK;
}

int di_device_sync(di_handle* h) {
    if (!h) return DI_ERR_ARG;
    Engine& e = h->e;
    DI_CUDA(cudaSetDevice(e.cfg.device));
    return sync_check(e);
}

int di_timer_start(di_handle* h) {
    if (!h) return DI_ERR_ARG;
    Engine& e = h->e;
    DI_CUDA(cudaSetDevice(e.cfg.device));
    DI_CUDA(cudaEventRecord(e.ev_t0, e.stream));
    return DI_OK;
}

int di_timer_stop(di_handle* h, float* ms_out) {
    if (!h || !ms_out) return DI_ERR_ARG;
    Engine& e = h->e;
    DI_CUDA(cudaSetDevice(e.cfg.device));
    DI_CUDA(cudaEventRecord(e.ev_t1, e.stream));
    DI_CUDA(cudaEventSynchronize(e.ev_t1));
    DI_CUDA(cudaEventElapsedTime(ms_out, e.ev_t0, e.ev_t1));
    return DI_OK;
}

int64_t di_launch_count(const di_handle* h) { return h ? h->e.launches : 0; }
float di_last_device_ms(const di_handle* h) { return h ? h->e.last_ms : -1.f; }

int di_set_profiling(di_handle* h, int32_t on) {
    if (!h) return DI_ERR_ARG;
    resolve_timers(h->e);
    h->e.profiling = on != 0;
    h->e.kernel_ms.clear();
    return DI_OK;
}

float di_kernel_ms(const di_handle* h, const char* which) {
    if (!h || !which) return -1.f;
    auto it = h->e.kernel_ms.find(which);
    if (it == h->e.kernel_ms.end() || it->second.second == 0) return -1.f;
    return (float)(it->second.first / (double)it->second.second);
}

int64_t di_kernel_launches(const di_handle* h, const char* which) {
    if (!h || !which) return 0;
    auto it = h->e.kernel_ms.find(which);
    return it == h->e.kernel_ms.end() ? 0 : it->second.second;
}

const char* di_describe(di_handle* h) {
    if (!h) return "";
    return h->e.cfg.math_mode == DI_MATH_FP32 ? "fp32 CUDA-core kernels" : tc_describe(h->e);
}

int64_t di_graph_fallbacks(di_handle* h) { return (h && h->e.cfg.math_mode != DI_MATH_FP32) ? tc_fallbacks(h->e) : 0; }

int di_debug_read(di_handle* h, const char* which, float* out, int64_t capacity_floats, int64_t* ld) {
    if (!h || !which || !out || !ld) return DI_ERR_ARG;
    Engine& e = h->e;
    const float* src; int64_t pitch, rows = e.Bp;
    if (!strcmp(which, "h")) { src = e.Hact; pitch = (int64_t)e.S * e.Hp; }
    else if (!strcmp(which, "dz2")) { src = e.DZ2; pitch = (int64_t)e.S * e.Op; }
    else if (!strcmp(which, "dz1")) { src = e.DZ1; pitch = (int64_t)e.S * e.Hp; }
    else if (!strcmp(which, "norm")) { src = e.d_norm; pitch = e.G; rows = e.N; }   // the resident normalised matrix
    else return fail(e, DI_ERR_ARG, "di_debug_read: unknown buffer");
    if (!src || capacity_floats < pitch * rows) return fail(e, DI_ERR_ARG, "di_debug_read: buffer too small");
    DI_CUDA(cudaSetDevice(e.cfg.device));
    DI_CUDA(cudaMemcpyAsync(out, src, (size_t)pitch * rows * sizeof(float), cudaMemcpyDeviceToHost, e.stream));
    *ld = pitch;
    return sync_check(e);
}

}  // extern "C"
