// internal.h — shared internals of libpbn_cuda.so (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <functional>
#include <string>
#include <utility>
#include <vector>

#include "../../include/pbn_cuda.h"
#include "pair_kernel.cuh"

int pbn_set_error(int code, const std::string& msg);
#define set_error pbn_set_error

// ------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------
#define PBN_CUDA_TRY(expr)                                                                                  \
    do {                                                                                                    \
        cudaError_t err__ = (expr);                                                                         \
        if (err__ != cudaSuccess) {                                                                         \
            return set_error(PBN_ERR_CUDA, std::string("CUDA error ") + cudaGetErrorName(err__) + " (" +    \
                                               cudaGetErrorString(err__) + ") at " #expr);                  \
        }                                                                                                   \
    } while (0)

#define PBN_TRY(expr)                 \
    do {                              \
        int rc__ = (expr);            \
        if (rc__ != PBN_OK) return rc__; \
    } while (0)

// ------------------------------------------------------------------------------------
// objects
// ------------------------------------------------------------------------------------
// A context drives one GPU.  A MULTI-DEVICE context (pbn_ctx_create_multi) is the context of its first device plus
// `peers`, the contexts of the others: tables, fitted KDEs, cross-validation objects and UCV scorers created through it
// carry one replica per peer (`rep[i]` lives on `peers[i]`), and the sharding entry points (pbn_kde_logl, pbn_ckde_cdf,
// pbn_cv_scores, pbn_ucv_score) split their work over all devices from one host thread per device, adding the
// per-device scalars in device order.  Everything else runs on the first device as before.
struct pbn_ctx {
    std::vector<pbn_ctx*> peers;
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t own_stream = nullptr;
    int sm_count = 0;
    double* d_exp_tab = nullptr;  // T'[j] = 2^(j/K) with (j << (20 - log2 K)) taken off the high word, j = 0..K-1 (pair_kernel.cuh)
    int64_t launches = 0, h2d = 0, d2h = 0;
    int64_t last_fallback_rows = 0;     // rows of the last logl call that took the shifted second pass
    int64_t last_row_kernel_rows = 0;   // ... of which the per-row kernel had to finish (farther than 2^31 kernel units)
    // tile skipping (spatial.cu): on by default for large single-model calls; the unit counts of the last such call
    bool skipping = true;
    int64_t last_units_total = 0, last_units_done = 0;
    int64_t units_total = 0, units_done = 0;  // accumulated while `timing` is on
    // optional device timing of the pair kernel (CUDA events on the launching stream)
    bool timing = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timed;
    double pair_ms = 0;
    int64_t pair_launches = 0;
    int64_t pair_units = 0;  // train x test pairs processed by timed launches (x2 for CKDE)
};

struct pbn_table {
    std::vector<pbn_table*> rep;  // replicas on ctx->peers (multi-device context), else empty
    pbn_ctx* ctx;
    int ncols;
    int64_t nrows;
    int64_t stride;  // elements between columns
    int dtype;
    void* data;  // [ncols][stride]
};

struct pbn_kde {
    std::vector<pbn_kde*> rep;  // replicas on ctx->peers
    pbn_ctx* ctx;
    int d;
    int dtype;
    bool ckde;  // fused joint+marginal (d >= 2, variable stored last)
    int64_t n;
    void* y;  // whitened training rows AoS [n_pad][d]
    float* d_bound;  // device scalar: max |whitened training coordinate|
    double* nrm;     // f64, d <= kMaxFastD only (else null): -sum_{c<dn} y_c^2 per training row, dn = d-1 if ckde else d
    // tile skipping (spatial.cu), large training sets with d <= kMaxFastD only (else null): the same rows in Morton order
    // (y itself keeps the caller's order: CKDE::sample addresses training rows by index), their norms, and the bounding
    // box of every training tile, [n_box_tiles][2 d] floats (min, max)
    void* ys = nullptr;
    double* nrm_s = nullptr;
    float* box = nullptr;
    int n_box_tiles = 0;
    double W[PBN_MAX_DIM * PBN_MAX_DIM];  // row-major lower-triangular whitening matrix (incl. unit scale)
    double mu[PBN_MAX_DIM];
    int perm[PBN_MAX_DIM];  // internal column k = caller column perm[k]
    double lognorm_joint;
    double lognorm_marg;
};

// ---- multi-device helpers (runtime.cu) ----
static inline int pbn_num_devices(const pbn_ctx* c) { return 1 + (int)c->peers.size(); }
static inline pbn_ctx* pbn_device_ctx(pbn_ctx* c, int i) { return i == 0 ? c : c->peers[i - 1]; }
template <typename Obj>
static inline Obj* pbn_replica(Obj* o, int i) { return i == 0 ? o : o->rep[i - 1]; }
template <typename Obj>
static inline bool pbn_replicated(const pbn_ctx* c, const Obj* o) { return !c->peers.empty() && o && o->rep.size() == c->peers.size(); }
// fn(i) for i = 0 .. n-1 on one host thread per device (i = 0 on the caller's); the first failure (lowest i) is returned
// and becomes the caller's pbn_last_error()
int pbn_run_on_devices(int n, const std::function<int(int)>& fn);
// rows [begin, end) of the virtual row order of a two-segment range
pbn_rows pbn_sub_rows(const pbn_rows& r, int64_t begin, int64_t end);

// coordinates of the Morton key of a fitted model (spatial.cu): the evidence coordinates for a CKDE unless PBN_MORTON_JOINT=1
int pbn_morton_dims(const pbn_kde* k);

static inline size_t elem_size(int dtype) { return dtype == PBN_F64 ? 8 : 4; }
static inline int64_t seg_count(const pbn_rows& r) { return (r.e0 - r.b0) + (r.e1 - r.b1); }

struct DevSetter {
    int prev = -1;
    bool ok = true;
    explicit DevSetter(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) ok = false;
        if (ok && prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DevSetter() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};


// ---- host linear algebra (runtime.cu) ----
bool chol_lower(const double* H, int d, double* L);                  // column-major lower Cholesky
void tri_inverse_rowmajor(const double* L, int d, double* Winv);     // inverse of lower-tri, row-major
bool is_psd(const double* cov, int d, int dtype);                    // util/basic_eigen_ops.hpp:136-147
double unit_scale(int dtype);  // kernel exponent units per natural-log unit of -s/2

struct ColPtrs {
    const void* p[PBN_MAX_DIM];
};
struct Vec32 {
    double v[PBN_MAX_DIM];
};
struct WhitenParams {
    ColPtrs cols;     // already permuted to internal order
    double W[PBN_MAX_DIM * (PBN_MAX_DIM + 1) / 2];  // packed lower triangle, row-major
    double mu[PBN_MAX_DIM];
    int d;
    int dn;           // coordinates entering the row norm (see pbn_kde::nrm)
    int64_t b0, n0, b1, n;
};

int check_cols(const pbn_table* tbl, const int* cols, int d);
int check_rows(const pbn_table* tbl, const pbn_rows& r);
const void* col_ptr(const pbn_table* t, int c);
std::string var_list(const int* cols, int d);
int moments_impl(pbn_ctx* ctx, const pbn_table* tbl, const int* cols, int d, pbn_rows rows, double* mean_out,
                 double* cov_out);
// y = W (x - mu) for the rows of a range, AoS output in the table's dtype; `Wfull` row-major d x d lower.
int whiten_raw_launch(pbn_ctx* ctx, const pbn_table* tbl, const int* cols_internal_order, int d, pbn_rows rows,
                      const double* Wfull, const double* mu, void* out, float* bound, double* nrm = nullptr, int dn = 0);

// whitened rows of `rows` of `tbl` under a fitted KDE's whitening matrix (cols in the KDE's variable order)
int pbn_whiten_kde(pbn_ctx* ctx, const pbn_kde* k, const pbn_table* tbl, const int* cols, pbn_rows rows, void* out,
                   float* bound, double* nrm);
// one fitted KDE against one row range (runtime.cu); device and/or host outputs, each may be null
int pbn_logl_impl(pbn_ctx* ctx, const pbn_kde* k, const pbn_table* test, const int* cols, pbn_rows rows,
                  double* d_out_logl, double* d_out_slogl, double* h_out_logl, double* h_out_slogl);

// MLE<LinearGaussianCPD> from centred moments (cv.cu): Cm = sum (x_a - mean_a)(x_b - mean_b), column-major d x d,
// variable first; writes beta[p + 1], returns the variance.
double lg_fit_from_moments(int64_t rows, int p, const double* mean, const double* Cm, double* beta);

// Stream-ordered scratch of one call: everything allocated through it is returned to the pool when the call leaves,
// on the error paths too (PBN_CUDA_TRY / PBN_TRY return early).
struct Scratch {
    cudaStream_t st;
    std::vector<void*> ptrs;
    std::vector<cudaEvent_t> events;  // destroyed on exit unless released
    explicit Scratch(cudaStream_t s) : st(s) {}
    ~Scratch() {
        for (void* p : ptrs) cudaFreeAsync(p, st);
        for (cudaEvent_t e : events) cudaEventDestroy(e);
    }
    template <typename T>
    cudaError_t alloc(T** out, size_t bytes) {
        void* p = nullptr;
        cudaError_t e = cudaMallocAsync(&p, bytes ? bytes : 8, st);
        if (e == cudaSuccess) ptrs.push_back(p);
        *out = static_cast<T*>(p);
        return e;
    }
};

// The shifted second pass over rows whose unshifted sums underflowed (runtime.cu)
struct ShiftPass {
    const void* train;   // whitened training rows AoS [n_train][d]
    int64_t n_train;
    const void* test;    // whitened test rows the flagged ids index
    int64_t m_cap;       // upper bound of the flagged count (sizes the scratch)
    int d;
    bool ckde, f64;
    double lognorm_joint, lognorm_marg;
    int* flagged;        // device: ascending row ids ...
    int* n_flagged;      // ... and their count
    double* out;         // out[row id]
    int* n_row_kernel;   // device int, may be null: how many rows the per-row kernel had to finish
};
int pbn_shift_pass(pbn_ctx* ctx, Scratch& sc, const ShiftPass& A);

// ---- spatial order and tile skipping (spatial.cu) ----
int pbn_spatial_sort(pbn_ctx* ctx, int dtype, int d, int dk, const void* y, const double* nrm, int64_t n, const float* bound, void* ys,
                     double* nrm_s, int* perm);
int pbn_spatial_boxes(pbn_ctx* ctx, int dtype, int d, const void* ys, int64_t n, int tile_rows, float* box);
int pbn_skip_nearest(pbn_ctx* ctx, const float* box_test, int n_test_tiles, const float* box_train, int n_train_tiles, int d, int K,
                     int* nearest, long long* first);
int pbn_skip_count(pbn_ctx* ctx, const pbn::PairJob* d_jobA, long long upbA, int tb, int ckde, int dtype, int64_t n_train,
                   const float* box_test, int n_test_tiles, const float* box_train, int n_train_tiles, int d, const int* nearest,
                   int K, float* thr, double* sumsA, long long* count, long long* tile_first, long long* total_out);
int pbn_skip_fill(pbn_ctx* ctx, int ckde, const float* box_test, int n_test_tiles, const float* box_train, int n_train_tiles, int d,
                  const int* nearest, int K, const float* thr, const long long* tile_first, int* unit_list);
int pbn_scatter_out(pbn_ctx* ctx, const double* src, const int* perm, int64_t n, double* dst);

namespace pbn {
cudaError_t launch_pair_f64(int D, bool ckde, const PairJob* jobs, int n_jobs, long long total_units, long long upb,
                            int grid, const double* tab, cudaStream_t stream);
cudaError_t launch_pair_f32(int D, bool ckde, const PairJob* jobs, int n_jobs, long long total_units, long long upb,
                            int grid, const double* tab, cudaStream_t stream);
cudaError_t launch_pair_gskip_f64(int D, bool ckde, const PairJob* jobs, int n_jobs, long long total_units, long long upb,
                                  int grid, const double* tab, cudaStream_t stream);
cudaError_t launch_pair_gskip_f32(int D, bool ckde, const PairJob* jobs, int n_jobs, long long total_units, long long upb,
                                  int grid, const double* tab, cudaStream_t stream);
cudaError_t launch_pair_shift_f64(int D, bool ckde, const PairJob* job, const long long* dyn, int grid, const double* tab,
                                  cudaStream_t stream);
cudaError_t launch_pair_shift_f32(int D, bool ckde, const PairJob* job, const long long* dyn, int grid, const double* tab,
                                  cudaStream_t stream);
cudaError_t launch_cdf_f64(int D, const PairJob* jobs, int n_jobs, long long total_units, long long upb, int grid,
                           const double* tab, double inv_c, cudaStream_t stream);
cudaError_t launch_cdf_f32(int D, const PairJob* jobs, int n_jobs, long long total_units, long long upb, int grid,
                           const double* tab, double inv_c, cudaStream_t stream);
int pair_tile_f64(int D);
int pair_tile_f32(int D);
int pair_tb_f64();  // UCV kernel and the CDF mode of the pair kernel (rows per thread of PairCfg)
int pair_tb_f32();
// test rows per tile of pair_kernel<T, D, CKDE> (the rows per thread depend on the kernel shape, pair_rows)
int pair_tb_for_f64(int D, bool ckde);
int pair_tb_for_f32(int D, bool ckde);
int pair_tb_cdf_f64(int D);  // CDF mode of pair_kernel
int pair_ctas_per_sm_f64();  // persistent CTAs per SM the kernels are register-bounded for (the launch grid multiplier)
int pair_ctas_per_sm_f32();
cudaError_t warm_pair_f64();
cudaError_t warm_pair_f32();
cudaError_t warm_pair_gskip_f64();
cudaError_t warm_pair_gskip_f32();
cudaError_t warm_pair_shift_f64();
cudaError_t warm_pair_shift_f32();
int pair_tb_cdf_f32(int D);
}  // namespace pbn
