// cv.cu — batched cross-validated likelihood scores (pbn_cv_*).
//
// Replaces the reference's strictly serial CVLikelihood / HoldoutLikelihood inner loop
//   learning/scores/cv_likelihood.cpp:11-25, holdout_likelihood.cpp:14-23,
//   dataset/crossvalidation_adaptator.cpp:5-31 (one Arrow Take + upload per fold per call),
//   factors/continuous/CKDE.hpp:182-287, learning/parameters/mle_LinearGaussianCPD.hpp:11-221
// with a design that touches the data set once:
//   * the table is stored on the device in SHUFFLED order, so every fold is a contiguous row range
//     and its training set is "everything else" (two segments);
//   * one pass accumulates, per fold, the column sums and centred cross products of ALL columns;
//     the training mean / covariance of any (fold, variable subset) is then O(d^2) host arithmetic
//     (total - fold), which gives every bandwidth H = k * cov and every LinearGaussianCPD fit
//     without reading the data again;
//   * all (candidate, fold) CKDE jobs of a call are whitened by ONE kernel launch and scored by ONE
//     multi-job launch of the pair kernel (pair_kernel.cuh), finalised and reduced by two more.
#include <cuda_runtime.h>
#include <math.h>

#include <algorithm>
#include <limits>
#include <string>
#include <vector>

#include "internal.h"

using pbn::PairJob;

struct pbn_cv {
    std::vector<pbn_cv*> rep;  // replicas on ctx->peers (multi-device context): the (item, fold) jobs are dealt over them
    pbn_ctx* ctx = nullptr;
    pbn_table* tbl = nullptr;  // shuffled-order copy (owned)
    int k = 0;
    int C = 0;
    int64_t n = 0;
    std::vector<int64_t> limits;   // k + 1
    std::vector<double> centre;    // [C] centring point (global mean)
    std::vector<double> S;         // [k][C]     per-fold sums of (x - centre)
    std::vector<double> G;         // [k][C][C]  per-fold sums of (x_a - centre_a)(x_b - centre_b)
    std::vector<double> S_tot, G_tot;
};

namespace {

#include "batch_kernels.cuh"

// out[c][r] = in[c][idx[r]]  (grid.y = column): the shuffled-order column store
template <typename T>
__global__ void gather_rows_kernel(const T* __restrict__ in, int64_t in_stride, const int32_t* __restrict__ idx, int64_t n,
                                   T* __restrict__ out, int64_t out_stride) {
    const T* src = in + (int64_t)blockIdx.y * in_stride;
    T* dst = out + (int64_t)blockIdx.y * out_stride;
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x)
        dst[r] = src[idx[r]];
}

// column sums of the whole table: partial[c][block]
template <typename T>
__global__ void colsum_all_kernel(const T* __restrict__ data, int64_t stride, int64_t n, double* __restrict__ partial) {
    __shared__ double sh[32];
    const T* x = data + (int64_t)blockIdx.y * stride;
    double s = 0;
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x)
        s += static_cast<double>(x[r]);
    double tot = block_sum(s, sh);
    if (threadIdx.x == 0) partial[(int64_t)blockIdx.y * gridDim.x + blockIdx.x] = tot;
}

// Per-fold centred sums and cross products of a 4 x 4 tile of column pairs.
//   grid = (blocks, tile pairs (ta <= tb), folds);  partial[((fold * npairs + pair) * blocks + block) * 24 + q]
//   q < 16: products a*4+b;  q = 16..19: sums of tile ta;  q = 20..23: sums of tile tb.
struct GramParams {
    const void* data;
    int64_t stride;
    int C;
    int ntiles;
    const int64_t* limits;  // device, k + 1
    const double* centre;   // device, C
    double* partial;
    int pair0;              // first tile pair of this launch (the host loops over chunks of pairs)
};

template <typename T>
__global__ void fold_gram_kernel(GramParams P) {
    __shared__ double sh[32];
    // decode the tile pair
    int pair = P.pair0 + blockIdx.y, ta = 0;
    while (pair >= P.ntiles - ta) { pair -= P.ntiles - ta; ++ta; }
    int tb = ta + pair;
    const int fold = blockIdx.z;
    const int64_t r0 = P.limits[fold], r1 = P.limits[fold + 1];
    const T* xa[4];
    const T* xb[4];
    double ca[4], cb[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        int a = ta * 4 + q, b = tb * 4 + q;
        xa[q] = a < P.C ? static_cast<const T*>(P.data) + (int64_t)a * P.stride : nullptr;
        xb[q] = b < P.C ? static_cast<const T*>(P.data) + (int64_t)b * P.stride : nullptr;
        ca[q] = a < P.C ? P.centre[a] : 0.0;
        cb[q] = b < P.C ? P.centre[b] : 0.0;
    }
    double acc[24];
#pragma unroll
    for (int q = 0; q < 24; ++q) acc[q] = 0.0;
    for (int64_t r = r0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < r1; r += (int64_t)gridDim.x * blockDim.x) {
        double va[4], vb[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            va[q] = xa[q] ? static_cast<double>(xa[q][r]) - ca[q] : 0.0;
            vb[q] = xb[q] ? static_cast<double>(xb[q][r]) - cb[q] : 0.0;
        }
#pragma unroll
        for (int a = 0; a < 4; ++a) {
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a * 4 + b] = fma(va[a], vb[b], acc[a * 4 + b]);
            acc[16 + a] += va[a];
            acc[20 + a] += vb[a];
        }
    }
    int npairs = gridDim.y;
    double* dst = P.partial + (((int64_t)fold * npairs + blockIdx.y) * gridDim.x + blockIdx.x) * 24;
    for (int q = 0; q < 24; ++q) {
        double tot = block_sum(acc[q], sh);
        if (threadIdx.x == 0) dst[q] = tot;
    }
}

// ---- batched whitening: one launch for all (candidate, fold) jobs of a chunk -------------------
struct WhitenJob {
    const void* cols[kMaxFast];  // column base pointers (shuffled table), internal variable order
    double W[kMaxFast * (kMaxFast + 1) / 2];  // packed lower triangle, row-major, includes the unit scale
    double mu[kMaxFast];
    long long t0, t1;  // rows [t0, t1) are the test fold, all other rows the training set
    void* y_train;     // AoS [n - (t1 - t0)][D]
    void* y_test;      // AoS [t1 - t0][D]
    float* bound;      // [0] max |coordinate| over training rows, [1] over test rows
    double* nrm_train; // f64 only (else null): -sum_{c<dn} y_c^2 per row, for the pair kernel's dot-product form
    double* nrm_test;
    int dn;
};

template <typename T, int D>
__global__ void whiten_batch_kernel(const WhitenJob* __restrict__ jobs, long long n) {
    const WhitenJob& jb = jobs[blockIdx.y];
    long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    float mx_tr = 0.f, mx_te = 0.f;
    if (r < n) {
        double x[D];
#pragma unroll
        for (int c = 0; c < D; ++c) x[c] = static_cast<double>(static_cast<const T*>(jb.cols[c])[r]) - jb.mu[c];
        const bool is_test = r >= jb.t0 && r < jb.t1;
        T* out = is_test ? static_cast<T*>(jb.y_test) + (r - jb.t0) * D
                         : static_cast<T*>(jb.y_train) + (r < jb.t0 ? r : r - (jb.t1 - jb.t0)) * D;
        float mx = 0.f;
        int w = 0;
        double nn = 0;
#pragma unroll
        for (int i = 0; i < D; ++i) {
            double s = 0;
#pragma unroll
            for (int k = 0; k <= i; ++k) s = fma(jb.W[w++], x[k], s);
            out[i] = static_cast<T>(s);
            if (i < jb.dn) nn = fma(-s, s, nn);
            float a = fabsf(static_cast<float>(s));
            mx = (a > mx || a != a) ? (a != a ? INFINITY : a) : mx;  // NaN counts as unbounded
        }
        if (sizeof(T) == 8 && jb.nrm_train) {
            if (is_test) jb.nrm_test[r - jb.t0] = nn;
            else jb.nrm_train[r < jb.t0 ? r : r - (jb.t1 - jb.t0)] = nn;
        }
        if (is_test) mx_te = mx; else mx_tr = mx;
    }
    for (int o = 16; o > 0; o >>= 1) {
        mx_tr = fmaxf(mx_tr, __shfl_xor_sync(0xffffffffu, mx_tr, o));
        mx_te = fmaxf(mx_te, __shfl_xor_sync(0xffffffffu, mx_te, o));
    }
    if ((threadIdx.x & 31) == 0) {
        // non-negative floats order like their bit patterns
        // issued only when the value can still raise the maximum (see whiten_kernel in runtime.cu)
        if (mx_tr > 0.f) {
            const int v = __float_as_int(mx_tr * 1.0001f);
            if (v > *reinterpret_cast<volatile int*>(jb.bound)) atomicMax(reinterpret_cast<int*>(jb.bound), v);
        }
        if (mx_te > 0.f) {
            const int v = __float_as_int(mx_te * 1.0001f);
            if (v > *reinterpret_cast<volatile int*>(jb.bound + 1)) atomicMax(reinterpret_cast<int*>(jb.bound + 1), v);
        }
    }
}

template <typename T>
cudaError_t launch_whiten_batch(int d, const WhitenJob* jobs, int n_jobs, long long n, cudaStream_t st) {
    dim3 grid((unsigned)((n + 255) / 256), (unsigned)n_jobs);
    switch (d) {
        case 1: whiten_batch_kernel<T, 1><<<grid, 256, 0, st>>>(jobs, n); break;
        case 2: whiten_batch_kernel<T, 2><<<grid, 256, 0, st>>>(jobs, n); break;
        case 3: whiten_batch_kernel<T, 3><<<grid, 256, 0, st>>>(jobs, n); break;
        case 4: whiten_batch_kernel<T, 4><<<grid, 256, 0, st>>>(jobs, n); break;
        case 5: whiten_batch_kernel<T, 5><<<grid, 256, 0, st>>>(jobs, n); break;
        case 6: whiten_batch_kernel<T, 6><<<grid, 256, 0, st>>>(jobs, n); break;
        case 7: whiten_batch_kernel<T, 7><<<grid, 256, 0, st>>>(jobs, n); break;
        case 8: whiten_batch_kernel<T, 8><<<grid, 256, 0, st>>>(jobs, n); break;
        case 9: whiten_batch_kernel<T, 9><<<grid, 256, 0, st>>>(jobs, n); break;
        case 10: whiten_batch_kernel<T, 10><<<grid, 256, 0, st>>>(jobs, n); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

// ---- host statistics -----------------------------------------------------------------------------
// Moments of the rows of `cv` EXCLUDING fold `f` restricted to `vars` (d of them):
//   mean[d], C[d*d] = sum over training rows of (x_a - mean_a)(x_b - mean_b)   (not yet divided).
void train_moments(const pbn_cv* cv, int f, const int* vars, int d, int64_t* n_out, double* mean, double* Cm) {
    const int C = cv->C;
    const int64_t nf = cv->limits[f + 1] - cv->limits[f];
    const int64_t ntr = cv->n - nf;
    const double* Sf = &cv->S[(size_t)f * C];
    const double* Gf = &cv->G[(size_t)f * C * C];
    double delta[PBN_MAX_DIM];
    for (int a = 0; a < d; ++a) {
        delta[a] = (cv->S_tot[vars[a]] - Sf[vars[a]]) / (double)ntr;
        mean[a] = cv->centre[vars[a]] + delta[a];
    }
    for (int a = 0; a < d; ++a)
        for (int b = 0; b < d; ++b) {
            size_t ij = (size_t)vars[a] * C + vars[b];
            Cm[a + b * d] = (cv->G_tot[ij] - Gf[ij]) - (double)ntr * delta[a] * delta[b];
        }
    *n_out = ntr;
}

}  // namespace

constexpr double kMachineTol = 1.4901161193847656e-08;  // util/math_constants.hpp:30 (sqrt of DBL_EPSILON)

// Solves Cxx b = Cxy for p >= 3 parents (centred normal equations, the least-squares solution the reference
// gets from colPivHouseholderQr on [1 X], mle_LinearGaussianCPD.hpp:157-191): Cholesky with complete
// pivoting; directions whose pivot is negligible get a zero coefficient.
static void solve_spd_pivoted(const double* A, const double* rhs, int p, double* x) {
    std::vector<double> M(A, A + p * p), L(p * p, 0.0), y(p, 0.0);
    std::vector<int> perm(p);
    for (int i = 0; i < p; ++i) perm[i] = i;
    double max_diag = 0;
    for (int i = 0; i < p; ++i) max_diag = std::max(max_diag, M[i + i * p]);
    int rank = 0;
    for (int j = 0; j < p; ++j) {
        int piv = j;
        for (int i = j + 1; i < p; ++i)
            if (M[perm[i] + perm[i] * p] > M[perm[piv] + perm[piv] * p]) piv = i;
        std::swap(perm[j], perm[piv]);
        // rows of L follow the permutation: swap the already computed parts
        for (int c = 0; c < j; ++c) std::swap(L[j + c * p], L[piv + c * p]);
        double djj = M[perm[j] + perm[j] * p];
        if (!(djj > max_diag * p * 2.220446049250313e-16) || !(djj > 0)) break;
        double ljj = sqrt(djj);
        L[j + j * p] = ljj;
        for (int i = j + 1; i < p; ++i) {
            double s = A[perm[i] + perm[j] * p];
            for (int c = 0; c < j; ++c) s -= L[i + c * p] * L[j + c * p];
            L[i + j * p] = s / ljj;
            M[perm[i] + perm[i] * p] -= L[i + j * p] * L[i + j * p];
        }
        ++rank;
    }
    for (int i = 0; i < rank; ++i) {
        double s = rhs[perm[i]];
        for (int c = 0; c < i; ++c) s -= L[i + c * p] * y[c];
        y[i] = s / L[i + i * p];
    }
    std::vector<double> z(p, 0.0);
    for (int i = rank - 1; i >= 0; --i) {
        double s = y[i];
        for (int c = i + 1; c < rank; ++c) s -= L[c + i * p] * z[c];
        z[i] = s / L[i + i * p];
    }
    for (int i = 0; i < p; ++i) x[perm[i]] = i < rank ? z[i] : 0.0;
}

// MLE<LinearGaussianCPD> (mle_LinearGaussianCPD.hpp:11-221) from centred moments.
// vars[0] = variable, vars[1..] = parents; beta[p + 1]; returns the variance.
double lg_fit_from_moments(int64_t rows, int p, const double* mean, const double* Cm, double* beta) {
    const int d = p + 1;
    const double inf = std::numeric_limits<double>::infinity();
    auto c = [&](int a, int b) { return Cm[a + b * d]; };
    const double my = mean[0];
    if (p == 0) {
        beta[0] = my;
        if (rows == 1) return inf;
        return c(0, 0) / (double)(rows - 1);
    }
    if (p == 1) {
        double var_x = c(1, 1) / (double)(rows - 1);
        if (var_x < kMachineTol) {
            beta[0] = my;
            beta[1] = 0;
            return rows <= 2 ? inf : c(0, 0) / (double)(rows - 2);
        }
        double cov_yx = c(0, 1) / (double)(rows - 1);
        double b = cov_yx / var_x;
        beta[0] = my - b * mean[1];
        beta[1] = b;
        if (rows <= 2) return inf;
        return (c(0, 0) - 2 * b * c(0, 1) + b * b * c(1, 1)) / (double)(rows - 2);
    }
    if (p == 2) {
        double var_x1 = c(1, 1) / (double)(rows - 1);
        bool singular1 = var_x1 < kMachineTol;
        double var_x2 = c(2, 2) / (double)(rows - 1);
        double cov_xx = c(1, 2) / (double)(rows - 1);
        bool singular2 = var_x2 < kMachineTol || fabs(cov_xx / sqrt(var_x1 * var_x2)) > (1 - kMachineTol);
        double b1 = 0, b2 = 0;
        if (singular1) {
            if (!singular2) b2 = (c(0, 2) / (double)(rows - 1)) / var_x2;
        } else if (singular2) {
            b1 = (c(0, 1) / (double)(rows - 1)) / var_x1;
        } else {
            double cov_yx1 = c(0, 1) / (double)(rows - 1), cov_yx2 = c(0, 2) / (double)(rows - 1);
            double den = var_x1 * var_x2 - cov_xx * cov_xx;
            b1 = (var_x2 * cov_yx1 - cov_xx * cov_yx2) / den;
            b2 = (cov_yx2 - b1 * cov_xx) / var_x2;
        }
        beta[0] = my - b1 * mean[1] - b2 * mean[2];
        beta[1] = b1;
        beta[2] = b2;
        if (rows <= 3) return inf;
        double rss = c(0, 0) + b1 * b1 * c(1, 1) + b2 * b2 * c(2, 2) - 2 * b1 * c(0, 1) - 2 * b2 * c(0, 2) + 2 * b1 * b2 * c(1, 2);
        return rss / (double)(rows - 3);
    }
    std::vector<double> A(p * p), rhs(p), b(p);
    for (int i = 0; i < p; ++i) {
        rhs[i] = c(0, i + 1);
        for (int j = 0; j < p; ++j) A[i + j * p] = c(i + 1, j + 1);
    }
    solve_spd_pivoted(A.data(), rhs.data(), p, b.data());
    double a = my;
    for (int i = 0; i < p; ++i) a -= b[i] * mean[i + 1];
    beta[0] = a;
    for (int i = 0; i < p; ++i) beta[i + 1] = b[i];
    if (rows <= d) return inf;
    double rss = c(0, 0);
    for (int i = 0; i < p; ++i) {
        rss -= 2 * b[i] * rhs[i];
        for (int j = 0; j < p; ++j) rss += b[i] * b[j] * A[i + j * p];
    }
    return rss / (double)(rows - d);
}

namespace {

// slogl of a fitted LinearGaussianCPD over the rows of fold f (LinearGaussianCPD.cpp:92-149), from the
// fold's centred sums / cross products.
double lg_fold_slogl(const pbn_cv* cv, int f, const int* vars, int p, const double* beta, double variance) {
    const int C = cv->C;
    const int64_t m = cv->limits[f + 1] - cv->limits[f];
    if (m == 0) return 0.0;
    if (!(variance < std::numeric_limits<double>::infinity())) return -std::numeric_limits<double>::infinity();
    const double* Sf = &cv->S[(size_t)f * C];
    const double* Gf = &cv->G[(size_t)f * C * C];
    auto g = [&](int a, int b) { return Gf[(size_t)vars[a] * C + vars[b]]; };
    // residual r = (y - cy) - sum_j b_j (x_j - c_j) - c0,  c0 = beta0 - cy + sum_j b_j c_j
    double c0 = beta[0] - cv->centre[vars[0]];
    for (int j = 1; j <= p; ++j) c0 += beta[j] * cv->centre[vars[j]];
    double rss = g(0, 0), lin = Sf[vars[0]];
    for (int i = 1; i <= p; ++i) {
        rss -= 2 * beta[i] * g(0, i);
        lin -= beta[i] * Sf[vars[i]];
        for (int j = 1; j <= p; ++j) rss += beta[i] * beta[j] * g(i, j);
    }
    rss += -2 * c0 * lin + (double)m * c0 * c0;
    const double log2pi = 1.8378770664093454836;
    return -0.5 * rss / variance - (double)m * (0.5 * log(variance) + 0.5 * log2pi);
}

struct CkdeJobHost {
    int item, fold;
    int64_t n_train, m;
    double lognorm_joint, lognorm_marg;
};

}  // namespace

extern "C" {

static int cv_create_one(pbn_ctx* ctx, const pbn_table* tbl, const int32_t* indices, int64_t n, const int32_t* limits, int k,
                         pbn_cv** out);

int pbn_cv_create(pbn_ctx* ctx, const pbn_table* tbl, const int32_t* indices, int64_t n, const int32_t* limits, int k,
                  pbn_cv** out) {
    if (!ctx || !tbl || !indices || !limits || !out) return set_error(PBN_ERR_ARG, "null argument");
    if (!pbn_replicated(ctx, tbl)) return cv_create_one(ctx, tbl, indices, n, limits, k, out);
    // multi-device context: the shuffled-order table (16 MB at config 4) and the fold statistics on every device
    const int nd = pbn_num_devices(ctx);
    std::vector<pbn_cv*> c(nd, nullptr);
    int rc = pbn_run_on_devices(nd, [&](int i) {
        return cv_create_one(pbn_device_ctx(ctx, i), pbn_replica(const_cast<pbn_table*>(tbl), i), indices, n, limits, k, &c[i]);
    });
    if (rc != PBN_OK) {
        for (pbn_cv* q : c)
            if (q) pbn_cv_free(q);
        return rc;
    }
    c[0]->rep.assign(c.begin() + 1, c.end());
    *out = c[0];
    return PBN_OK;
}

static int cv_create_one(pbn_ctx* ctx, const pbn_table* tbl, const int32_t* indices, int64_t n, const int32_t* limits, int k,
                         pbn_cv** out) {
    if (!ctx || !tbl || !indices || !limits || !out) return set_error(PBN_ERR_ARG, "null argument");
    if (k < 1 || n < 1) return set_error(PBN_ERR_ARG, "invalid fold structure");
    if (limits[0] != 0 || limits[k] != n) return set_error(PBN_ERR_ARG, "fold limits must span [0, n]");
    for (int f = 0; f < k; ++f)
        if (limits[f + 1] < limits[f]) return set_error(PBN_ERR_ARG, "fold limits must be non-decreasing");
    for (int64_t i = 0; i < n; ++i)
        if (indices[i] < 0 || indices[i] >= tbl->nrows) return set_error(PBN_ERR_ARG, "row index out of range");
    DevSetter ds(ctx->device);
    cudaStream_t st = ctx->stream;
    const int C = tbl->ncols;
    const size_t es = elem_size(tbl->dtype);
    const bool f64 = tbl->dtype == PBN_F64;

    pbn_cv* cv = new pbn_cv();
    cv->ctx = ctx;
    cv->k = k;
    cv->C = C;
    cv->n = n;
    cv->limits.assign(limits, limits + k + 1);
    pbn_table* t = new pbn_table();
    t->ctx = ctx;
    t->ncols = C;
    t->nrows = n;
    t->dtype = tbl->dtype;
    t->stride = (n + 63) / 64 * 64 + 64;
    t->data = nullptr;
    cv->tbl = t;
    auto fail = [&](int rc) {
        if (t->data) cudaFreeAsync(t->data, st);
        delete t;
        delete cv;
        return rc;
    };
#define CV_TRY(expr)                                                                                                   \
    do {                                                                                                               \
        cudaError_t e__ = (expr);                                                                                      \
        if (e__ != cudaSuccess)                                                                                        \
            return fail(set_error(PBN_ERR_CUDA, std::string("CUDA error ") + cudaGetErrorName(e__) + " (" +            \
                                                    cudaGetErrorString(e__) + ") at " #expr));                         \
    } while (0)

    // 1. shuffled-order column store
    int32_t* d_idx = nullptr;
    CV_TRY(cudaMallocAsync(&t->data, (size_t)C * t->stride * es, st));
    CV_TRY(cudaMemsetAsync(t->data, 0, (size_t)C * t->stride * es, st));
    CV_TRY(cudaMallocAsync(&d_idx, (size_t)n * sizeof(int32_t), st));
    CV_TRY(cudaMemcpyAsync(d_idx, indices, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    ctx->h2d += n * 4;
    {
        int bx = (int)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * 8);
        dim3 grid(bx, C);
        if (f64)
            gather_rows_kernel<double><<<grid, 256, 0, st>>>(static_cast<const double*>(tbl->data), tbl->stride, d_idx, n,
                                                             static_cast<double*>(t->data), t->stride);
        else
            gather_rows_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(tbl->data), tbl->stride, d_idx, n,
                                                            static_cast<float*>(t->data), t->stride);
        ctx->launches++;
        CV_TRY(cudaGetLastError());
    }
    CV_TRY(cudaFreeAsync(d_idx, st));

    // 2. centring point: the global column means
    const int blocks = (int)std::min<int64_t>((n + 255) / 256, 64);
    {
        double* d_part = nullptr;
        CV_TRY(cudaMallocAsync(&d_part, (size_t)C * blocks * sizeof(double), st));
        dim3 grid(blocks, C);
        if (f64) colsum_all_kernel<double><<<grid, 256, 0, st>>>(static_cast<const double*>(t->data), t->stride, n, d_part);
        else colsum_all_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(t->data), t->stride, n, d_part);
        ctx->launches++;
        CV_TRY(cudaGetLastError());
        std::vector<double> h((size_t)C * blocks);
        CV_TRY(cudaMemcpyAsync(h.data(), d_part, h.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
        CV_TRY(cudaStreamSynchronize(st));
        ctx->d2h += (int64_t)h.size() * 8;
        CV_TRY(cudaFreeAsync(d_part, st));
        cv->centre.resize(C);
        for (int c = 0; c < C; ++c) {
            double s = 0;
            for (int b = 0; b < blocks; ++b) s += h[(size_t)c * blocks + b];
            cv->centre[c] = s / (double)n;
        }
    }

    // 3. per-fold centred sums and cross products of all column pairs
    {
        const int ntiles = (C + 3) / 4;
        const int npairs = ntiles * (ntiles + 1) / 2;
        const int fb = (int)std::max<int64_t>(1, std::min<int64_t>(32, (n / k + 255) / 256));
        std::vector<int64_t> lim64(cv->limits);
        int64_t* d_lim = nullptr;
        double* d_centre = nullptr;
        double* d_part = nullptr;
        // tile pairs are processed in chunks: grid.y stays far below the 65535 launch limit and the partial buffer
        // (device and host) is bounded however many columns the frame has
        const int chunk = std::min(npairs, 2048);
        size_t np = (size_t)k * chunk * fb * 24;
        CV_TRY(cudaMallocAsync(&d_lim, (k + 1) * sizeof(int64_t), st));
        CV_TRY(cudaMallocAsync(&d_centre, C * sizeof(double), st));
        CV_TRY(cudaMallocAsync(&d_part, np * sizeof(double), st));
        CV_TRY(cudaMemcpyAsync(d_lim, lim64.data(), (k + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
        CV_TRY(cudaMemcpyAsync(d_centre, cv->centre.data(), C * sizeof(double), cudaMemcpyHostToDevice, st));
        GramParams P;
        P.data = t->data;
        P.stride = t->stride;
        P.C = C;
        P.ntiles = ntiles;
        P.limits = d_lim;
        P.centre = d_centre;
        P.partial = d_part;
        std::vector<double> h(np);
        std::vector<double> hsum((size_t)k * npairs * 24);  // block partials added in block order
        for (int pair0 = 0; pair0 < npairs; pair0 += chunk) {
            const int np_here = std::min(chunk, npairs - pair0);
            P.pair0 = pair0;
            dim3 grid(fb, np_here, k);
            if (f64) fold_gram_kernel<double><<<grid, 256, 0, st>>>(P);
            else fold_gram_kernel<float><<<grid, 256, 0, st>>>(P);
            ctx->launches++;
            CV_TRY(cudaGetLastError());
            const size_t cnt = (size_t)k * np_here * fb * 24;
            CV_TRY(cudaMemcpyAsync(h.data(), d_part, cnt * sizeof(double), cudaMemcpyDeviceToHost, st));
            CV_TRY(cudaStreamSynchronize(st));
            ctx->d2h += (int64_t)cnt * 8;
            for (int f = 0; f < k; ++f)
                for (int pl = 0; pl < np_here; ++pl)
                    for (int q = 0; q < 24; ++q) {
                        double s = 0;
                        for (int b = 0; b < fb; ++b) s += h[(((size_t)f * np_here + pl) * fb + b) * 24 + q];
                        hsum[((size_t)f * npairs + pair0 + pl) * 24 + q] = s;
                    }
        }
        CV_TRY(cudaFreeAsync(d_lim, st));
        CV_TRY(cudaFreeAsync(d_centre, st));
        CV_TRY(cudaFreeAsync(d_part, st));
        cv->S.assign((size_t)k * C, 0.0);
        cv->G.assign((size_t)k * C * C, 0.0);
        for (int f = 0; f < k; ++f) {
            int pair = 0;
            for (int ta = 0; ta < ntiles; ++ta)
                for (int tb = ta; tb < ntiles; ++tb, ++pair) {
                    double acc[24];
                    for (int q = 0; q < 24; ++q) acc[q] = hsum[((size_t)f * npairs + pair) * 24 + q];
                    for (int a = 0; a < 4; ++a)
                        for (int b = 0; b < 4; ++b) {
                            int ca = ta * 4 + a, cb = tb * 4 + b;
                            if (ca >= C || cb >= C) continue;
                            cv->G[((size_t)f * C + ca) * C + cb] = acc[a * 4 + b];
                            cv->G[((size_t)f * C + cb) * C + ca] = acc[a * 4 + b];
                        }
                    if (ta == tb)
                        for (int a = 0; a < 4; ++a)
                            if (ta * 4 + a < C) cv->S[(size_t)f * C + ta * 4 + a] = acc[16 + a];
                }
        }
        cv->S_tot.assign(C, 0.0);
        cv->G_tot.assign((size_t)C * C, 0.0);
        for (int f = 0; f < k; ++f) {
            for (int c = 0; c < C; ++c) cv->S_tot[c] += cv->S[(size_t)f * C + c];
            for (size_t q = 0; q < (size_t)C * C; ++q) cv->G_tot[q] += cv->G[(size_t)f * C * C + q];
        }
    }
#undef CV_TRY
    *out = cv;
    return PBN_OK;
}

int pbn_cv_free(pbn_cv* cv) {
    if (!cv) return PBN_OK;
    for (pbn_cv* r : cv->rep) pbn_cv_free(r);
    cv->rep.clear();
    DevSetter ds(cv->ctx->device);
    if (cv->tbl) {
        if (cv->tbl->data) cudaFreeAsync(cv->tbl->data, cv->ctx->stream);
        delete cv->tbl;
    }
    delete cv;
    return PBN_OK;
}

const pbn_table* pbn_cv_table(const pbn_cv* cv) { return cv ? cv->tbl : nullptr; }
int pbn_cv_folds(const pbn_cv* cv) { return cv ? cv->k : 0; }

int pbn_cv_train_moments(const pbn_cv* cv, int fold, const int* vars, int d, double* mean_out, double* cov_out) {
    if (!cv || !vars) return set_error(PBN_ERR_ARG, "null argument");
    if (fold < 0 || fold >= cv->k) return set_error(PBN_ERR_ARG, "fold out of range");
    PBN_TRY(check_cols(cv->tbl, vars, d));
    std::vector<double> mean(d), Cm((size_t)d * d);
    int64_t ntr;
    train_moments(cv, fold, vars, d, &ntr, mean.data(), Cm.data());
    if (mean_out) std::copy(mean.begin(), mean.end(), mean_out);
    if (cov_out) {
        if (ntr < 2) return set_error(PBN_ERR_ARG, "covariance needs at least 2 rows");
        for (int i = 0; i < d * d; ++i) cov_out[i] = Cm[i] / (double)(ntr - 1);
    }
    return PBN_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------
// scoring
// ---------------------------------------------------------------------------------------------------
namespace {

// bandwidth of (item, fold) from the fold statistics: NormalReferenceRule / ScottsBandwidth ::bandwidth
// (kde/NormalReferenceRule.hpp:109-134, ScottsBandwidth.hpp:91-117) on the fold's training rows.
int fold_bandwidth(const pbn_cv* cv, int f, const int* vars, int d, int rule, int64_t* ntr, double* mean, double* H) {
    std::vector<double> Cm((size_t)d * d);
    train_moments(cv, f, vars, d, ntr, mean, Cm.data());
    const int dtype = cv->tbl->dtype;
    if (*ntr <= d)
        return set_error(PBN_ERR_SINGULAR, "Bandwidth matrix of " + std::to_string(d) + " variables " + var_list(vars, d) +
                                               " cannot be estimated with " + std::to_string(*ntr) + " instances");
    std::vector<double> cov((size_t)d * d);
    for (int i = 0; i < d * d; ++i) {
        cov[i] = Cm[i] / (double)(*ntr - 1);
        if (dtype == PBN_F32) cov[i] = (double)(float)cov[i];  // the reference holds the covariance in the data type
    }
    if (!is_psd(cov.data(), d, dtype))
        return set_error(PBN_ERR_SINGULAR,
                         "Covariance matrix for variables " + var_list(vars, d) + " is not positive-definite.");
    double N = (double)*ntr, dd = (double)d, kfac;
    if (dtype == PBN_F32) N = (double)(float)N;
    if (rule == PBN_BW_NORMAL_REFERENCE) kfac = pow(4. / (N * (dd + 2.)), 2. / (dd + 4.));
    else if (rule == PBN_BW_SCOTT) kfac = pow(N, -2. / (dd + 4.));
    else return set_error(PBN_ERR_ARG, "unknown bandwidth rule");
    for (int i = 0; i < d * d; ++i) H[i] = kfac * cov[i];
    return PBN_OK;
}

struct CvJob {
    int item, fold;
};

// scores the CKDE (item, fold) jobs `jobs[sel[.]]` (all items with the same number of variables d <= kMaxFast): ONE
// whitening launch and ONE multi-job pair-kernel launch per chunk of jobs; job_scores[sel[q]] = slogl of that fold
int score_ckde_jobs(pbn_ctx* ctx, pbn_cv* cv, const pbn_cv_item* items, const CvJob* jobs, const std::vector<int>& sel, int d,
                    double* job_scores, int* status, std::string* first_error) {
    DevSetter ds(ctx->device);
    cudaStream_t st = ctx->stream;
    const pbn_table* tbl = cv->tbl;
    const bool f64 = tbl->dtype == PBN_F64;
    const size_t es = elem_size(tbl->dtype);
    const bool ckde = d >= 2;
    const int TILE = f64 ? pbn::pair_tile_f64(d) : pbn::pair_tile_f32(d);
    const int TB = f64 ? pbn::pair_tb_for_f64(d, ckde) : pbn::pair_tb_for_f32(d, ckde);
    const double unit = unit_scale(tbl->dtype);
    const double cscale = sqrt(0.5 * unit);
    const double log2pi = 1.8378770664093454836;

    // bytes of whitened rows per (item, fold) job
    auto job_bytes = [&](int64_t ntr, int64_t m) {
        size_t a = ((size_t)ntr * d * es + 16 + 255) / 256 * 256;
        size_t b = ((size_t)m * d * es + 16 + 255) / 256 * 256;
        if (f64) b += ((size_t)ntr * 8 + 16 + 255) / 256 * 256 + ((size_t)m * 8 + 255) / 256 * 256;  // row norms
        return a + b;
    };
    const size_t budget = (size_t)3 << 30;  // whitened-row scratch per chunk

    for (size_t c0 = 0; c0 < sel.size();) {
        // a chunk: as many consecutive jobs as fit the scratch budget (at least one)
        size_t c1 = c0, bytes = 0;
        while (c1 < sel.size()) {
            const int f = jobs[sel[c1]].fold;
            const int64_t m = cv->limits[f + 1] - cv->limits[f];
            const size_t jb = job_bytes(cv->n - m, m);
            if (c1 > c0 && bytes + jb > budget) break;
            bytes += jb;
            ++c1;
        }
        std::vector<WhitenJob> wj;
        std::vector<PairJob> pj;
        std::vector<FinJob> fj;
        std::vector<CkdeJobHost> hj;
        std::vector<size_t> y_off_train, y_off_test;
        size_t ybytes = 0;
        long long out_total = 0;
        for (size_t q = c0; q < c1; ++q) {
            const int jid = sel[q];
            const int it = jobs[jid].item, f = jobs[jid].fold;
            const pbn_cv_item& item = items[it];
            if (status[it] != PBN_OK) continue;  // an earlier fold of this item already failed
            // internal order: variable last, so that the marginal's whitened coordinates are a prefix of the joint's
            int vars[kMaxFast], pvars[kMaxFast];
            for (int i = 0; i < d; ++i) vars[i] = item.vars[i];
            for (int i = 0; i < d; ++i) pvars[i] = ckde ? vars[(i + 1) % d] : vars[i];
            {
                const int64_t m = cv->limits[f + 1] - cv->limits[f];
                int64_t ntr;
                double mean[kMaxFast], H[kMaxFast * kMaxFast], L[kMaxFast * kMaxFast], Winv[kMaxFast * kMaxFast];
                int rc = fold_bandwidth(cv, f, pvars, d, item.rule, &ntr, mean, H);
                if (rc == PBN_OK && !chol_lower(H, d, L)) rc = set_error(PBN_ERR_SINGULAR, "bandwidth matrix is not positive definite");
                if (rc != PBN_OK) {
                    status[it] = rc;
                    if (first_error->empty()) *first_error = pbn_last_error();
                    continue;
                }
                if (m == 0) continue;  // an empty test fold adds 0
                tri_inverse_rowmajor(L, d, Winv);
                WhitenJob w;
                memset(&w, 0, sizeof(w));
                for (int i = 0; i < d; ++i) w.cols[i] = col_ptr(tbl, pvars[i]);
                int wi = 0;
                for (int i = 0; i < d; ++i)
                    for (int j = 0; j <= i; ++j) w.W[wi++] = cscale * Winv[i * d + j];
                for (int i = 0; i < d; ++i) w.mu[i] = mean[i];
                w.t0 = cv->limits[f];
                w.t1 = cv->limits[f + 1];
                double slog = 0, slog_m = 0;
                for (int i = 0; i < d; ++i) {
                    slog += log(L[i + i * d]);
                    if (i < d - 1) slog_m += log(L[i + i * d]);
                }
                CkdeJobHost h;
                h.item = jid;  // position in `jobs`: where the fold's slogl goes
                h.fold = f;
                h.n_train = ntr;
                h.m = m;
                h.lognorm_joint = -slog - 0.5 * d * log2pi - log((double)ntr);
                h.lognorm_marg = -slog_m - 0.5 * (d - 1) * log2pi - log((double)ntr);
                y_off_train.push_back(ybytes);
                ybytes += ((size_t)ntr * d * es + 16 + 255) / 256 * 256;
                y_off_test.push_back(ybytes);
                ybytes += ((size_t)m * d * es + 16 + 255) / 256 * 256;
                if (f64) ybytes += ((size_t)ntr * 8 + 16 + 255) / 256 * 256 + ((size_t)m * 8 + 255) / 256 * 256;  // row norms
                FinJob fin;
                fin.lognorm_joint = h.lognorm_joint;
                fin.lognorm_marg = h.lognorm_marg;
                fin.out_off = 0;  // set below
                wj.push_back(w);
                fj.push_back(fin);
                hj.push_back(h);
                PairJob p;
                memset(&p, 0, sizeof(p));
                p.n_train = ntr;
                p.m = m;
                pj.push_back(p);
            }
        }
        c0 = c1;
        // out offsets are recomputed here because failed items may have removed jobs
        const int J = (int)hj.size();
        if (J == 0) continue;
        out_total = 0;
        for (int j = 0; j < J; ++j) {
            fj[j].out_off = out_total;
            out_total += hj[j].m;
        }

        // unit schedule of the multi-job launch
        long long U = 0;
        long long max_m = 0;
        for (int j = 0; j < J; ++j) {
            pj[j].n_train_tiles = (int)((hj[j].n_train + TILE - 1) / TILE);
            pj[j].n_test_tiles = (int)((hj[j].m + TB - 1) / TB);
            pj[j].unit_begin = U;
            U += (long long)pj[j].n_train_tiles * pj[j].n_test_tiles;
            max_m = std::max<long long>(max_m, hj[j].m);
        }
        int grid = (int)std::min<long long>(U, (long long)ctx->sm_count * (f64 ? pbn::pair_ctas_per_sm_f64() : pbn::pair_ctas_per_sm_f32()));
        long long upb = (U + grid - 1) / grid;
        grid = (int)((U + upb - 1) / upb);
        const int n_acc = ckde ? 2 : 1;
        size_t part_elems = 0;
        std::vector<size_t> part_off(J);
        for (int j = 0; j < J; ++j) {
            pj[j].slots = (int)std::min<long long>((pj[j].n_train_tiles + upb - 1) / upb + 1, grid);
            pj[j].m_pad = (hj[j].m + 31) / 32 * 32;
            part_off[j] = part_elems;
            part_elems += (size_t)n_acc * pj[j].slots * pj[j].m_pad;
        }

        // device scratch: one allocation
        size_t off = 0;
        auto carve = [&](size_t bytes) {
            size_t o = off;
            off += (bytes + 255) / 256 * 256;
            return o;
        };
        size_t o_y = carve(ybytes), o_bound = carve((size_t)J * 2 * sizeof(float)), o_wj = carve((size_t)J * sizeof(WhitenJob));
        size_t o_pj = carve((size_t)J * sizeof(PairJob)), o_fj = carve((size_t)J * sizeof(FinJob));
        size_t o_part = carve(part_elems * sizeof(double)), o_out = carve((size_t)out_total * sizeof(double));
        size_t o_flag = carve((size_t)out_total * sizeof(int2)), o_nflag = carve(256), o_sums = carve((size_t)J * sizeof(double));
        char* base = nullptr;
        PBN_CUDA_TRY(cudaMallocAsync(&base, off, st));
        for (int j = 0; j < J; ++j) {
            wj[j].y_train = base + o_y + y_off_train[j];
            wj[j].y_test = base + o_y + y_off_test[j];
            wj[j].bound = reinterpret_cast<float*>(base + o_bound) + 2 * j;
            if (f64) {
                char* after_test = static_cast<char*>(wj[j].y_test) + ((size_t)hj[j].m * d * es + 16 + 255) / 256 * 256;
                wj[j].nrm_train = reinterpret_cast<double*>(after_test);
                wj[j].nrm_test = reinterpret_cast<double*>(after_test + ((size_t)hj[j].n_train * 8 + 16 + 255) / 256 * 256);
                wj[j].dn = ckde ? d - 1 : d;
                pj[j].train_nrm = wj[j].nrm_train;
                pj[j].test_nrm = wj[j].nrm_test;
            }
            pj[j].train = wj[j].y_train;
            pj[j].test = wj[j].y_test;
            pj[j].bound_train = wj[j].bound;
            pj[j].bound_test = wj[j].bound + 1;
            pj[j].part = reinterpret_cast<double*>(base + o_part) + part_off[j];
        }
        PBN_CUDA_TRY(cudaMemsetAsync(base + o_bound, 0, (size_t)J * 2 * sizeof(float), st));
        PBN_CUDA_TRY(cudaMemsetAsync(base + o_nflag, 0, 256, st));
        PBN_CUDA_TRY(cudaMemcpyAsync(base + o_wj, wj.data(), (size_t)J * sizeof(WhitenJob), cudaMemcpyHostToDevice, st));
        PBN_CUDA_TRY(cudaMemcpyAsync(base + o_pj, pj.data(), (size_t)J * sizeof(PairJob), cudaMemcpyHostToDevice, st));
        PBN_CUDA_TRY(cudaMemcpyAsync(base + o_fj, fj.data(), (size_t)J * sizeof(FinJob), cudaMemcpyHostToDevice, st));
        ctx->h2d += (int64_t)J * (sizeof(WhitenJob) + sizeof(PairJob) + sizeof(FinJob));
        const WhitenJob* d_wj = reinterpret_cast<const WhitenJob*>(base + o_wj);
        const PairJob* d_pj = reinterpret_cast<const PairJob*>(base + o_pj);
        const FinJob* d_fj = reinterpret_cast<const FinJob*>(base + o_fj);
        double* d_out = reinterpret_cast<double*>(base + o_out);
        int2* d_flag = reinterpret_cast<int2*>(base + o_flag);
        int* d_nflag = reinterpret_cast<int*>(base + o_nflag);
        double* d_sums = reinterpret_cast<double*>(base + o_sums);

        PBN_CUDA_TRY(f64 ? launch_whiten_batch<double>(d, d_wj, J, cv->n, st) : launch_whiten_batch<float>(d, d_wj, J, cv->n, st));
        ctx->launches++;
        cudaEvent_t ev0 = nullptr, ev1 = nullptr;
        if (ctx->timing) {
            PBN_CUDA_TRY(cudaEventCreate(&ev0));
            PBN_CUDA_TRY(cudaEventCreate(&ev1));
            PBN_CUDA_TRY(cudaEventRecord(ev0, st));
        }
        PBN_CUDA_TRY(f64 ? pbn::launch_pair_f64(d, ckde, d_pj, J, U, upb, grid, ctx->d_exp_tab, st)
                         : pbn::launch_pair_f32(d, ckde, d_pj, J, U, upb, grid, ctx->d_exp_tab, st));
        ctx->launches++;
        if (ctx->timing) {
            PBN_CUDA_TRY(cudaEventRecord(ev1, st));
            ctx->timed.emplace_back(ev0, ev1);
            for (int j = 0; j < J; ++j) ctx->pair_units += hj[j].n_train * hj[j].m * (ckde ? 2 : 1);
        }
        const double thresh = f64 ? ldexp(1.0, -900) : ldexp(1.0, -64);
        {
            dim3 fgrid((unsigned)std::min<long long>((max_m + 255) / 256, 64), (unsigned)J);
            finalize_batch_kernel<<<fgrid, 256, 0, st>>>(d_pj, d_fj, upb, TB, ckde ? 1 : 0, thresh, d_out, d_flag, d_nflag);
            ctx->launches++;
            PBN_CUDA_TRY(cudaGetLastError());
        }
        // rows whose unshifted sums underflowed (test rows of a fold far from all its training rows): the shifted second
        // pass of runtime.cu, job by job (round 1: one CTA per row).  Rare, so the flagged (job, row) list is read back.
        int nflag = 0;
        PBN_CUDA_TRY(cudaMemcpyAsync(&nflag, d_nflag, sizeof(int), cudaMemcpyDeviceToHost, st));
        PBN_CUDA_TRY(cudaStreamSynchronize(st));
        ctx->d2h += 4;
        Scratch sc(st);
        if (nflag > 0) {
            std::vector<int2> fl((size_t)nflag);
            PBN_CUDA_TRY(cudaMemcpyAsync(fl.data(), d_flag, (size_t)nflag * sizeof(int2), cudaMemcpyDeviceToHost, st));
            PBN_CUDA_TRY(cudaStreamSynchronize(st));
            ctx->d2h += (int64_t)nflag * 8;
            std::vector<std::vector<int>> by_job(J);
            for (const int2& e : fl) by_job[e.x].push_back(e.y);
            for (int j = 0; j < J; ++j) {
                std::vector<int>& rows = by_job[j];
                if (rows.empty()) continue;
                std::sort(rows.begin(), rows.end());
                const int cnt = (int)rows.size();
                rows.push_back(cnt);  // the count travels behind the list
                int* d_rows = nullptr;
                PBN_CUDA_TRY(sc.alloc(&d_rows, rows.size() * sizeof(int)));
                PBN_CUDA_TRY(cudaMemcpyAsync(d_rows, rows.data(), rows.size() * sizeof(int), cudaMemcpyHostToDevice, st));
                PBN_CUDA_TRY(cudaStreamSynchronize(st));  // `rows` is reused by the next job
                ctx->h2d += (int64_t)rows.size() * 4;
                ShiftPass A;
                A.train = wj[j].y_train;
                A.n_train = hj[j].n_train;
                A.test = wj[j].y_test;
                A.m_cap = cnt;
                A.d = d;
                A.ckde = ckde;
                A.f64 = f64;
                A.lognorm_joint = fj[j].lognorm_joint;
                A.lognorm_marg = fj[j].lognorm_marg;
                A.flagged = d_rows;
                A.n_flagged = d_rows + cnt;
                A.out = d_out + fj[j].out_off;
                A.n_row_kernel = nullptr;
                PBN_TRY(pbn_shift_pass(ctx, sc, A));
            }
        }
        segsum_kernel<<<J, 256, 0, st>>>(d_pj, d_fj, d_out, d_sums);
        ctx->launches++;
        PBN_CUDA_TRY(cudaGetLastError());
        std::vector<double> sums(J);
        PBN_CUDA_TRY(cudaMemcpyAsync(sums.data(), d_sums, (size_t)J * sizeof(double), cudaMemcpyDeviceToHost, st));
        PBN_CUDA_TRY(cudaStreamSynchronize(st));
        ctx->d2h += (int64_t)J * 8;
        ctx->last_fallback_rows = nflag;
        PBN_CUDA_TRY(cudaFreeAsync(base, st));
        for (int j = 0; j < J; ++j) job_scores[hj[j].item] = sums[j];
    }
    return PBN_OK;
}

}  // namespace

// One fold of one item, the unit everything below is scheduled in.  LinearGaussianCPD jobs are closed forms of the fold
// statistics (host); CKDE jobs of families up to kMaxFast variables are grouped by family size and batched
// (score_ckde_jobs); wider CKDEs take one fit + slogl per fold through the single-model entry points.
// Multi-device context: the batched CKDE jobs are dealt over the devices most expensive first (cost = training rows x
// test rows x variables), every device runs its share as its own batch on its replica of the shuffled table, from one
// host thread per device.  Each job's value is produced by exactly one device.
extern "C" int pbn_cv_score_jobs(pbn_ctx* ctx, pbn_cv* cv, const pbn_cv_item* items, int n_items, const int32_t* job_item,
                                 const int32_t* job_fold, int n_jobs, double* job_scores, int* status) {
    if (!ctx || !cv || (n_items > 0 && !items) || (n_jobs > 0 && (!job_item || !job_fold || !job_scores)))
        return set_error(PBN_ERR_ARG, "null argument");
    DevSetter ds(ctx->device);
    std::vector<int> st_local(n_items, PBN_OK);
    int* stat = status ? status : st_local.data();
    for (int i = 0; i < n_items; ++i) {
        const pbn_cv_item& it = items[i];
        stat[i] = PBN_OK;
        if (it.n_vars < 1 || it.n_vars > PBN_MAX_DIM) return set_error(PBN_ERR_UNSUPPORTED, "number of variables must be in [1, 32]");
        PBN_TRY(check_cols(cv->tbl, it.vars, it.n_vars));
        if (it.factor != PBN_FACTOR_LINEAR_GAUSSIAN && it.factor != PBN_FACTOR_CKDE) return set_error(PBN_ERR_ARG, "unknown factor type");
    }
    std::string first_error;
    std::vector<CvJob> jobs(n_jobs);
    std::vector<std::vector<int>> groups(kMaxFast + 1);
    std::vector<int> slow;
    for (int j = 0; j < n_jobs; ++j) {
        jobs[j].item = job_item[j];
        jobs[j].fold = job_fold[j];
        job_scores[j] = 0.0;
        if (jobs[j].item < 0 || jobs[j].item >= n_items) return set_error(PBN_ERR_ARG, "job refers to an item out of range");
        if (jobs[j].fold < 0 || jobs[j].fold >= cv->k) return set_error(PBN_ERR_ARG, "fold range out of bounds");
        const pbn_cv_item& it = items[jobs[j].item];
        if (it.factor == PBN_FACTOR_LINEAR_GAUSSIAN) {
            // fit on the fold's training rows and slogl on its test rows, both from the fold statistics
            const int d = it.n_vars, p = d - 1;
            std::vector<double> mean(d), Cm((size_t)d * d), beta(d);
            int64_t ntr;
            train_moments(cv, jobs[j].fold, it.vars, d, &ntr, mean.data(), Cm.data());
            if (ntr < 1) return set_error(PBN_ERR_ARG, "empty training fold");
            double var = lg_fit_from_moments(ntr, p, mean.data(), Cm.data(), beta.data());
            job_scores[j] = lg_fold_slogl(cv, jobs[j].fold, it.vars, p, beta.data(), var);
        } else if (it.n_vars <= kMaxFast) {
            groups[it.n_vars].push_back(j);
        } else {
            slow.push_back(j);
        }
    }
    // ---- batched CKDE jobs ----
    const int nd = pbn_num_devices(ctx);
    double total_cost = 0;
    auto job_cost = [&](int j) {
        const int f = jobs[j].fold;
        const double m = (double)(cv->limits[f + 1] - cv->limits[f]);
        return ((double)cv->n - m) * m * items[jobs[j].item].n_vars;
    };
    for (int d = 1; d <= kMaxFast; ++d)
        for (int j : groups[d]) total_cost += job_cost(j);
    // (cost = training rows x test rows x variables; below ~1 ms of kernel time per device one device is faster)
    if (nd > 1 && pbn_replicated(ctx, cv) && total_cost >= 1.5e9 * nd) {
        // deal: most expensive first (ties by position), round-robin; a device keeps its jobs in their original order
        std::vector<int> order;
        for (int d = 1; d <= kMaxFast; ++d) order.insert(order.end(), groups[d].begin(), groups[d].end());
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
            const double ca = job_cost(a), cb = job_cost(b);
            return ca != cb ? ca > cb : a < b;
        });
        std::vector<std::vector<std::vector<int>>> mine(nd, std::vector<std::vector<int>>(kMaxFast + 1));
        for (size_t pos = 0; pos < order.size(); ++pos) mine[pos % nd][items[jobs[order[pos]].item].n_vars].push_back(order[pos]);
        std::vector<std::vector<int>> dev_stat(nd, std::vector<int>(n_items, PBN_OK));
        std::vector<std::string> dev_err(nd);
        int rc = pbn_run_on_devices(nd, [&](int i) {
            for (int d = 1; d <= kMaxFast; ++d) {
                std::sort(mine[i][d].begin(), mine[i][d].end());
                if (mine[i][d].empty()) continue;
                PBN_TRY(score_ckde_jobs(pbn_device_ctx(ctx, i), pbn_replica(cv, i), items, jobs.data(), mine[i][d], d, job_scores,
                                        dev_stat[i].data(), &dev_err[i]));
            }
            return PBN_OK;
        });
        if (rc != PBN_OK) return rc;
        int64_t fb = 0;
        for (int i = 0; i < nd; ++i) {
            fb += pbn_device_ctx(ctx, i)->last_fallback_rows;
            for (int q = 0; q < n_items; ++q)
                if (stat[q] == PBN_OK && dev_stat[i][q] != PBN_OK) stat[q] = dev_stat[i][q];
            if (first_error.empty() && !dev_err[i].empty()) first_error = dev_err[i];
        }
        ctx->last_fallback_rows = fb;
    } else {
        for (int d = 1; d <= kMaxFast; ++d)
            if (!groups[d].empty()) PBN_TRY(score_ckde_jobs(ctx, cv, items, jobs.data(), groups[d], d, job_scores, stat, &first_error));
    }
    // wide CKDEs (d > kMaxFast = 10): one fit + slogl per fold through the single-model entry points
    for (int j : slow) {
        const int i = jobs[j].item, f = jobs[j].fold;
        if (stat[i] != PBN_OK) continue;
        const pbn_cv_item& it = items[i];
        const int d = it.n_vars;
        std::vector<double> mean(d), H((size_t)d * d);
        int64_t ntr;
        int rc = fold_bandwidth(cv, f, it.vars, d, it.rule, &ntr, mean.data(), H.data());
        pbn_kde* kde = nullptr;
        pbn_rows tr = {0, cv->limits[f], cv->limits[f + 1], cv->n};
        pbn_rows te = {cv->limits[f], cv->limits[f + 1], 0, 0};
        if (rc == PBN_OK) rc = pbn_ckde_fit(ctx, cv->tbl, it.vars, d, tr, H.data(), &kde);
        double sc = 0;
        if (rc == PBN_OK) rc = pbn_kde_logl(ctx, kde, cv->tbl, it.vars, te, nullptr, &sc);
        if (kde) pbn_kde_free(kde);
        if (rc == PBN_ERR_CUDA) return rc;
        if (rc != PBN_OK) {
            stat[i] = rc;
            if (first_error.empty()) first_error = pbn_last_error();
            continue;
        }
        job_scores[j] = sc;
    }
    if (!first_error.empty()) set_error(PBN_ERR_SINGULAR, first_error);  // message of the first failed item
    return PBN_OK;
}

extern "C" int pbn_cv_scores(pbn_ctx* ctx, pbn_cv* cv, const pbn_cv_item* items, int n_items, int fold_begin, int fold_end,
                             double* scores, int* status) {
    if (!ctx || !cv || (n_items > 0 && (!items || !scores))) return set_error(PBN_ERR_ARG, "null argument");
    if (fold_begin < 0 || fold_end > cv->k || fold_begin >= fold_end) return set_error(PBN_ERR_ARG, "fold range out of bounds");
    const int nf = fold_end - fold_begin;
    std::vector<int32_t> ji((size_t)n_items * nf), jf((size_t)n_items * nf);
    for (int i = 0; i < n_items; ++i)
        for (int q = 0; q < nf; ++q) {
            ji[(size_t)i * nf + q] = i;
            jf[(size_t)i * nf + q] = fold_begin + q;
        }
    std::vector<double> js((size_t)n_items * nf);
    PBN_TRY(pbn_cv_score_jobs(ctx, cv, items, n_items, ji.data(), jf.data(), n_items * nf, js.data(), status));
    // CVLikelihood::local_score: loglik += slogl(fold) in fold order (cv_likelihood.cpp:19-23)
    for (int i = 0; i < n_items; ++i) {
        double tot = 0;
        for (int q = 0; q < nf; ++q) tot += js[(size_t)i * nf + q];
        scores[i] = tot;
    }
    return PBN_OK;
}
