// ref_driver.cpp — runs the REFERENCE'S OWN OpenCL-C kernels (compiled as C++ through
// cl_shim.h) on the host, driven by a restatement of the reference's host enqueue logic.
// TEST INFRASTRUCTURE ONLY; built into oracle/_ref/libref_kernels.so by `make -C oracle ref`
// when /root/reference is present.  It pins oracle/pbn_oracle.cpp (the standalone
// restatement) against the arithmetic of the real kernel source.
//
// Host logic restated from (paths relative to /root/reference/pybnesian/):
//   kde/KDE.hpp:592-640 (_logl_impl: chunks of <= 64 test rows), 43-67 and 123-212
//   (execute_logl_mat, column and row variants), opencl/opencl_config.hpp:344-536
//   (reduction1d, reduction_cols_offset, logsumexp_cols_offset; local size = min(length, 256)),
//   kde/UCV.cpp:28-53, 132-178, 296-358 (score_unconstrained_impl, 64-bit chunk offsets),
//   factors/continuous/CKDE.hpp:256-287 (joint - marginal, sum1d).
#include "cl_shim.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <vector>
#include <omp.h>

thread_local WorkItem g_wi;

#include "kde_kernels_expanded.inc"  // generated under oracle/_ref/ from the reference tree

namespace {

constexpr int kMaxLocal = 256;  // emulated device work-group limit

// barrier-free kernel over a 1-D NDRange (local size chosen by the "driver": irrelevant)
template <typename F>
void run_1d(size_t global, F&& body) {
    WorkItem& w = g_wi;
    std::memset(&w, 0, sizeof(w));
    w.global_size[0] = global; w.global_size[1] = 1;
    w.local_size[0] = 1; w.num_groups[0] = global;
    for (size_t i = 0; i < global; ++i) {
        w.global_id[0] = i; w.group_id[0] = i; w.local_id[0] = 0;
        body();
    }
}

// kernel with barriers: global = groups*local (x) by cols (y), local = (local, 1)
template <typename F>
void run_groups(int groups, int local, int cols, F&& body) {
    for (int col = 0; col < cols; ++col)
        for (int g = 0; g < groups; ++g) {
#pragma omp parallel num_threads(local)
            {
                WorkItem& w = g_wi;
                std::memset(&w, 0, sizeof(w));
                int lid = omp_get_thread_num();
                w.global_size[0] = (size_t)groups * local; w.global_size[1] = cols;
                w.local_size[0] = local; w.local_size[1] = 1;
                w.num_groups[0] = groups; w.num_groups[1] = cols;
                w.group_id[0] = g; w.group_id[1] = col;
                w.local_id[0] = lid; w.local_id[1] = 0;
                w.global_id[0] = (size_t)g * local + lid; w.global_id[1] = col;
                body();
            }
        }
}

template <typename T> struct K;
template <> struct K<double> {
    static void max_mat_cols(double* m, uint r, double* l, double* o, uint off) { max_mat_cols_double(m, r, l, o, off); }
    static void sum_mat_cols(double* m, uint r, double* l, double* o, uint off) { sum_mat_cols_double(m, r, l, o, off); }
    static void sum1d(double* i, uint n, double* l, double* o, uint off) { sum1d_double(i, n, l, o, off); }
    static void lse_coeffs(double* i, uint r, double* mx) { logsumexp_coeffs_double(i, r, mx); }
    static void finish_lse(double* res, uint off, double* mx) { finish_lse_offset_double(res, off, mx); }
    static void solve(double* d, uint r, uint c, double* ch) { solve_double(d, r, c, ch); }
    static void square(double* m) { square_double(m); }
    static void substract(double* tr, uint pr, uint off, uint rows, double* te, uint tpr, uint toff, uint idx, double* res) { substract_double(tr, pr, off, rows, te, tpr, toff, idx, res); }
    static void logl_col(double* sq, uint c, double* sol, uint sr, uint idx, double ln) { logl_values_mat_column_double(sq, c, sol, sr, idx, ln); }
    static void logl_row(double* sq, uint c, double* sol, uint sr, uint idx, double ln) { logl_values_mat_row_double(sq, c, sol, sr, idx, ln); }
    static void logl_1d(double* tr, uint n, double* te, uint off, const double* sd, double ln, double* res) { logl_values_1d_mat_double(tr, n, te, off, sd, ln, res); }
    static void add_logl_1d(double* tr, uint n, double* te, uint off, const double* sd, double* res) { add_logl_values_1d_mat_double(tr, n, te, off, sd, res); }
    static void sub_vec(double* a, double* b) { substract_vectors_double(a, b); }
    static void ucv_1d(double* d, uint off, double* h, double l2, double l1, double* s2, double* s1) { sum_ucv_1d_double(d, off, h, l2, l1, s2, s1); }
    static void tri_sub(double* d, uint pr, uint c, uint off, uint rows, double* res) { triangular_substract_mat_double(d, pr, c, off, rows, res); }
    static void ucv_mat(double* sq, uint c, double l2, double l1, double* s2, double* s1) { sum_ucv_mat_double(sq, c, l2, l1, s2, s1); }
    static void exp_el(double* m) { exp_elementwise_double(m); }
    static void accum(double* m, uint r, double* l, double* sums) { accum_sum_mat_cols_double(m, r, l, sums); }
    static void add_accum(double* m, uint r, uint off, uint spg, uint ng, double* sums) { add_accum_sum_mat_cols_double(m, r, off, spg, ng, sums); }
    static void norm_accum(double* m, uint r, double* sums) { normalize_accum_sum_mat_cols_double(m, r, sums); }
    static void find_idx(double* m, uint r, uint off, double* rn, int* idx) { find_random_indices_double(m, r, off, rn, idx); }
    static void cmeans_1d(double* tr, uint pr, double* te, uint tpr, uint toff, const double* tm, double* res) { conditional_means_1d_double(tr, pr, te, tpr, toff, tm, res); }
    static void cmeans_col(double* tr, uint pr, double* sub, uint spr, double* tv, uint ec, double* res, uint ci, uint rpr) { conditional_means_column_double(tr, pr, sub, spr, tv, ec, res, ci, rpr); }
    static void cmeans_row(double* tr, uint pr, double* sub, uint spr, double* tv, uint ec, double* res, uint ri, uint rpr) { conditional_means_row_double(tr, pr, sub, spr, tv, ec, res, ri, rpr); }
    static void uni_cdf(double* means, uint pr, double* x, uint off, double inv_std, double inv_N, double* mat) { univariate_normal_cdf_double(means, pr, x, off, inv_std, inv_N, mat); }
    static void ncdf(double* means, uint pr, double* x, uint off, double inv_std) { normal_cdf_double(means, pr, x, off, inv_std); }
    static void prod_el(double* a, double* b) { product_elementwise_double(a, b); }
    static void div_el(double* a, uint off, double* b) { division_elementwise_double(a, off, b); }
};
template <> struct K<float> {
    static void max_mat_cols(float* m, uint r, float* l, float* o, uint off) { max_mat_cols_float(m, r, l, o, off); }
    static void sum_mat_cols(float* m, uint r, float* l, float* o, uint off) { sum_mat_cols_float(m, r, l, o, off); }
    static void sum1d(float* i, uint n, float* l, float* o, uint off) { sum1d_float(i, n, l, o, off); }
    static void lse_coeffs(float* i, uint r, float* mx) { logsumexp_coeffs_float(i, r, mx); }
    static void finish_lse(float* res, uint off, float* mx) { finish_lse_offset_float(res, off, mx); }
    static void solve(float* d, uint r, uint c, float* ch) { solve_float(d, r, c, ch); }
    static void square(float* m) { square_float(m); }
    static void substract(float* tr, uint pr, uint off, uint rows, float* te, uint tpr, uint toff, uint idx, float* res) { substract_float(tr, pr, off, rows, te, tpr, toff, idx, res); }
    static void logl_col(float* sq, uint c, float* sol, uint sr, uint idx, float ln) { logl_values_mat_column_float(sq, c, sol, sr, idx, ln); }
    static void logl_row(float* sq, uint c, float* sol, uint sr, uint idx, float ln) { logl_values_mat_row_float(sq, c, sol, sr, idx, ln); }
    static void logl_1d(float* tr, uint n, float* te, uint off, const float* sd, float ln, float* res) { logl_values_1d_mat_float(tr, n, te, off, sd, ln, res); }
    static void add_logl_1d(float* tr, uint n, float* te, uint off, const float* sd, float* res) { add_logl_values_1d_mat_float(tr, n, te, off, sd, res); }
    static void sub_vec(float* a, float* b) { substract_vectors_float(a, b); }
    static void ucv_1d(float* d, uint off, float* h, float l2, float l1, float* s2, float* s1) { sum_ucv_1d_float(d, off, h, l2, l1, s2, s1); }
    static void tri_sub(float* d, uint pr, uint c, uint off, uint rows, float* res) { triangular_substract_mat_float(d, pr, c, off, rows, res); }
    static void ucv_mat(float* sq, uint c, float l2, float l1, float* s2, float* s1) { sum_ucv_mat_float(sq, c, l2, l1, s2, s1); }
    static void exp_el(float* m) { exp_elementwise_float(m); }
    static void accum(float* m, uint r, float* l, float* sums) { accum_sum_mat_cols_float(m, r, l, sums); }
    static void add_accum(float* m, uint r, uint off, uint spg, uint ng, float* sums) { add_accum_sum_mat_cols_float(m, r, off, spg, ng, sums); }
    static void norm_accum(float* m, uint r, float* sums) { normalize_accum_sum_mat_cols_float(m, r, sums); }
    static void find_idx(float* m, uint r, uint off, float* rn, int* idx) { find_random_indices_float(m, r, off, rn, idx); }
    static void cmeans_1d(float* tr, uint pr, float* te, uint tpr, uint toff, const float* tm, float* res) { conditional_means_1d_float(tr, pr, te, tpr, toff, tm, res); }
    static void cmeans_col(float* tr, uint pr, float* sub, uint spr, float* tv, uint ec, float* res, uint ci, uint rpr) { conditional_means_column_float(tr, pr, sub, spr, tv, ec, res, ci, rpr); }
    static void cmeans_row(float* tr, uint pr, float* sub, uint spr, float* tv, uint ec, float* res, uint ri, uint rpr) { conditional_means_row_float(tr, pr, sub, spr, tv, ec, res, ri, rpr); }
    static void uni_cdf(float* means, uint pr, float* x, uint off, float inv_std, float inv_N, float* mat) { univariate_normal_cdf_float(means, pr, x, off, inv_std, inv_N, mat); }
    static void ncdf(float* means, uint pr, float* x, uint off, float inv_std) { normal_cdf_float(means, pr, x, off, inv_std); }
    static void prod_el(float* a, float* b) { product_elementwise_float(a, b); }
    static void div_el(float* a, uint off, float* b) { division_elementwise_float(a, off, b); }
};

// OpenCLConfig::reduction_cols_offset (opencl_config.hpp:461-515) with Max or Sum kernels.
template <typename T, bool IsMax>
void reduction_cols(std::vector<T> mat, int rows, int cols, T* out, int out_offset) {
    int length = rows;
    std::vector<T> next;
    while (true) {
        int local = std::min(length, kMaxLocal);
        int groups = (length + local - 1) / local;
        bool last = groups == 1;
        next.assign((size_t)groups * cols, T(0));
        T* target = last ? out : next.data();
        uint off = last ? (uint)out_offset : 0u;
        std::vector<std::vector<T>> locals(1);
        run_groups(groups, local, cols, [&]() {
            // one __local buffer per work-group: groups run one after another, so a static
            // buffer shared by the threads of the current group is exactly that
            static std::vector<T> lbuf;
#pragma omp single
            lbuf.assign(local, T(0));
            if (IsMax) K<T>::max_mat_cols(mat.data(), (uint)length, lbuf.data(), target, off);
            else K<T>::sum_mat_cols(mat.data(), (uint)length, lbuf.data(), target, off);
        });
        if (last) return;
        mat.swap(next);
        length = groups;
    }
}

// OpenCLConfig::sum1d / reduction1d (opencl_config.hpp:344-397)
template <typename T>
T sum1d(std::vector<T> v) {
    int length = (int)v.size();
    std::vector<T> next;
    while (true) {
        int local = std::min(length, kMaxLocal);
        int groups = (length + local - 1) / local;
        next.assign(groups, T(0));
        run_groups(groups, local, 1, [&]() {
            static std::vector<T> lbuf;
#pragma omp single
            lbuf.assign(local, T(0));
            K<T>::sum1d(v.data(), (uint)length, lbuf.data(), next.data(), 0u);
        });
        if (groups == 1) return next[0];
        v.swap(next);
        length = groups;
    }
}

// KDE::_logl_impl (KDE.hpp:592-640): result values in T
template <typename T>
void kde_logl(const T* train_c, int N, const T* test_c, int m, int d, const T* chol_c, T lognorm, T* res) {
    std::vector<T> train(train_c, train_c + (size_t)N * d), test(test_c, test_c + (size_t)m * d), chol(chol_c, chol_c + d * d);
    int allocated_m = std::min(m, 64);
    std::vector<T> mat((size_t)N * allocated_m), tmp;
    if (d > 1) tmp.resize((size_t)std::max(N, allocated_m) * d);
    int iterations = (int)std::ceil((double)m / (double)allocated_m);
    auto exec = [&](int test_offset, int test_length) {
        if (d == 1) {
            run_1d((size_t)N * test_length, [&]() { K<T>::logl_1d(train.data(), N, test.data(), test_offset, chol.data(), lognorm, mat.data()); });
        } else if (N > test_length) {
            for (int i = 0; i < test_length; ++i) {
                run_1d((size_t)N * d, [&]() { K<T>::substract(train.data(), N, 0u, N, test.data(), m, test_offset, i, tmp.data()); });
                run_1d(N, [&]() { K<T>::solve(tmp.data(), N, d, chol.data()); });
                run_1d((size_t)N * d, [&]() { K<T>::square(tmp.data()); });
                run_1d(N, [&]() { K<T>::logl_col(tmp.data(), d, mat.data(), N, i, lognorm); });
            }
        } else {
            for (int i = 0; i < N; ++i) {
                run_1d((size_t)test_length * d, [&]() { K<T>::substract(test.data(), m, test_offset, test_length, train.data(), N, 0, i, tmp.data()); });
                run_1d(test_length, [&]() { K<T>::solve(tmp.data(), test_length, d, chol.data()); });
                run_1d((size_t)test_length * d, [&]() { K<T>::square(tmp.data()); });
                run_1d(test_length, [&]() { K<T>::logl_row(tmp.data(), d, mat.data(), N, i, lognorm); });
            }
        }
        // logsumexp_cols_offset (opencl_config.hpp:517-536)
        std::vector<T> sub(mat.begin(), mat.begin() + (size_t)N * test_length), mx(test_length);
        reduction_cols<T, true>(sub, N, test_length, mx.data(), 0);
        run_1d((size_t)N * test_length, [&]() { K<T>::lse_coeffs(sub.data(), N, mx.data()); });
        reduction_cols<T, false>(sub, N, test_length, res, test_offset);
        run_1d(test_length, [&]() { K<T>::finish_lse(res, test_offset, mx.data()); });
    };
    for (int i = 0; i < iterations - 1; ++i) exec(i * allocated_m, allocated_m);
    int remaining = m - (iterations - 1) * allocated_m;
    exec(m - remaining, remaining);
}

// ProductKDE::_logl_impl + product_logl_mat (kde/ProductKDE.hpp:233-296): one training column per variable,
// test block column-major m x d, sd[c] = sqrt(h_c) in T, chunks of <= 64 test rows.
template <typename T>
void product_kde_logl(const T* train_c, int N, const T* test_c, int m, int d, const T* sd, T lognorm, T* res) {
    std::vector<T> train(train_c, train_c + (size_t)N * d), test(test_c, test_c + (size_t)m * d);
    int allocated_m = std::min(m, 64);
    std::vector<T> mat((size_t)N * allocated_m);
    int iterations = (int)std::ceil((double)m / (double)allocated_m);
    auto exec = [&](int test_offset, int test_length) {
        run_1d((size_t)N * test_length, [&]() { K<T>::logl_1d(train.data(), N, test.data(), test_offset, sd, lognorm, mat.data()); });
        for (int c = 1; c < d; ++c)
            run_1d((size_t)N * test_length, [&]() {
                K<T>::add_logl_1d(train.data() + (size_t)c * N, N, test.data(), (uint)(c * m) + test_offset, sd + c, mat.data());
            });
        std::vector<T> sub(mat.begin(), mat.begin() + (size_t)N * test_length), mx(test_length);
        reduction_cols<T, true>(sub, N, test_length, mx.data(), 0);
        run_1d((size_t)N * test_length, [&]() { K<T>::lse_coeffs(sub.data(), N, mx.data()); });
        reduction_cols<T, false>(sub, N, test_length, res, test_offset);
        run_1d(test_length, [&]() { K<T>::finish_lse(res, test_offset, mx.data()); });
    };
    for (int i = 0; i < iterations - 1; ++i) exec(i * allocated_m, allocated_m);
    int remaining = m - (iterations - 1) * allocated_m;
    exec(m - remaining, remaining);
}

template <typename T>
void ucv_sums(const T* X_c, int N, int d, const T* chol_c, T l2H, T lH, T* s2h_out, T* sh_out) {
    std::vector<T> X(X_c, X_c + (size_t)N * d), chol(chol_c, chol_c + d * d);
    uint64_t n_dist = (uint64_t)N * (N - 1) / 2;
    uint64_t per_it = std::min<uint64_t>(1000000, n_dist);
    int iterations = (int)std::ceil((double)n_dist / (double)per_it);
    std::vector<T> sum2h(per_it, T(0)), sumh(per_it, T(0)), tmp;
    if (d > 1) tmp.resize(per_it * d);
    auto exec = [&](uint64_t offset, uint64_t length) {
        if (d == 1) {
            run_1d(length, [&]() { K<T>::ucv_1d(X.data(), (uint)offset, chol.data(), l2H, lH, sum2h.data(), sumh.data()); });
        } else {
            run_1d(length * d, [&]() { K<T>::tri_sub(X.data(), N, d, (uint)offset, (uint)length, tmp.data()); });
            run_1d(length, [&]() { K<T>::solve(tmp.data(), (uint)length, d, chol.data()); });
            run_1d(length * d, [&]() { K<T>::square(tmp.data()); });
            run_1d(length, [&]() { K<T>::ucv_mat(tmp.data(), d, l2H, lH, sum2h.data(), sumh.data()); });
        }
    };
    for (int i = 0; i < iterations - 1; ++i) exec((uint64_t)i * per_it, per_it);
    uint64_t remaining = n_dist - (uint64_t)(iterations - 1) * per_it;
    exec((uint64_t)(iterations - 1) * per_it, remaining);
    *s2h_out = sum1d<T>(sum2h);
    *sh_out = sum1d<T>(sumh);
}

// UnivariateKDE / MultivariateKDE::execute_logl_mat (KDE.hpp:43-67, 123-212): N x test_length log-kernel values
template <typename T>
void logl_mat(std::vector<T>& train, int N, std::vector<T>& test, int m, int test_offset, int test_length, int d,
              std::vector<T>& chol, T lognorm, std::vector<T>& tmp, std::vector<T>& mat) {
    if (d == 1) {
        run_1d((size_t)N * test_length, [&]() { K<T>::logl_1d(train.data(), N, test.data(), test_offset, chol.data(), lognorm, mat.data()); });
    } else if (N > test_length) {
        for (int i = 0; i < test_length; ++i) {
            run_1d((size_t)N * d, [&]() { K<T>::substract(train.data(), N, 0u, N, test.data(), m, test_offset, i, tmp.data()); });
            run_1d(N, [&]() { K<T>::solve(tmp.data(), N, d, chol.data()); });
            run_1d((size_t)N * d, [&]() { K<T>::square(tmp.data()); });
            run_1d(N, [&]() { K<T>::logl_col(tmp.data(), d, mat.data(), N, i, lognorm); });
        }
    } else {
        for (int i = 0; i < N; ++i) {
            run_1d((size_t)test_length * d, [&]() { K<T>::substract(test.data(), m, test_offset, test_length, train.data(), N, 0, i, tmp.data()); });
            run_1d(test_length, [&]() { K<T>::solve(tmp.data(), test_length, d, chol.data()); });
            run_1d((size_t)test_length * d, [&]() { K<T>::square(tmp.data()); });
            run_1d(test_length, [&]() { K<T>::logl_row(tmp.data(), d, mat.data(), N, i, lognorm); });
        }
    }
}

// kernel over a 2-D NDRange without barriers
template <typename F>
void run_2d(size_t g0, size_t g1, F&& body) {
    WorkItem& w = g_wi;
    std::memset(&w, 0, sizeof(w));
    w.global_size[0] = g0; w.global_size[1] = g1;
    w.local_size[0] = 1; w.local_size[1] = 1; w.num_groups[0] = g0; w.num_groups[1] = g1;
    for (size_t j = 0; j < g1; ++j)
        for (size_t i = 0; i < g0; ++i) {
            w.global_id[0] = i; w.group_id[0] = i; w.global_id[1] = j; w.group_id[1] = j;
            body();
        }
}

inline int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

// OpenCLConfig::accum_sum_cols (opencl_config.hpp:539-579): in-place exclusive prefix sum of every column,
// returns the per-column totals.  Work-group size = the emulated device limit.
template <typename T>
std::vector<T> accum_sum_cols(std::vector<T>& mat, int rows, int cols) {
    int local_wg = rows > kMaxLocal ? kMaxLocal : next_pow2(rows);
    int num_groups = (int)std::ceil((double)rows / (double)(2 * local_wg));
    std::vector<T> group_sums((size_t)num_groups * cols, T(0));
    run_groups(num_groups, local_wg, cols, [&]() {
        static std::vector<T> lbuf;
#pragma omp single
        lbuf.assign(2 * local_wg, T(0));
        K<T>::accum(mat.data(), (uint)rows, lbuf.data(), group_sums.data());
    });
    if (num_groups > 1) {
        std::vector<T> total = accum_sum_cols<T>(group_sums, num_groups, cols);
        run_2d((size_t)(rows - 2 * local_wg), cols, [&]() {
            K<T>::add_accum(mat.data(), (uint)rows, (uint)(2 * local_wg), (uint)(2 * local_wg), (uint)num_groups, group_sums.data());
        });
        return total;
    }
    return group_sums;
}

// CKDE::_sample_indices_from_weights (factors/continuous/CKDE.hpp:402-504), chunks of <= 64 evidence rows
template <typename T>
void sample_indices(const T* mtrain_c, int N, const T* etest_c, int n, int p, const T* chol_c, T lognorm, const T* rp_c, int* res) {
    std::vector<T> train(mtrain_c, mtrain_c + (size_t)N * p), test(etest_c, etest_c + (size_t)n * p), chol(chol_c, chol_c + p * p);
    std::vector<T> rp(rp_c, rp_c + n);
    for (int i = 0; i < n; ++i) res[i] = N - 1;
    int allocated_m = std::min(n, 64);
    std::vector<T> mat((size_t)N * allocated_m), tmp;
    if (p > 1) tmp.resize((size_t)std::max(N, allocated_m) * p);
    int iterations = (int)std::ceil((double)n / (double)allocated_m);
    auto exec = [&](int off, int len) {
        logl_mat<T>(train, N, test, n, off, len, p, chol, lognorm, tmp, mat);
        run_1d((size_t)N * len, [&]() { K<T>::exp_el(mat.data()); });
        std::vector<T> total = accum_sum_cols<T>(mat, N, len);
        run_2d((size_t)(N - 1), len, [&]() { K<T>::norm_accum(mat.data(), (uint)N, total.data()); });
        run_2d((size_t)(N - 1), len, [&]() { K<T>::find_idx(mat.data(), (uint)N, (uint)off, rp.data(), res); });
    };
    for (int i = 0; i < iterations - 1; ++i) exec(i * allocated_m, allocated_m);
    int remaining = n - (iterations - 1) * allocated_m;
    exec(n - remaining, remaining);
}

// CKDE::_cdf_univariate / _cdf_multivariate (factors/continuous/CKDE.hpp:558-728), chunks of <= 64 test rows.
// train: N x d column-major (variable first); xtest: m values of the variable; etest: m x p evidence block.
template <typename T>
void ckde_cdf(const T* train_c, int N, const T* xtest_c, const T* etest_c, int m, int d, const T* chol_marg_c, T lognorm_w,
              const T* transform_c, T inv_std, T* res) {
    const int p = d - 1;
    std::vector<T> train(train_c, train_c + (size_t)N * d), x(xtest_c, xtest_c + m);
    int allocated_m = std::min(m, 64);
    int iterations = (int)std::ceil((double)m / (double)allocated_m);
    std::vector<T> mu((size_t)N * allocated_m);
    if (p == 0) {
        T inv_N = (T)(1.0 / N);
        auto exec = [&](int off, int len) {
            run_1d((size_t)N * len, [&]() { K<T>::uni_cdf(train.data(), (uint)N, x.data(), (uint)off, inv_std, inv_N, mu.data()); });
            std::vector<T> sub(mu.begin(), mu.begin() + (size_t)N * len);
            reduction_cols<T, false>(sub, N, len, res, off);
        };
        for (int i = 0; i < iterations - 1; ++i) exec(i * allocated_m, allocated_m);
        int remaining = m - (iterations - 1) * allocated_m;
        exec(m - remaining, remaining);
        return;
    }
    std::vector<T> mtrain(train.begin() + N, train.end()), etest(etest_c, etest_c + (size_t)m * p);
    std::vector<T> chol(chol_marg_c, chol_marg_c + p * p), transform(transform_c, transform_c + p);
    std::vector<T> W((size_t)N * allocated_m), sum_W(allocated_m), tmp;
    if (p > 1) tmp.resize((size_t)std::max(N, allocated_m) * p);
    auto exec = [&](int off, int len) {
        logl_mat<T>(mtrain, N, etest, m, off, len, p, chol, lognorm_w, tmp, W);
        run_1d((size_t)N * len, [&]() { K<T>::exp_el(W.data()); });
        {
            std::vector<T> sub(W.begin(), W.begin() + (size_t)N * len);
            reduction_cols<T, false>(sub, N, len, sum_W.data(), 0);
        }
        if (p == 1) {
            run_1d((size_t)N * len, [&]() { K<T>::cmeans_1d(train.data(), (uint)N, etest.data(), (uint)m, (uint)off, transform.data(), mu.data()); });
        } else if (N > len) {
            for (int i = 0; i < len; ++i) {
                run_1d((size_t)N * p, [&]() { K<T>::substract(mtrain.data(), N, 0u, N, etest.data(), m, off, i, tmp.data()); });
                run_1d(N, [&]() { K<T>::cmeans_col(train.data(), (uint)N, tmp.data(), (uint)N, transform.data(), (uint)p, mu.data(), (uint)i, (uint)N); });
            }
        } else {
            for (int i = 0; i < N; ++i) {
                run_1d((size_t)len * p, [&]() { K<T>::substract(etest.data(), m, off, len, mtrain.data(), N, 0, i, tmp.data()); });
                run_1d(len, [&]() { K<T>::cmeans_row(train.data(), (uint)N, tmp.data(), (uint)len, transform.data(), (uint)p, mu.data(), (uint)i, (uint)N); });
            }
        }
        run_1d((size_t)N * len, [&]() { K<T>::ncdf(mu.data(), (uint)N, x.data(), (uint)off, inv_std); });
        run_1d((size_t)N * len, [&]() { K<T>::prod_el(mu.data(), W.data()); });
        {
            std::vector<T> sub(mu.begin(), mu.begin() + (size_t)N * len);
            reduction_cols<T, false>(sub, N, len, res, off);
        }
        run_1d(len, [&]() { K<T>::div_el(res, (uint)off, sum_W.data()); });
    };
    for (int i = 0; i < iterations - 1; ++i) exec(i * allocated_m, allocated_m);
    int remaining = m - (iterations - 1) * allocated_m;
    exec(m - remaining, remaining);
}

}  // namespace

extern "C" {

// train/test column-major in the data type; chol = lower Cholesky factor in the data type
// (column-major d x d); lognorm already rounded to the data type.  out: m values widened to double.
int ref_kde_logl(const void* train, int N, const void* test, int m, int d, int dtype, const void* chol, double lognorm,
                 double* out, double* out_sum) {
    if (dtype == 0) {
        std::vector<double> res(m);
        kde_logl<double>((const double*)train, N, (const double*)test, m, d, (const double*)chol, lognorm, res.data());
        for (int i = 0; i < m; ++i) out[i] = res[i];
        if (out_sum) *out_sum = sum1d<double>(res);
    } else {
        std::vector<float> res(m);
        kde_logl<float>((const float*)train, N, (const float*)test, m, d, (const float*)chol, (float)lognorm, res.data());
        for (int i = 0; i < m; ++i) out[i] = res[i];
        if (out_sum) *out_sum = sum1d<float>(res);
    }
    return 0;
}

// ProductKDE logl/slogl: sd = sqrt(h) in the data type (ProductKDE.hpp:172-179), lognorm rounded to the data type.
int ref_product_kde_logl(const void* train, int N, const void* test, int m, int d, int dtype, const void* sd,
                         double lognorm, double* out, double* out_sum) {
    if (dtype == 0) {
        std::vector<double> res(m);
        product_kde_logl<double>((const double*)train, N, (const double*)test, m, d, (const double*)sd, lognorm, res.data());
        for (int i = 0; i < m; ++i) out[i] = res[i];
        if (out_sum) *out_sum = sum1d<double>(res);
    } else {
        std::vector<float> res(m);
        product_kde_logl<float>((const float*)train, N, (const float*)test, m, d, (const float*)sd, (float)lognorm, res.data());
        for (int i = 0; i < m; ++i) out[i] = res[i];
        if (out_sum) *out_sum = sum1d<float>(res);
    }
    return 0;
}

// CKDE::_slogl: joint - marginal (`substract_vectors`), then sum1d.  Inputs: the two logl vectors in T.
int ref_ckde_combine(const double* joint, const double* marg, int m, int dtype, double* out, double* out_sum) {
    if (dtype == 0) {
        std::vector<double> a(joint, joint + m), b(marg, marg + m);
        run_1d(m, [&]() { K<double>::sub_vec(a.data(), b.data()); });
        for (int i = 0; i < m; ++i) out[i] = a[i];
        if (out_sum) *out_sum = sum1d<double>(a);
    } else {
        std::vector<float> a(m), b(m);
        for (int i = 0; i < m; ++i) { a[i] = (float)joint[i]; b[i] = (float)marg[i]; }
        run_1d(m, [&]() { K<float>::sub_vec(a.data(), b.data()); });
        for (int i = 0; i < m; ++i) out[i] = a[i];
        if (out_sum) *out_sum = sum1d<float>(a);
    }
    return 0;
}

// The two pair sums of UCVScorer::score_unconstrained_impl (UCV.cpp:296-358) in the data type.
int ref_ucv_sums(const void* X, int N, int d, int dtype, const void* chol, double lognorm_2H, double lognorm_H,
                 double* s2h, double* sh) {
    if (dtype == 0) {
        double a, b;
        ucv_sums<double>((const double*)X, N, d, (const double*)chol, lognorm_2H, lognorm_H, &a, &b);
        *s2h = a; *sh = b;
    } else {
        float a, b;
        ucv_sums<float>((const float*)X, N, d, (const float*)chol, (float)lognorm_2H, (float)lognorm_H, &a, &b);
        *s2h = a; *sh = b;
    }
    return 0;
}


// CKDE::cdf through the reference kernels.  chol_marg (p x p), transform (p) and the scalars are the host-side
// quantities of CKDE.hpp:594-616 rounded to the data type by the caller; lognorm_w = lognorm_marg + log N.
int ref_ckde_cdf(const void* train, int N, const void* xtest, const void* etest, int m, int d, int dtype,
                 const void* chol_marg, double lognorm_w, const void* transform, double inv_std, double* out) {
    if (dtype == 0) {
        std::vector<double> res(m);
        ckde_cdf<double>((const double*)train, N, (const double*)xtest, (const double*)etest, m, d, (const double*)chol_marg,
                         lognorm_w, (const double*)transform, inv_std, res.data());
        for (int i = 0; i < m; ++i) out[i] = res[i];
    } else {
        std::vector<float> res(m);
        ckde_cdf<float>((const float*)train, N, (const float*)xtest, (const float*)etest, m, d, (const float*)chol_marg,
                        (float)lognorm_w, (const float*)transform, (float)inv_std, res.data());
        for (int i = 0; i < m; ++i) out[i] = res[i];
    }
    return 0;
}

// CKDE::_sample_indices_from_weights through the reference kernels (exp_elementwise, accum_sum_mat_cols,
// add_accum_sum_mat_cols, normalize_accum_sum_mat_cols, find_random_indices).
int ref_ckde_sample_indices(const void* mtrain, int N, const void* etest, int n, int p, int dtype, const void* chol_marg,
                            double lognorm_marg, const void* random_prob, int* out) {
    if (dtype == 0)
        sample_indices<double>((const double*)mtrain, N, (const double*)etest, n, p, (const double*)chol_marg, lognorm_marg,
                               (const double*)random_prob, out);
    else
        sample_indices<float>((const float*)mtrain, N, (const float*)etest, n, p, (const float*)chol_marg,
                              (float)lognorm_marg, (const float*)random_prob, out);
    return 0;
}

}  // extern "C"
