// pbn_oracle.cpp — CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
//
// A plain C++17 restatement (no Eigen, no OpenCL, no CUDA) of the arithmetic the
// reference (davenza/PyBNesian v0.5.1) performs on its KDE / CKDE log-likelihood
// hot path.  Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
// `--impl reference` legs of `bench.py` may load this library; the product
// (`pybnesian_b200/`) never links, imports or calls it.
//
// Parity pin: the reference ships no golden vectors for this path;
// its own tests compare against SciPy.  This oracle is pinned (tests/test_oracle*.py)
//   (1) against scipy.stats.gaussian_kde exactly the way the reference's tests do
//       (tests/factors/continuous/KDE_test.py:167-203, CKDE_test.py:146-179), and
//   (2) against `oracle/_ref/libref_kernels.so`: the reference's own OpenCL-C kernel
//       source compiled as C through a work-item shim (oracle/ref_shim/), and
//   (3) against a long-double direct evaluation (`orc_kde_logl_ld`).
// The optimiser used by UCV (NLopt Nelder-Mead, un-vendored) is NOT restated here:
// "parity unpinned" for the optimiser trajectory; the UCV objective is pinned.
//
// Every function cites the reference file:line it follows (paths relative to
// /root/reference/pybnesian/).  Matrices are column-major, contiguous, null rows
// already removed by the caller (what DataFrame::to_eigen does, dataset.hpp:236-338).
//
// Build: see oracle/Makefile  (g++ -O2 -ffp-contract=off -fopenmp).

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <numeric>
#include <random>
#include <unordered_set>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

constexpr double kPi = 3.141592653589793238462643383279502884;
// util/math_constants.hpp:30  machine_tol = sqrt(eps<double>)
const double kMachineTol = std::sqrt(std::numeric_limits<double>::epsilon());

// Work-group size used for the emulated OpenCL tree reductions
// (opencl/opencl_config.hpp:344-397: local = min(len, device max work-group)).
// The reference value is device dependent; 256 is a common device maximum.
constexpr int kLocalSize = 256;

// ---- tree reductions ------------------------------------------------------------
// One work-group of kde/opencl_kernels/KDE.cl.src:22-65 (`sum1d` / `max1d`):
// halving strides, the odd element is kept at slot 0.
template <typename T, bool IsMax>
T group_reduce(T* a, int gs) {
    while (gs > 1) {
        int stride = gs / 2;
        if (gs % 2 == 0) {
            for (int id = 0; id < stride; ++id) {
                if (IsMax) a[id] = std::max(a[id], a[id + stride]);
                else a[id] = a[id] + a[id + stride];
            }
            gs = gs / 2;
        } else {
            for (int id = 0; id < stride; ++id) {
                if (IsMax) a[id + 1] = std::max(a[id + 1], a[id + 1 + stride]);
                else a[id + 1] = a[id + 1] + a[id + 1 + stride];
            }
            gs = gs / 2 + 1;
        }
    }
    return a[0];
}

// Multi-level reduction, opencl/opencl_config.hpp:344-397 (+ update_reduction_status,
// opencl_config.cpp:278-283).  `v` is consumed.
template <typename T, bool IsMax>
T tree_reduce(std::vector<T>& v) {
    int length = static_cast<int>(v.size());
    if (length == 0) return IsMax ? -std::numeric_limits<T>::infinity() : T(0);
    std::vector<T> next;
    while (true) {
        int local = std::min(length, kLocalSize);
        int groups = (length + local - 1) / local;
        next.resize(groups);
        for (int g = 0; g < groups; ++g) {
            int gs = (g == groups - 1) ? (length - g * local) : local;
            next[g] = group_reduce<T, IsMax>(v.data() + static_cast<size_t>(g) * local, gs);
        }
        if (groups == 1) return next[0];
        v.swap(next);
        length = groups;
    }
}

// ---- small dense helpers ----------------------------------------------------------
// Lower Cholesky factor (what Eigen `llt().matrixLLT()` holds in its lower triangle,
// kde/KDE.hpp:459-460).  Column-major d x d.  Returns false if not PD.
bool cholesky_lower(const double* H, int d, double* L) {
    std::fill(L, L + d * d, 0.0);
    for (int j = 0; j < d; ++j) {
        double s = H[j + j * d];
        for (int k = 0; k < j; ++k) s -= L[j + k * d] * L[j + k * d];
        if (!(s > 0.0)) return false;
        double ljj = std::sqrt(s);
        L[j + j * d] = ljj;
        for (int i = j + 1; i < d; ++i) {
            double t = H[i + j * d];
            for (int k = 0; k < j; ++k) t -= L[i + k * d] * L[j + k * d];
            L[i + j * d] = t / ljj;
        }
    }
    return true;
}

// Eigenvalues of a symmetric matrix by cyclic Jacobi (stands in for
// Eigen::SelfAdjointEigenSolver in util/basic_eigen_ops.hpp:136-147).
void sym_eigenvalues(std::vector<double> a, int d, std::vector<double>& ev) {
    for (int sweep = 0; sweep < 100; ++sweep) {
        double off = 0;
        for (int p = 0; p < d; ++p)
            for (int q = p + 1; q < d; ++q) off += a[p + q * d] * a[p + q * d];
        if (off < 1e-300) break;
        for (int p = 0; p < d; ++p)
            for (int q = p + 1; q < d; ++q) {
                double apq = a[p + q * d];
                if (apq == 0.0) continue;
                double app = a[p + p * d], aqq = a[q + q * d];
                double theta = (aqq - app) / (2.0 * apq);
                double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < d; ++k) {
                    double akp = a[k + p * d], akq = a[k + q * d];
                    a[k + p * d] = c * akp - s * akq;
                    a[k + q * d] = s * akp + c * akq;
                }
                for (int k = 0; k < d; ++k) {
                    double apk = a[p + k * d], aqk = a[q + k * d];
                    a[p + k * d] = c * apk - s * aqk;
                    a[q + k * d] = s * apk + c * aqk;
                }
            }
    }
    ev.resize(d);
    for (int i = 0; i < d; ++i) ev[i] = a[i + i * d];
}

// util/basic_eigen_ops.hpp:136-147
template <typename T>
bool is_psd(const std::vector<T>& cov, int d) {
    std::vector<double> a(cov.begin(), cov.end()), ev;
    sym_eigenvalues(a, d, ev);
    double mx = *std::max_element(ev.begin(), ev.end());
    double mn = *std::min_element(ev.begin(), ev.end());
    double tol = mx * d * static_cast<double>(std::numeric_limits<T>::epsilon());
    return !(mn < tol);
}

// Sum in T with 8 interleaved partial sums (the order a packet-vectorised
// Eigen `.sum()` / `.mean()` / `.dot()` produces up to the packet width; the exact
// width depends on the reference's build flags, so only rounding-level agreement
// with the reference is claimed here).
template <typename T, typename F>
T lane_sum(int64_t n, F&& term) {
    T acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int64_t i = 0;
    for (; i + 8 <= n; i += 8)
        for (int l = 0; l < 8; ++l) acc[l] += term(i + l);
    T s = ((acc[0] + acc[4]) + (acc[2] + acc[6])) + ((acc[1] + acc[5]) + (acc[3] + acc[7]));
    for (; i < n; ++i) s += term(i);
    return s;
}

// dataset/dataset.hpp:341-396  two-pass covariance in T, 1/(N-1) in T.
template <typename T>
void cov_T(const T* X, int64_t n, int d, std::vector<T>& cov) {
    std::vector<std::vector<T>> c(d, std::vector<T>(n));
    for (int j = 0; j < d; ++j) {
        const T* x = X + static_cast<size_t>(j) * n;
        T mean = lane_sum<T>(n, [&](int64_t i) { return x[i]; }) / static_cast<T>(n);
        for (int64_t i = 0; i < n; ++i) c[j][i] = x[i] - mean;
    }
    T inv_N = 1 / static_cast<T>(n - 1);
    cov.assign(static_cast<size_t>(d) * d, T(0));
    for (int i = 0; i < d; ++i) {
        cov[i + i * d] = lane_sum<T>(n, [&](int64_t r) { return c[i][r] * c[i][r]; }) * inv_N;
        for (int j = i + 1; j < d; ++j) {
            T v = lane_sum<T>(n, [&](int64_t r) { return c[i][r] * c[j][r]; }) * inv_N;
            cov[i + j * d] = cov[j + i * d] = v;
        }
    }
}

// kde/NormalReferenceRule.hpp:109-134 and kde/ScottsBandwidth.hpp:91-117.
// status: 0 ok, 1 = valid_rows <= d, 2 = covariance not positive definite.
template <typename T>
int bandwidth_T(const T* X, int64_t n, int d, int rule, double* H) {
    if (n <= d) return 1;
    std::vector<T> cov;
    cov_T<T>(X, n, d, cov);
    if (!is_psd<T>(cov, d)) return 2;
    T N = static_cast<T>(n), dd = static_cast<T>(d);
    double k;
    if (rule == 0)
        k = std::pow(4. / (N * (dd + 2.)), 2. / (dd + 4));
    else
        k = std::pow(static_cast<double>(N), -2. / (dd + 4));
    for (int i = 0; i < d * d; ++i) H[i] = k * static_cast<double>(cov[i]);
    return 0;
}

// kde/KDE.hpp:476-477
double lognorm_const(const double* L, int d, int64_t N) {
    double s = 0;
    for (int i = 0; i < d; ++i) s += std::log(L[i + i * d]);
    return -s - 0.5 * d * std::log(2 * kPi) - std::log(static_cast<double>(N));
}

template <typename T> inline T exp_T(T x);
template <> inline double exp_T<double>(double x) { return std::exp(x); }
template <> inline float exp_T<float>(float x) { return expf(x); }
template <typename T> inline T log_T(T x);
template <> inline double log_T<double>(double x) { return std::log(x); }
template <> inline float log_T<float>(float x) { return logf(x); }

// One (train i, test t) log-kernel value in T.
//  d == 1: kde/opencl_kernels/KDE.cl.src:143-156 (`logl_values_1d_mat`)
//  d >= 2: `substract` (173-187) -> `solve` (123-135) -> `square` (137-141, through a
//          double temporary) -> `logl_values_mat_column/_row` (190-226).
// The `_row` variant (KDE.hpp:180-211) subtracts in the opposite order; forward
// substitution is odd and the result is squared, so both give identical values.
template <typename T>
inline T pair_logl(const T* train, int64_t N, int64_t i, const T* test, int64_t m, int64_t t, int d,
                   const T* Lt, T lognorm, T* delta) {
    if (d == 1) {
        T u = (train[i] - test[t]) / Lt[0];
        return static_cast<T>((-0.5 * static_cast<double>(u)) * static_cast<double>(u) +
                              static_cast<double>(lognorm));
    }
    for (int c = 0; c < d; ++c) delta[c] = test[t + static_cast<size_t>(c) * m] - train[i + static_cast<size_t>(c) * N];
    for (int c = 0; c < d; ++c) {
        for (int k = 0; k < c; ++k) delta[c] -= Lt[c + k * d] * delta[k];
        delta[c] /= Lt[c + c * d];
    }
    T summation = 0;
    for (int c = 0; c < d; ++c) {
        double dd = delta[c];
        T sq = static_cast<T>(dd * dd);
        if (c == 0) summation = sq; else summation += sq;
    }
    return static_cast<T>((-0.5 * static_cast<double>(summation)) + static_cast<double>(lognorm));
}

// KDE::_logl_impl (kde/KDE.hpp:592-640) + OpenCLConfig::logsumexp_cols_offset
// (opencl/opencl_config.hpp:517-536; kernels KDE.cl.src:67-121, 228-233):
// per test row: N log-kernel values, max, exp(x - max), tree-sum, log + max.
// Output values are in T (returned widened to double as KDE::_logl does, 523-526).
template <typename T>
void kde_logl_T(const T* train, int64_t N, const T* test, int64_t m, int d, const double* H, T* out) {
    std::vector<double> L(d * d);
    cholesky_lower(H, d, L.data());
    std::vector<T> Lt(d * d);
    for (int i = 0; i < d * d; ++i) Lt[i] = static_cast<T>(L[i]);  // KDE.hpp:464-470
    T lognorm = static_cast<T>(lognorm_const(L.data(), d, N));
#pragma omp parallel
    {
        std::vector<T> col(N), work(N), delta(d > 0 ? d : 1);
#pragma omp for schedule(dynamic, 8)
        for (int64_t t = 0; t < m; ++t) {
            for (int64_t i = 0; i < N; ++i)
                col[i] = pair_logl<T>(train, N, i, test, m, t, d, Lt.data(), lognorm, delta.data());
            work = col;
            T mx = tree_reduce<T, true>(work);
            for (int64_t i = 0; i < N; ++i) col[i] = exp_T<T>(col[i] - mx);
            T s = tree_reduce<T, false>(col);
            col.resize(N);
            out[t] = log_T<T>(s) + mx;
        }
    }
}

template <typename T>
double slogl_T(const T* logl, int64_t m) {
    // KDE::_slogl (kde/KDE.hpp:549-562): `sum1d` tree in T, then widened.
    std::vector<T> v(logl, logl + m);
    return static_cast<double>(tree_reduce<T, false>(v));
}

// Pair index enumeration of the UCV kernels (KDE.cl.src:479-482 and siblings).
inline void ucv_pair(uint64_t k, unsigned& i1, unsigned& i2) {
    double ii = static_cast<double>(k) + 1;
    i1 = static_cast<unsigned>(std::ceil(std::sqrt(2.0 * ii + 0.25) - 0.5));
    i2 = static_cast<unsigned>(ii - (static_cast<double>(i1) - 1) * i1 * 0.5 - 1);
}

// UCVScorer::score_unconstrained_impl (kde/UCV.cpp:296-358) with the kernels
// `sum_ucv_1d` (KDE.cl.src:471-489) / `triangular_substract_mat`+`solve`+`square`+
// `sum_ucv_mat` (492-526).  Pairs are processed in chunks of 1e6 accumulating into
// 1e6 slots of T, then `sum1d`.  The reference passes the chunk offset as a 32-bit
// unsigned (UCV.cpp:32,136) and wraps for N > 92 682; this oracle uses the intended
// 64-bit offset (SURVEY.md §7 "Reference bugs").
template <typename T>
double ucv_unconstrained_T(const T* X, int64_t N, int d, const double* H) {
    std::vector<T> Ht(d * d);
    for (int i = 0; i < d * d; ++i) Ht[i] = static_cast<T>(H[i]);  // UCV.cpp:388-392 cast<float>()
    std::vector<double> Hd(Ht.begin(), Ht.end()), Ld(d * d);
    // copy_unconstrained_bandwidth (UCV.cpp:221-233): llt() of the T matrix.  A T-precision
    // Cholesky is emulated by factorising in double and rounding the factor to T.
    cholesky_lower(Hd.data(), d, Ld.data());
    std::vector<T> Lt(d * d);
    for (int i = 0; i < d * d; ++i) Lt[i] = static_cast<T>(Ld[i]);
    double slog = 0;
    for (int i = 0; i < d; ++i) slog += std::log(static_cast<double>(Lt[i + i * d]));
    T lognorm_H = static_cast<T>(-slog - 0.5 * d * std::log(2 * static_cast<double>(static_cast<T>(kPi))));
    double lognorm_2H = static_cast<double>(lognorm_H) - 0.5 * d * std::log(2.);
    T l2H = static_cast<T>(lognorm_2H), lH = lognorm_H;

    uint64_t n_dist = static_cast<uint64_t>(N) * (N - 1) / 2;
    uint64_t per_it = std::min<uint64_t>(1000000, n_dist);
    std::vector<T> sum2h(per_it, T(0)), sumh(per_it, T(0));
#pragma omp parallel
    {
        std::vector<T> delta(d);
#pragma omp for schedule(static)
        for (int64_t slot = 0; slot < static_cast<int64_t>(per_it); ++slot) {
            for (uint64_t k = slot; k < n_dist; k += per_it) {
                unsigned i1, i2;
                ucv_pair(k, i1, i2);
                T s;
                if (d == 1) {
                    T u = (X[i1] - X[i2]) / Lt[0];
                    s = u * u;
                } else {
                    for (int c = 0; c < d; ++c)
                        delta[c] = X[i1 + static_cast<size_t>(c) * N] - X[i2 + static_cast<size_t>(c) * N];
                    for (int c = 0; c < d; ++c) {
                        for (int kk = 0; kk < c; ++kk) delta[c] -= Lt[c + kk * d] * delta[kk];
                        delta[c] /= Lt[c + c * d];
                    }
                    s = 0;
                    for (int c = 0; c < d; ++c) {
                        double dd = delta[c];
                        T sq = static_cast<T>(dd * dd);
                        if (c == 0) s = sq; else s += sq;
                    }
                }
                // `exp(-0.25*s + lognorm)`: double literal => evaluated in double, stored to T.
                sum2h[slot] = static_cast<T>(static_cast<double>(sum2h[slot]) +
                                             std::exp(-0.25 * static_cast<double>(s) + static_cast<double>(l2H)));
                sumh[slot] = static_cast<T>(static_cast<double>(sumh[slot]) +
                                            std::exp(-0.5 * static_cast<double>(s) + static_cast<double>(lH)));
            }
        }
    }
    T s2h = tree_reduce<T, false>(sum2h);
    T sh = tree_reduce<T, false>(sumh);
    // UCV.cpp:357  (N is size_t -> converted to T in the mixed expressions)
    return std::exp(lognorm_2H) + static_cast<double>(2 * s2h / static_cast<T>(N)) -
           static_cast<double>(4 * sh / static_cast<T>(N - 1));
}

// UCVScorer::score_diagonal_impl (kde/UCV.cpp:235-294), kernels `ucv_diag`,
// `sum_ucv_diag`, `copy_ucv_diag` (KDE.cl.src:529-574).  `hdiag` = diagonal of H.
template <typename T>
double ucv_diagonal_T(const T* X, int64_t N, int d, const double* hdiag) {
    std::vector<T> h(d);
    for (int i = 0; i < d; ++i) h[i] = std::sqrt(static_cast<T>(hdiag[i]));  // UCV.cpp:366-369
    double slog = 0;
    for (int i = 0; i < d; ++i) slog += std::log(static_cast<double>(h[i]));
    T lognorm_H = static_cast<T>(-slog - 0.5 * d * std::log(2 * static_cast<double>(static_cast<T>(kPi))));
    double lognorm_2H = static_cast<double>(lognorm_H) - 0.5 * d * std::log(2.);
    T l2H = static_cast<T>(lognorm_2H), lH = lognorm_H;
    uint64_t n_dist = static_cast<uint64_t>(N) * (N - 1) / 2;
    uint64_t per_it = std::min<uint64_t>(1000000, n_dist);
    std::vector<T> sum2h(per_it, T(0)), sumh(per_it, T(0));
#pragma omp parallel for schedule(static)
    for (int64_t slot = 0; slot < static_cast<int64_t>(per_it); ++slot) {
        for (uint64_t k = slot; k < n_dist; k += per_it) {
            unsigned i1, i2;
            ucv_pair(k, i1, i2);
            T s = 0;
            for (int c = 0; c < d; ++c) {
                T u = (X[i1 + static_cast<size_t>(c) * N] - X[i2 + static_cast<size_t>(c) * N]) / h[c];
                u = u * u;
                if (c == 0) s = u; else s += u;
            }
            sum2h[slot] = static_cast<T>(static_cast<double>(sum2h[slot]) +
                                         std::exp(-0.25 * static_cast<double>(s) + static_cast<double>(l2H)));
            sumh[slot] = static_cast<T>(static_cast<double>(sumh[slot]) +
                                        std::exp(-0.5 * static_cast<double>(s) + static_cast<double>(lH)));
        }
    }
    T s2h = tree_reduce<T, false>(sum2h);
    T sh = tree_reduce<T, false>(sumh);
    return std::exp(lognorm_2H) + static_cast<double>(2 * s2h / static_cast<T>(N)) -
           static_cast<double>(4 * sh / static_cast<T>(N - 1));
}

// ---- ProductKDE -----------------------------------------------------
// NormalReferenceRule::diag_bandwidth (kde/NormalReferenceRule.hpp:72-106, eq. 3.4 of Chacon & Duong 2018)
// and ScottsBandwidth::diag_bandwidth (kde/ScottsBandwidth.hpp:66-89), in T like the reference.
// `inverse()` / `determinant()` of the d x d matrix are Eigen PartialPivLU in the reference; Gauss-Jordan
// with partial pivoting in T stands in (rounding-level agreement only).
// status: 0 ok, 1 = valid_rows <= d (<= 1 for Scott), 2 = covariance not positive definite.
template <typename T>
int diag_bandwidth_T(const T* X, int64_t n, int d, int rule, double* h) {
    if (n <= (rule == 0 ? d : 1)) return 1;
    std::vector<T> cov;
    cov_T<T>(X, n, d, cov);
    T N = static_cast<T>(n), dd = static_cast<T>(d);
    if (rule != 0) {
        double k = std::pow(static_cast<double>(N), -2. / (dd + 4.));
        for (int i = 0; i < d; ++i) h[i] = k * static_cast<double>(cov[i + i * d]);
        return 0;
    }
    if (!is_psd<T>(cov, d)) return 2;
    std::vector<T> A(d * d), Inv(d * d, T(0));
    for (int i = 0; i < d; ++i)
        for (int j = 0; j < d; ++j) A[i + j * d] = cov[i + j * d] * (T(1) / cov[i + i * d]);
    for (int i = 0; i < d; ++i) Inv[i + i * d] = 1;
    T det = 1;
    for (int c = 0; c < d; ++c) {
        int piv = c;
        for (int r = c + 1; r < d; ++r)
            if (std::fabs(A[r + c * d]) > std::fabs(A[piv + c * d])) piv = r;
        if (piv != c) {
            for (int j = 0; j < d; ++j) {
                std::swap(A[c + j * d], A[piv + j * d]);
                std::swap(Inv[c + j * d], Inv[piv + j * d]);
            }
            det = -det;
        }
        T pv = A[c + c * d];
        det *= pv;
        for (int j = 0; j < d; ++j) { A[c + j * d] /= pv; Inv[c + j * d] /= pv; }
        for (int r = 0; r < d; ++r) {
            if (r == c) continue;
            T f = A[r + c * d];
            for (int j = 0; j < d; ++j) { A[r + j * d] -= f * A[c + j * d]; Inv[r + j * d] -= f * Inv[c + j * d]; }
        }
    }
    T tr = 0, tr2 = 0;
    for (int i = 0; i < d; ++i) {
        tr += Inv[i + i * d];
        for (int j = 0; j < d; ++j) tr2 += Inv[i + j * d] * Inv[j + i * d];
    }
    T k = 4 * dd * std::sqrt(det) / (2 * tr2 + tr * tr);
    // `std::pow(k / N, 2. / (d + 4.)) * diag`: pow in double; Eigen promotes the scalar to the vector's
    // scalar type (float for float data), the product is evaluated in T, then cast to double.
    T f = static_cast<T>(std::pow(k / N, 2. / (dd + 4.)));
    for (int i = 0; i < d; ++i) h[i] = static_cast<double>(f * cov[i + i * d]);
    return 0;
}

// ProductKDE::_fit lognorm (kde/ProductKDE.hpp:190-192) and ProductKDE::_logl_impl (276-296) with the kernels
// `logl_values_1d_mat` / `add_logl_values_1d_mat` (KDE.cl.src:143-170) followed by logsumexp_cols_offset:
// per pair  l = -0.5 u_0^2 + lognorm;  l += -0.5 u_c^2 (c = 1..d-1),  u_c = (train_c - test_c) / sqrt(h_c),
// all in T except the double literal 0.5 (the products are evaluated in double, then stored to T).
template <typename T>
void product_kde_logl_T(const T* train, int64_t N, const T* test, int64_t m, int d, const double* h, T* out) {
    std::vector<T> sd(d);
    double slog = 0;
    for (int c = 0; c < d; ++c) {
        sd[c] = sizeof(T) == 8 ? static_cast<T>(std::sqrt(h[c])) : std::sqrt(static_cast<T>(h[c]));
        slog += std::log(h[c]);
    }
    T lognorm = static_cast<T>(-0.5 * d * std::log(2 * kPi) - 0.5 * slog - std::log(static_cast<double>(N)));
#pragma omp parallel
    {
        std::vector<T> col(N), work(N);
#pragma omp for schedule(dynamic, 8)
        for (int64_t t = 0; t < m; ++t) {
            for (int64_t i = 0; i < N; ++i) {
                T u = (train[i] - test[t]) / sd[0];
                T l = static_cast<T>((-0.5 * static_cast<double>(u)) * static_cast<double>(u) + static_cast<double>(lognorm));
                for (int c = 1; c < d; ++c) {
                    T v = (train[i + static_cast<size_t>(c) * N] - test[t + static_cast<size_t>(c) * m]) / sd[c];
                    l = static_cast<T>(static_cast<double>(l) + (-0.5 * static_cast<double>(v)) * static_cast<double>(v));
                }
                col[i] = l;
            }
            work = col;
            T mx = tree_reduce<T, true>(work);
            for (int64_t i = 0; i < N; ++i) col[i] = exp_T<T>(col[i] - mx);
            T s = tree_reduce<T, false>(col);
            col.resize(N);
            out[t] = log_T<T>(s) + mx;
        }
    }
}

// ---- LinearGaussianCPD ------------------------------------------------------------
// Householder QR least squares with column pivoting (stands in for Eigen's
// `colPivHouseholderQr().solve`, mle_LinearGaussianCPD.hpp:166).  A is n x q col-major (consumed).
template <typename T>
void lstsq_qr(std::vector<T>& A, int64_t n, int q, std::vector<T> y, std::vector<T>& beta) {
    std::vector<int> perm(q);
    std::iota(perm.begin(), perm.end(), 0);
    std::vector<T> norms(q);
    for (int j = 0; j < q; ++j) {
        T s = 0;
        for (int64_t i = 0; i < n; ++i) s += A[i + j * n] * A[i + j * n];
        norms[j] = s;
    }
    int rank = q;
    for (int k = 0; k < q && k < n; ++k) {
        int piv = k;
        for (int j = k + 1; j < q; ++j)
            if (norms[j] > norms[piv]) piv = j;
        if (piv != k) {
            for (int64_t i = 0; i < n; ++i) std::swap(A[i + k * n], A[i + piv * n]);
            std::swap(perm[k], perm[piv]);
            std::swap(norms[k], norms[piv]);
        }
        T nrm = 0;
        for (int64_t i = k; i < n; ++i) nrm += A[i + k * n] * A[i + k * n];
        nrm = std::sqrt(nrm);
        if (nrm == 0) { rank = k; break; }
        T alpha = (A[k + k * n] > 0) ? -nrm : nrm;
        std::vector<T> v(n - k);
        for (int64_t i = k; i < n; ++i) v[i - k] = A[i + k * n];
        v[0] -= alpha;
        T vnorm2 = 0;
        for (auto x : v) vnorm2 += x * x;
        if (vnorm2 > 0) {
            for (int j = k; j < q; ++j) {
                T dot = 0;
                for (int64_t i = k; i < n; ++i) dot += v[i - k] * A[i + j * n];
                T f = 2 * dot / vnorm2;
                for (int64_t i = k; i < n; ++i) A[i + j * n] -= f * v[i - k];
            }
            T dot = 0;
            for (int64_t i = k; i < n; ++i) dot += v[i - k] * y[i];
            T f = 2 * dot / vnorm2;
            for (int64_t i = k; i < n; ++i) y[i] -= f * v[i - k];
        }
        for (int j = k + 1; j < q; ++j) {
            T s = 0;
            for (int64_t i = k + 1; i < n; ++i) s += A[i + j * n] * A[i + j * n];
            norms[j] = s;
        }
    }
    std::vector<T> z(q, T(0));
    for (int k = std::min<int64_t>(rank, n) - 1; k >= 0; --k) {
        T s = y[k];
        for (int j = k + 1; j < rank; ++j) s -= A[k + j * n] * z[j];
        z[k] = s / A[k + k * n];
    }
    beta.assign(q, T(0));
    for (int j = 0; j < q; ++j) beta[perm[j]] = z[j];
}

// learning/parameters/mle_LinearGaussianCPD.hpp:11-221.  cols[0] = y, cols[1..p] = parents.
template <typename T>
void lg_fit_T(const T* const* cols, int64_t rows, int p, double* beta, double* variance) {
    const double inf = std::numeric_limits<double>::infinity();
    auto mean = [&](const T* x) { return lane_sum<T>(rows, [&](int64_t i) { return x[i]; }) / static_cast<T>(rows); };
    const T* y = cols[0];
    if (p == 0) {
        T m = mean(y);
        beta[0] = m;
        if (rows == 1) { *variance = inf; return; }
        T var = lane_sum<T>(rows, [&](int64_t i) { T dlt = y[i] - m; return dlt * dlt; });
        *variance = var / (rows - 1);
        return;
    }
    if (p == 1) {
        const T* x = cols[1];
        T my = mean(y), mx = mean(x);
        T var_x = lane_sum<T>(rows, [&](int64_t i) { T dlt = x[i] - mx; return dlt * dlt; }) / (rows - 1);
        if (var_x < kMachineTol) {
            beta[0] = my; beta[1] = 0;
            T v = lane_sum<T>(rows, [&](int64_t i) { T dlt = y[i] - my; return dlt * dlt; }) / (rows - 2);
            *variance = (rows <= 2) ? inf : static_cast<double>(v);
            return;
        }
        T cov_yx = lane_sum<T>(rows, [&](int64_t i) { return (y[i] - my) * (x[i] - mx); }) / (rows - 1);
        T b = cov_yx / var_x;
        T a = my - b * mx;
        beta[0] = a; beta[1] = b;
        if (rows <= 2) { *variance = inf; return; }
        T v = lane_sum<T>(rows, [&](int64_t i) { T r = (y[i] - my) - b * (x[i] - mx); return r * r; }) / (rows - 2);
        *variance = v;
        return;
    }
    if (p == 2) {
        const T *x1 = cols[1], *x2 = cols[2];
        T m1 = mean(x1), m2 = mean(x2), my = mean(y);
        T var1 = lane_sum<T>(rows, [&](int64_t i) { T dlt = x1[i] - m1; return dlt * dlt; }) / (rows - 1);
        bool singular1 = var1 < kMachineTol;
        T var2 = lane_sum<T>(rows, [&](int64_t i) { T dlt = x2[i] - m2; return dlt * dlt; }) / (rows - 1);
        T cov_xx = lane_sum<T>(rows, [&](int64_t i) { return (x1[i] - m1) * (x2[i] - m2); }) / (rows - 1);
        bool singular2 = var2 < kMachineTol || std::abs(cov_xx / std::sqrt(var1 * var2)) > (1 - kMachineTol);
        double var = 0;
        if (singular1) {
            if (singular2) {
                beta[0] = my; beta[1] = 0; beta[2] = 0;
                var = lane_sum<T>(rows, [&](int64_t i) { T dlt = y[i] - my; return dlt * dlt; }) / (rows - 3);
            } else {
                T cyx2 = lane_sum<T>(rows, [&](int64_t i) { return (y[i] - my) * (x2[i] - m2); }) / (rows - 1);
                T b2 = cyx2 / var2;
                beta[0] = my - b2 * m2; beta[1] = 0; beta[2] = b2;
                var = lane_sum<T>(rows, [&](int64_t i) { T r = (y[i] - my) - b2 * (x2[i] - m2); return r * r; }) / (rows - 3);
            }
        } else {
            if (singular2) {
                T cyx1 = lane_sum<T>(rows, [&](int64_t i) { return (y[i] - my) * (x1[i] - m1); }) / (rows - 1);
                T b1 = cyx1 / var1;
                beta[0] = my - b1 * m1; beta[1] = b1; beta[2] = 0;
                var = lane_sum<T>(rows, [&](int64_t i) { T r = (y[i] - my) - b1 * (x1[i] - m1); return r * r; }) / (rows - 3);
            } else {
                T cyx1 = lane_sum<T>(rows, [&](int64_t i) { return (y[i] - my) * (x1[i] - m1); }) / (rows - 1);
                T cyx2 = lane_sum<T>(rows, [&](int64_t i) { return (y[i] - my) * (x2[i] - m2); }) / (rows - 1);
                T den = var1 * var2 - cov_xx * cov_xx;
                T b1 = (var2 * cyx1 - cov_xx * cyx2) / den;
                T b2 = (cyx2 - b1 * cov_xx) / var2;
                beta[0] = my - b1 * m1 - b2 * m2; beta[1] = b1; beta[2] = b2;
                var = lane_sum<T>(rows, [&](int64_t i) {
                          T r = (y[i] - my) - b1 * (x1[i] - m1) - b2 * (x2[i] - m2);
                          return r * r;
                      }) / (rows - 3);
            }
        }
        *variance = (rows <= 3) ? inf : var;
        return;
    }
    int q = p + 1;
    std::vector<T> A(static_cast<size_t>(rows) * q);
    for (int64_t i = 0; i < rows; ++i) A[i] = 1;
    for (int j = 1; j < q; ++j) std::memcpy(&A[static_cast<size_t>(j) * rows], cols[j], rows * sizeof(T));
    std::vector<T> Acopy = A, yv(y, y + rows), b;
    lstsq_qr<T>(Acopy, rows, q, yv, b);
    for (int j = 0; j < q; ++j) beta[j] = b[j];
    if (rows <= q) { *variance = inf; return; }
    T v = lane_sum<T>(rows, [&](int64_t i) {
              T r = 0;
              for (int j = 0; j < q; ++j) r += A[i + static_cast<size_t>(j) * rows] * b[j];
              T e = y[i] - r;
              return e * e;
          }) / (rows - q);
    *variance = v;
}

// factors/continuous/LinearGaussianCPD.cpp:92-120 (logl_impl) and 138-149 (slogl = .sum()).
template <typename T>
void lg_logl_T(const T* const* cols, int64_t m, int p, const double* beta, double variance, T* out) {
    double inv_std = 1 / std::sqrt(variance);
    T cst = static_cast<T>(-0.5 * std::log(variance) - 0.5 * std::log(2 * static_cast<double>(static_cast<T>(kPi))));
    for (int64_t i = 0; i < m; ++i) {
        T mean = static_cast<T>(beta[0]);
        for (int j = 1; j <= p; ++j) mean += static_cast<T>(beta[j]) * cols[j][i];
        T z = static_cast<T>(inv_std * static_cast<double>(cols[0][i] - mean));
        T l = static_cast<T>(-0.5 * static_cast<double>(z * z));
        out[i] = l + cst;
    }
}

template <typename T>
std::vector<T> gather_cols(const T* X, int64_t n_rows, int d, const int32_t* idx, int64_t n_idx) {
    std::vector<T> out(static_cast<size_t>(n_idx) * d);
    for (int c = 0; c < d; ++c)
        for (int64_t i = 0; i < n_idx; ++i) out[i + static_cast<size_t>(c) * n_idx] = X[idx[i] + static_cast<size_t>(c) * n_rows];
    return out;
}

// CKDE::_fit + _slogl (factors/continuous/CKDE.hpp:182-200, 256-287): joint KDE on
// [variable]+evidence, marginal with H[1:,1:] on the same rows, joint - marginal in T,
// `sum1d`.  X* are column-major with the variable in column 0.
template <typename T>
void ckde_logl_T(const T* train, int64_t N, const T* test, int64_t m, int d, const double* Hj, T* out) {
    kde_logl_T<T>(train, N, test, m, d, Hj, out);
    if (d > 1) {
        int p = d - 1;
        std::vector<double> Hm(p * p);
        for (int i = 0; i < p; ++i)
            for (int j = 0; j < p; ++j) Hm[i + j * p] = Hj[(i + 1) + (j + 1) * d];
        std::vector<T> marg(m);
        kde_logl_T<T>(train + N, N, test + m, m, p, Hm.data(), marg.data());
        for (int64_t t = 0; t < m; ++t) out[t] -= marg[t];  // `substract_vectors`, KDE.cl.src:236-239
    }
}

// CVLikelihood::local_score (learning/scores/cv_likelihood.cpp:11-25) for a CKDE or a
// LinearGaussianCPD node.  X: full n_rows x d column-major, variable first.
template <typename T>
int cv_score_T(const T* X, int64_t n_rows, int d, const int32_t* indices, const int32_t* limits, int k,
               int factor /*0 CKDE, 1 LinearGaussian*/, int rule, double* out) {
    double loglik = 0;
    int64_t total = limits[k];
    for (int f = 0; f < k; ++f) {
        int64_t ts = limits[f], te = limits[f + 1];
        std::vector<int32_t> tr;
        tr.insert(tr.end(), indices, indices + ts);
        tr.insert(tr.end(), indices + te, indices + total);
        std::vector<T> trainM = gather_cols<T>(X, n_rows, d, tr.data(), tr.size());
        std::vector<T> testM = gather_cols<T>(X, n_rows, d, indices + ts, te - ts);
        int64_t N = tr.size(), m = te - ts;
        if (factor == 0) {
            std::vector<double> H(d * d);
            int st = bandwidth_T<T>(trainM.data(), N, d, rule, H.data());
            if (st) return st;
            std::vector<T> l(m);
            ckde_logl_T<T>(trainM.data(), N, testM.data(), m, d, H.data(), l.data());
            loglik += slogl_T<T>(l.data(), m);
        } else {
            std::vector<const T*> ctr(d), cte(d);
            for (int c = 0; c < d; ++c) { ctr[c] = trainM.data() + static_cast<size_t>(c) * N; cte[c] = testM.data() + static_cast<size_t>(c) * m; }
            std::vector<double> beta(d);
            double var;
            lg_fit_T<T>(ctr.data(), N, d - 1, beta.data(), &var);
            std::vector<T> l(m);
            lg_logl_T<T>(cte.data(), m, d - 1, beta.data(), var, l.data());
            // Eigen `.sum()` in T
            loglik += static_cast<double>(lane_sum<T>(m, [&](int64_t i) { return l[i]; }));
        }
    }
    *out = loglik;
    return 0;
}

// ---- CKDE::cdf and CKDE::sample ----------------------------------------------------
template <typename T> inline T erfc_T(T x);
template <> inline double erfc_T<double>(double x) { return std::erfc(x); }
template <> inline float erfc_T<float>(float x) { return erfcf(x); }
template <typename T> inline T sqrt1_2();
template <> inline double sqrt1_2<double>() { return 0.70710678118654752440; }
template <> inline float sqrt1_2<float>() { return 0.70710678118654752440f; }

// The quantities CKDE::_cdf_multivariate / _sample_multivariate derive from the two bandwidths on the host
// (factors/continuous/CKDE.hpp:346-360, 586-600): L_marg = chol(H[1:,1:]), R = L_marg^-1 H[1:,0],
// cond_var = H[0,0] - |R|^2, transform = R^T L_marg^-1 (all in double).
struct CondParams {
    std::vector<double> Lm, transform;
    double cond_var;
};
inline CondParams cond_params(const double* Hj, int d) {
    int p = d - 1;
    CondParams c;
    std::vector<double> Hm(p * p);
    for (int i = 0; i < p; ++i)
        for (int j = 0; j < p; ++j) Hm[i + j * p] = Hj[(i + 1) + (j + 1) * d];
    c.Lm.resize(p * p);
    cholesky_lower(Hm.data(), p, c.Lm.data());
    std::vector<double> inv(p * p, 0.0);  // L_marg^-1 by forward substitution on the identity (column-major)
    for (int col = 0; col < p; ++col)
        for (int i = 0; i < p; ++i) {
            double v = (i == col) ? 1.0 : 0.0;
            for (int k = 0; k < i; ++k) v -= c.Lm[i + k * p] * inv[k + col * p];
            inv[i + col * p] = v / c.Lm[i + i * p];
        }
    std::vector<double> R(p, 0.0);
    for (int i = 0; i < p; ++i)
        for (int k = 0; k < p; ++k) R[i] += inv[i + k * p] * Hj[(k + 1) + 0 * d];
    double nrm = 0;
    for (int i = 0; i < p; ++i) nrm += R[i] * R[i];
    c.cond_var = Hj[0] - nrm;
    c.transform.assign(p, 0.0);
    for (int j = 0; j < p; ++j)
        for (int i = 0; i < p; ++i) c.transform[j] += R[i] * inv[i + j * p];
    return c;
}

// CKDE::_cdf (factors/continuous/CKDE.hpp:506-556), _cdf_univariate (558-592), _cdf_multivariate (594-728)
// with kernels univariate_normal_cdf, normal_cdf, conditional_means_*, exp_elementwise, product_elementwise,
// division_elementwise (kde/opencl_kernels/KDE.cl.src:241-245, 366-468) and sum_cols_offset
// (opencl/opencl_config.hpp:399-515).  train / test are column-major with the variable in column 0.
// The weights are NOT max-shifted in the reference: a test row whose weights all underflow yields 0/0 = NaN.
template <typename T>
void ckde_cdf_T(const T* train, int64_t N, const T* test, int64_t m, int d, const double* Hj, T* out) {
    if (d == 1) {
        T inv_std = static_cast<T>(1.0 / std::sqrt(Hj[0]));
        T inv_N = static_cast<T>(1.0 / N);
#pragma omp parallel
        {
            std::vector<T> col(N);
#pragma omp for schedule(dynamic, 8)
            for (int64_t t = 0; t < m; ++t) {
                col.resize(N);
                for (int64_t i = 0; i < N; ++i)
                    col[i] = static_cast<T>(inv_N * (0.5 * erfc_T<T>(sqrt1_2<T>() * inv_std * -(test[t] - train[i]))));
                out[t] = tree_reduce<T, false>(col);
            }
        }
        return;
    }
    const int p = d - 1;
    CondParams cp = cond_params(Hj, d);
    std::vector<T> Lt(p * p), tr(p);
    for (int i = 0; i < p * p; ++i) Lt[i] = static_cast<T>(cp.Lm[i]);
    for (int j = 0; j < p; ++j) tr[j] = static_cast<T>(cp.transform[j]);
    T lognorm = static_cast<T>(lognorm_const(cp.Lm.data(), p, N) + std::log(static_cast<double>(N)));
    T inv_std = static_cast<T>(1.0 / std::sqrt(cp.cond_var));
    const T* mtrain = train + N;  // evidence columns
    const T* mtest = test + m;
#pragma omp parallel
    {
        std::vector<T> W(N), mu(N), delta(p);
#pragma omp for schedule(dynamic, 8)
        for (int64_t t = 0; t < m; ++t) {
            W.resize(N);
            mu.resize(N);
            for (int64_t i = 0; i < N; ++i) {
                W[i] = exp_T<T>(pair_logl<T>(mtrain, N, i, mtest, m, t, p, Lt.data(), lognorm, delta.data()));
                T mean = train[i];
                for (int j = 0; j < p; ++j)
                    mean += tr[j] * (mtest[t + static_cast<size_t>(j) * m] - mtrain[i + static_cast<size_t>(j) * N]);
                T c = static_cast<T>(0.5 * erfc_T<T>(sqrt1_2<T>() * inv_std * (mean - test[t])));
                mu[i] = c * W[i];
            }
            T sumW = tree_reduce<T, false>(W);
            T num = tree_reduce<T, false>(mu);
            out[t] = num / sumW;
        }
    }
}

// CKDE::_sample_indices_from_weights (factors/continuous/CKDE.hpp:402-504): weights exp(logl_marg(i, t)) in T,
// exclusive prefix sum over the training rows (accum_sum_cols, opencl_config.hpp:539-579, KDE.cl.src:254-338),
// rows 1.. divided by the total (normalize_accum_sum_mat_cols, 340-348), index i in [0, N-2] with
// cum[i] <= u < cum[i+1], default N-1 (find_random_indices, 351-364).  The reference's scan is a work-group
// tree whose association depends on the device's work-group size; here the prefix sum is sequential, so an
// index can differ only when u falls within rounding error of a boundary.
template <typename T>
void ckde_sample_indices_T(const T* mtrain, int64_t N, const T* etest, int64_t n, int p, const double* Lm,
                           const T* random_prob, int32_t* out) {
    std::vector<T> Lt(p * p);
    for (int i = 0; i < p * p; ++i) Lt[i] = static_cast<T>(Lm[i]);
    T lognorm = static_cast<T>(lognorm_const(Lm, p, N));
#pragma omp parallel
    {
        std::vector<T> cum(N + 1), delta(p);
#pragma omp for schedule(dynamic, 8)
        for (int64_t t = 0; t < n; ++t) {
            cum[0] = 0;
            for (int64_t i = 0; i < N; ++i)
                cum[i + 1] = cum[i] + exp_T<T>(pair_logl<T>(mtrain, N, i, etest, n, t, p, Lt.data(), lognorm, delta.data()));
            T total = cum[N];
            for (int64_t i = 1; i < N; ++i) cum[i] /= total;
            T u = random_prob[t];
            int32_t idx = static_cast<int32_t>(N - 1);
            for (int64_t i = 0; i + 1 < N; ++i)
                if (cum[i] <= u && u < cum[i + 1]) idx = static_cast<int32_t>(i);
            out[t] = idx;
        }
    }
}

// CKDE::_sample / _sample_multivariate (factors/continuous/CKDE.hpp:289-400).  `evidence` is n x p column-major
// (ignored for d == 1).  libstdc++'s std::mt19937 / uniform_int_distribution / uniform_real_distribution /
// normal_distribution are the reference's generators (it is built with the same standard library).
template <typename T>
void ckde_sample_T(const T* train, int64_t N, int d, const double* Hj, const T* evidence, int64_t n, uint32_t seed,
                   T* out, int32_t* idx_out) {
    if (d == 1) {
        std::mt19937 rng{seed};
        std::uniform_int_distribution<> uniform(0, static_cast<int>(N - 1));
        std::normal_distribution<T> normal(0, std::sqrt(Hj[0]));
        for (int64_t i = 0; i < n; ++i) {
            auto index = uniform(rng);
            if (idx_out) idx_out[i] = index;
            out[i] = train[index] + normal(rng);
        }
        return;
    }
    const int p = d - 1;
    std::vector<T> random_prob(n);
    std::mt19937 rng{seed};
    std::uniform_real_distribution<T> uniform(0, 1);
    for (int64_t i = 0; i < n; ++i) random_prob[i] = uniform(rng);
    CondParams cp = cond_params(Hj, d);
    std::vector<int32_t> idx(n);
    ckde_sample_indices_T<T>(train + N, N, evidence, n, p, cp.Lm.data(), random_prob.data(), idx.data());
    if (idx_out) std::copy(idx.begin(), idx.end(), idx_out);
    std::vector<T> tr(p);
    for (int j = 0; j < p; ++j) tr[j] = static_cast<T>(cp.transform[j]);
    // cond_mean = evidence_substract * transform (Eigen row-times-vector product in T), CKDE.hpp:375-384
    std::vector<T> cond_mean(n);
    for (int64_t i = 0; i < n; ++i) {
        T acc = 0;
        for (int j = 0; j < p; ++j)
            acc += (evidence[i + static_cast<size_t>(j) * n] - train[idx[i] + static_cast<size_t>(j + 1) * N]) * tr[j];
        cond_mean[i] = acc;
    }
    std::normal_distribution<T> normal(0, std::sqrt(cp.cond_var));
    for (int64_t i = 0; i < n; ++i) {
        cond_mean[i] += train[idx[i]] + normal(rng);
        out[i] = cond_mean[i];
    }
}

}  // namespace

extern "C" {

int orc_num_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU-baseline legs of bench.py ask for all host threads back
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int orc_cov(const void* X, int64_t n, int d, int dtype, double* cov_out) {
    if (dtype == 0) {
        std::vector<double> c;
        cov_T<double>(static_cast<const double*>(X), n, d, c);
        std::copy(c.begin(), c.end(), cov_out);
    } else {
        std::vector<float> c;
        cov_T<float>(static_cast<const float*>(X), n, d, c);
        for (size_t i = 0; i < c.size(); ++i) cov_out[i] = c[i];
    }
    return 0;
}

int orc_bandwidth(const void* X, int64_t n, int d, int dtype, int rule, double* H) {
    return dtype == 0 ? bandwidth_T<double>(static_cast<const double*>(X), n, d, rule, H)
                      : bandwidth_T<float>(static_cast<const float*>(X), n, d, rule, H);
}

// lognorm + Cholesky of a bandwidth (kde/KDE.hpp:459-477). Returns 1 if H is not PD.
int orc_kde_prepare(const double* H, int d, int64_t N, double* L, double* lognorm) {
    if (!cholesky_lower(H, d, L)) return 1;
    *lognorm = lognorm_const(L, d, N);
    return 0;
}

int orc_kde_logl(const void* train, int64_t N, const void* test, int64_t m, int d, int dtype, const double* H,
                 double* out_logl, double* out_slogl) {
    if (dtype == 0) {
        std::vector<double> l(m);
        kde_logl_T<double>(static_cast<const double*>(train), N, static_cast<const double*>(test), m, d, H, l.data());
        if (out_logl) std::copy(l.begin(), l.end(), out_logl);
        if (out_slogl) *out_slogl = slogl_T<double>(l.data(), m);
    } else {
        std::vector<float> l(m);
        kde_logl_T<float>(static_cast<const float*>(train), N, static_cast<const float*>(test), m, d, H, l.data());
        if (out_logl) for (int64_t i = 0; i < m; ++i) out_logl[i] = l[i];
        if (out_slogl) *out_slogl = slogl_T<float>(l.data(), m);
    }
    return 0;
}

int orc_ckde_logl(const void* train, int64_t N, const void* test, int64_t m, int d, int dtype, const double* Hjoint,
                  double* out_logl, double* out_slogl) {
    if (dtype == 0) {
        std::vector<double> l(m);
        ckde_logl_T<double>(static_cast<const double*>(train), N, static_cast<const double*>(test), m, d, Hjoint, l.data());
        if (out_logl) std::copy(l.begin(), l.end(), out_logl);
        if (out_slogl) *out_slogl = slogl_T<double>(l.data(), m);
    } else {
        std::vector<float> l(m);
        ckde_logl_T<float>(static_cast<const float*>(train), N, static_cast<const float*>(test), m, d, Hjoint, l.data());
        if (out_logl) for (int64_t i = 0; i < m; ++i) out_logl[i] = l[i];
        if (out_slogl) *out_slogl = slogl_T<float>(l.data(), m);
    }
    return 0;
}

int orc_diag_bandwidth(const void* X, int64_t n, int d, int dtype, int rule, double* h) {
    return dtype == 0 ? diag_bandwidth_T<double>(static_cast<const double*>(X), n, d, rule, h)
                      : diag_bandwidth_T<float>(static_cast<const float*>(X), n, d, rule, h);
}

int orc_product_kde_logl(const void* train, int64_t N, const void* test, int64_t m, int d, int dtype, const double* h,
                         double* out_logl, double* out_slogl) {
    if (dtype == 0) {
        std::vector<double> l(m);
        product_kde_logl_T<double>(static_cast<const double*>(train), N, static_cast<const double*>(test), m, d, h, l.data());
        if (out_logl) std::copy(l.begin(), l.end(), out_logl);
        if (out_slogl) *out_slogl = slogl_T<double>(l.data(), m);
    } else {
        std::vector<float> l(m);
        product_kde_logl_T<float>(static_cast<const float*>(train), N, static_cast<const float*>(test), m, d, h, l.data());
        if (out_logl) for (int64_t i = 0; i < m; ++i) out_logl[i] = l[i];
        if (out_slogl) *out_slogl = slogl_T<float>(l.data(), m);
    }
    return 0;
}

// Independent check of the restatement: direct evaluation in long double
// (x87 80-bit) with a plain two-pass log-sum-exp.  Inputs are widened first.
int orc_kde_logl_ld(const void* train, int64_t N, const void* test, int64_t m, int d, int dtype, const double* H,
                    double* out_logl) {
    std::vector<long double> L(d * d, 0.0L);
    for (int j = 0; j < d; ++j) {
        long double s = H[j + j * d];
        for (int k = 0; k < j; ++k) s -= L[j + k * d] * L[j + k * d];
        if (!(s > 0)) return 1;
        L[j + j * d] = sqrtl(s);
        for (int i = j + 1; i < d; ++i) {
            long double t = H[i + j * d];
            for (int k = 0; k < j; ++k) t -= L[i + k * d] * L[j + k * d];
            L[i + j * d] = t / L[j + j * d];
        }
    }
    long double slog = 0;
    for (int i = 0; i < d; ++i) slog += logl(L[i + i * d]);
    long double lognorm = -slog - 0.5L * d * logl(2 * 3.141592653589793238462643383279502884L) - logl((long double)N);
    auto at = [&](const void* p, int64_t idx) -> long double {
        return dtype == 0 ? (long double)static_cast<const double*>(p)[idx] : (long double)static_cast<const float*>(p)[idx];
    };
#pragma omp parallel
    {
        std::vector<long double> v(N), delta(d);
#pragma omp for schedule(dynamic, 8)
        for (int64_t t = 0; t < m; ++t) {
            long double mx = -INFINITY;
            for (int64_t i = 0; i < N; ++i) {
                for (int c = 0; c < d; ++c) delta[c] = at(test, t + (int64_t)c * m) - at(train, i + (int64_t)c * N);
                long double s = 0;
                for (int c = 0; c < d; ++c) {
                    for (int k = 0; k < c; ++k) delta[c] -= L[c + k * d] * delta[k];
                    delta[c] /= L[c + c * d];
                    s += delta[c] * delta[c];
                }
                v[i] = -0.5L * s;
                mx = std::max(mx, v[i]);
            }
            long double acc = 0;
            for (int64_t i = 0; i < N; ++i) acc += expl(v[i] - mx);
            out_logl[t] = (double)(lognorm + mx + logl(acc));
        }
    }
    return 0;
}

int orc_ucv_score_unconstrained(const void* X, int64_t N, int d, int dtype, const double* H, double* out) {
    *out = dtype == 0 ? ucv_unconstrained_T<double>(static_cast<const double*>(X), N, d, H)
                      : ucv_unconstrained_T<float>(static_cast<const float*>(X), N, d, H);
    return 0;
}

int orc_ucv_score_diagonal(const void* X, int64_t N, int d, int dtype, const double* hdiag, double* out) {
    *out = dtype == 0 ? ucv_diagonal_T<double>(static_cast<const double*>(X), N, d, hdiag)
                      : ucv_diagonal_T<float>(static_cast<const float*>(X), N, d, hdiag);
    return 0;
}

// CrossValidationProperties (dataset/crossvalidation_adaptator.hpp:15-67).
// `indices` holds the valid row ids on entry (iota, or the non-null rows) and the
// shuffled ids on exit; `limits` gets k+1 entries.  Uses libstdc++'s std::shuffle,
// the same routine the reference is compiled against.
int orc_cv_indices(int32_t* indices, int64_t n, int k, uint32_t seed, int32_t* limits) {
    std::vector<int> v(indices, indices + n);
    std::mt19937 rng{seed};
    std::shuffle(v.begin(), v.end(), rng);
    std::copy(v.begin(), v.end(), indices);
    int fold_size = static_cast<int>(n / k), extra = static_cast<int>(n % k);
    int cur = 0, pos = 0;
    limits[pos++] = 0;
    for (int i = 0; i < extra; ++i) { cur += fold_size + 1; limits[pos++] = cur; }
    for (int i = extra; i < k; ++i) { cur += fold_size; limits[pos++] = cur; }
    return 0;
}

// HoldOut (dataset/holdout_adaptator.hpp:17-70): shuffle, test_rows = round(n*ratio),
// train = first n - test_rows shuffled ids.
int orc_holdout_indices(int32_t* indices, int64_t n, double test_ratio, uint32_t seed, int32_t* n_train) {
    std::vector<int> v(indices, indices + n);
    std::mt19937 rng{seed};
    std::shuffle(v.begin(), v.end(), rng);
    std::copy(v.begin(), v.end(), indices);
    int test_rows = static_cast<int>(std::round(n * test_ratio));
    *n_train = static_cast<int32_t>(n - test_rows);
    return 0;
}

int orc_lg_fit(const void* const* cols, int64_t rows, int p, int dtype, double* beta, double* variance) {
    if (dtype == 0) lg_fit_T<double>(reinterpret_cast<const double* const*>(cols), rows, p, beta, variance);
    else lg_fit_T<float>(reinterpret_cast<const float* const*>(cols), rows, p, beta, variance);
    return 0;
}

int orc_lg_logl(const void* const* cols, int64_t m, int p, int dtype, const double* beta, double variance,
                double* out_logl, double* out_slogl) {
    if (dtype == 0) {
        std::vector<double> l(m);
        lg_logl_T<double>(reinterpret_cast<const double* const*>(cols), m, p, beta, variance, l.data());
        if (out_logl) std::copy(l.begin(), l.end(), out_logl);
        if (out_slogl) *out_slogl = lane_sum<double>(m, [&](int64_t i) { return l[i]; });
    } else {
        std::vector<float> l(m);
        lg_logl_T<float>(reinterpret_cast<const float* const*>(cols), m, p, beta, variance, l.data());
        if (out_logl) for (int64_t i = 0; i < m; ++i) out_logl[i] = l[i];
        if (out_slogl) *out_slogl = lane_sum<float>(m, [&](int64_t i) { return l[i]; });
    }
    return 0;
}

int orc_cv_score(const void* X, int64_t n_rows, int d, int dtype, const int32_t* indices, const int32_t* limits, int k,
                 int factor, int rule, double* out) {
    return dtype == 0 ? cv_score_T<double>(static_cast<const double*>(X), n_rows, d, indices, limits, k, factor, rule, out)
                      : cv_score_T<float>(static_cast<const float*>(X), n_rows, d, indices, limits, k, factor, rule, out);
}

// ---- containers whose libstdc++ behaviour the reference's hill climbing depends on ----------------
// std::unordered_set<int> iteration order = order of BayesianNetwork::parents() (graph/graph_types.hpp:12-51)
void* orc_uset_new() { return new std::unordered_set<int>(); }
void* orc_uset_clone(const void* s) { return new std::unordered_set<int>(*static_cast<const std::unordered_set<int>*>(s)); }
void orc_uset_free(void* s) { delete static_cast<std::unordered_set<int>*>(s); }
void orc_uset_insert(void* s, int v) { static_cast<std::unordered_set<int>*>(s)->insert(v); }
void orc_uset_erase(void* s, int v) { static_cast<std::unordered_set<int>*>(s)->erase(v); }
int orc_uset_size(const void* s) { return static_cast<int>(static_cast<const std::unordered_set<int>*>(s)->size()); }
void orc_uset_list(const void* s, int* out) {
    int i = 0;
    for (auto v : *static_cast<const std::unordered_set<int>*>(s)) out[i++] = v;
}
// ArcOperatorSet::find_max_indegree: std::sort(sorted_idx, delta desc) (learning/operators/operators.hpp:494)
void orc_sort_desc(int* idx, int64_t n, const double* delta) {
    std::vector<int> v(idx, idx + n);
    std::sort(v.begin(), v.end(), [&delta](auto i1, auto i2) { return delta[i1] > delta[i2]; });
    std::copy(v.begin(), v.end(), idx);
}


// ---- CKDE::cdf / CKDE::sample, LinearGaussianCPD::cdf / sample ----
int orc_ckde_cdf(const void* train, int64_t N, const void* test, int64_t m, int d, int dtype, const double* Hjoint,
                 double* out) {
    if (dtype == 0) {
        ckde_cdf_T<double>(static_cast<const double*>(train), N, static_cast<const double*>(test), m, d, Hjoint, out);
    } else {
        std::vector<float> r(m);
        ckde_cdf_T<float>(static_cast<const float*>(train), N, static_cast<const float*>(test), m, d, Hjoint, r.data());
        for (int64_t i = 0; i < m; ++i) out[i] = r[i];
    }
    return 0;
}

// out: n values of the data's dtype; idx_out (may be NULL): the sampled training-row indices
int orc_ckde_sample(const void* train, int64_t N, int d, int dtype, const double* Hjoint, const void* evidence,
                    int64_t n, uint32_t seed, void* out, int32_t* idx_out) {
    if (dtype == 0)
        ckde_sample_T<double>(static_cast<const double*>(train), N, d, Hjoint, static_cast<const double*>(evidence), n,
                              seed, static_cast<double*>(out), idx_out);
    else
        ckde_sample_T<float>(static_cast<const float*>(train), N, d, Hjoint, static_cast<const float*>(evidence), n, seed,
                             static_cast<float*>(out), idx_out);
    return 0;
}

// the uniform draws of CKDE::_sample_multivariate (CKDE.hpp:336-341), for tests of the index kernel alone
int orc_uniform_real(int64_t n, uint32_t seed, int dtype, void* out) {
    std::mt19937 rng{seed};
    if (dtype == 0) {
        std::uniform_real_distribution<double> u(0, 1);
        for (int64_t i = 0; i < n; ++i) static_cast<double*>(out)[i] = u(rng);
    } else {
        std::uniform_real_distribution<float> u(0, 1);
        for (int64_t i = 0; i < n; ++i) static_cast<float*>(out)[i] = u(rng);
    }
    return 0;
}

int orc_ckde_sample_indices(const void* mtrain, int64_t N, const void* etest, int64_t n, int p, int dtype,
                            const double* Hmarg, const void* random_prob, int32_t* out) {
    std::vector<double> Lm(p * p);
    if (!cholesky_lower(Hmarg, p, Lm.data())) return 1;
    if (dtype == 0)
        ckde_sample_indices_T<double>(static_cast<const double*>(mtrain), N, static_cast<const double*>(etest), n, p,
                                      Lm.data(), static_cast<const double*>(random_prob), out);
    else
        ckde_sample_indices_T<float>(static_cast<const float*>(mtrain), N, static_cast<const float*>(etest), n, p,
                                     Lm.data(), static_cast<const float*>(random_prob), out);
    return 0;
}

// LinearGaussianCPD::sample (factors/continuous/LinearGaussianCPD.cpp:317-372): always double output;
// ev[j] is evidence column j (dtype of the evidence frame), beta[0] the intercept.
int orc_lg_sample(const double* beta, double variance, int p, const void* const* ev, int ev_dtype, int64_t n,
                  uint32_t seed, double* out) {
    std::mt19937 rng{seed};
    std::normal_distribution<> normal(beta[0], std::sqrt(variance));
    for (int64_t i = 0; i < n; ++i) out[i] = normal(rng);
    for (int j = 0; j < p; ++j)
        for (int64_t i = 0; i < n; ++i)
            out[i] += beta[j + 1] * (ev_dtype == 0 ? static_cast<const double*>(ev[j])[i]
                                                    : static_cast<double>(static_cast<const float*>(ev[j])[i]));
    return 0;
}

}  // extern "C"
