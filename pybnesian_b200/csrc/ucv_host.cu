// ucv_host.cu — UCVScorer and the UCV bandwidth selector (host side) on top of ucv_kernel.cu.
//
// Reference: kde/UCV.hpp:12-56, kde/UCV.cpp:209-525 (UCVScorer::score_*, wrap_ucv_optim,
// wrap_ucv_diag_optim, UCV::bandwidth / diag_bandwidth), util/vech_ops.cpp:6-69.
// The optimiser the reference calls is NLopt 2.7.1 `LN_NELDERMEAD` (un-vendored third party,
// vcpkg pin in /root/reference/vcpkg.json); `nelder_mead` below restates that algorithm
// (S. G. Johnson's neldermead.c: alpha=1, beta=1/2, gamma=2, delta=1/2, default initial step
// |x_i|, ftol_rel / xtol_rel stopping rules).  Trajectory parity with NLopt is UNPINNED
// (no reference test covers UCV); the objective itself is pinned by the oracle.
#include <functional>

#include "internal.h"

struct pbn_ucv {
    std::vector<pbn_ucv*> rep;  // replicas on ctx->peers (multi-device context): each sums its slice of the pair tiles
    pbn_ctx* ctx;
    const pbn_table* tbl;
    int cols[PBN_MAX_DIM];
    int d;
    int dtype;
    pbn_rows rows;
    int64_t n;
    double mu[PBN_MAX_DIM];
    void* y;            // whitened rows (re-filled per score call)
    float* d_bound;
    double* nrm;        // f64, d <= 8 only (else null): -sum_c y_c^2 per row, re-filled with y
    long long* d_prefix;
    int n_row_tiles;
    long long total_units;
    double* d_partial;
    int max_grid;
    int64_t evals;
};

static int ucv_sums(pbn_ucv* s, const double* Lchol /*col-major lower*/, long long part, long long nparts, double* S2,
                    double* S1) {
    pbn_ctx* ctx = s->ctx;
    const int d = s->d;
    cudaStream_t st = ctx->stream;
    std::vector<double> Winv(d * d), W(d * d);
    tri_inverse_rowmajor(Lchol, d, Winv.data());
    // exponent of exp(-s/4): half the pair kernel's scale
    double c = sqrt(0.25 * unit_scale(s->dtype));
    for (int i = 0; i < d * d; ++i) W[i] = c * Winv[i];
    PBN_CUDA_TRY(cudaMemsetAsync(s->d_bound, 0, sizeof(float), st));
    PBN_TRY(whiten_raw_launch(ctx, s->tbl, s->cols, d, s->rows, W.data(), s->mu, s->y, s->d_bound, s->nrm, d));
    long long ub = s->total_units * part / nparts, ue = s->total_units * (part + 1) / nparts;
    long long U = ue - ub;
    *S2 = 0;
    *S1 = 0;
    if (U <= 0) return PBN_OK;
    int grid = (int)std::min<long long>(U, s->max_grid);
    long long upb = (U + grid - 1) / grid;
    grid = (int)((U + upb - 1) / upb);
    pbn::UcvJob job;
    job.y = s->y;
    job.n = s->n;
    job.prefix = s->d_prefix;
    job.n_row_tiles = s->n_row_tiles;
    job.bound = s->d_bound;
    job.nrm = s->nrm;
    job.partial = s->d_partial;
    job.unit_begin = ub;
    job.unit_end = ue;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (ctx->timing) {
        PBN_CUDA_TRY(cudaEventCreate(&ev0));
        PBN_CUDA_TRY(cudaEventCreate(&ev1));
        PBN_CUDA_TRY(cudaEventRecord(ev0, st));
    }
    PBN_CUDA_TRY(pbn::launch_ucv(s->dtype == PBN_F64, d, job, upb, grid, ctx->d_exp_tab, st));
    ctx->launches++;
    if (ctx->timing) {
        PBN_CUDA_TRY(cudaEventRecord(ev1, st));
        ctx->timed.emplace_back(ev0, ev1);
        ctx->pair_units += (int64_t)((double)s->n * (s->n - 1) / 2 * (double)U / (double)s->total_units);
    }
    std::vector<double> h(2 * (size_t)grid);
    PBN_CUDA_TRY(cudaMemcpyAsync(h.data(), s->d_partial, h.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
    PBN_CUDA_TRY(cudaStreamSynchronize(st));
    ctx->d2h += (int64_t)h.size() * 8;
    double a = 0, b = 0;
    for (int g = 0; g < grid; ++g) { a += h[2 * g]; b += h[2 * g + 1]; }
    *S2 = a;
    *S1 = b;
    s->evals++;
    return PBN_OK;
}

// N * UCV(H) = e^{c2} + 2/N * e^{c2} * S2 - 4/(N-1) * e^{c1} * S1   (kde/UCV.cpp:357)
static double ucv_combine(const pbn_ucv* s, const double* Lchol, double S2, double S1) {
    const int d = s->d;
    double slog = 0;
    for (int i = 0; i < d; ++i) slog += log(Lchol[i + i * d]);
    const double log2pi = 1.8378770664093454836;
    double c1 = -slog - 0.5 * d * log2pi;
    double c2 = c1 - 0.5 * d * log(2.0);
    double N = (double)s->n;
    return exp(c2) + 2.0 * exp(c2) * S2 / N - 4.0 * exp(c1) * S1 / (N - 1.0);
}

// Cholesky factor used by a score call: diag(sqrt(h)) or chol(H), H cast to the data type first
// (UCV.cpp:388-392 casts the bandwidth to the data type)
static int ucv_factor(const pbn_ucv* s, const double* H, int is_diag, std::vector<double>& L) {
    const int d = s->d;
    L.assign(d * d, 0.0);
    if (is_diag) {
        for (int i = 0; i < d; ++i) {
            if (!(H[i] > 0)) return set_error(PBN_ERR_SINGULAR, "diagonal bandwidth must be positive");
            L[i + i * d] = sqrt(H[i]);
        }
        return PBN_OK;
    }
    std::vector<double> Hc(H, H + d * d);
    if (s->dtype == PBN_F32)
        for (auto& v : Hc) v = (double)(float)v;
    if (!chol_lower(Hc.data(), d, L.data())) return set_error(PBN_ERR_SINGULAR, "bandwidth matrix is not positive definite");
    return PBN_OK;
}

// slice `part` of `nparts` of the pair-tile schedule; a multi-device scorer cuts that slice once more, one piece per
// device (upper-triangular tile list split by tile count), and adds the pieces in device order
static int ucv_sums_devices(pbn_ucv* s, const double* Lchol, long long part, long long nparts, double* S2, double* S1) {
    // (below ~2e8 pairs one device finishes before the threads have started)
    if (!pbn_replicated(s->ctx, s) || s->n < 20000) return ucv_sums(s, Lchol, part, nparts, S2, S1);
    const int nd = pbn_num_devices(s->ctx);
    std::vector<double> a(nd, 0.0), b(nd, 0.0);
    PBN_TRY(pbn_run_on_devices(nd, [&](int i) {
        pbn_ucv* r = pbn_replica(s, i);
        DevSetter ds(r->ctx->device);
        return ucv_sums(r, Lchol, part * nd + i, nparts * nd, &a[i], &b[i]);
    }));
    *S2 = 0;
    *S1 = 0;
    for (int i = 0; i < nd; ++i) {
        *S2 += a[i];
        *S1 += b[i];
    }
    return PBN_OK;
}

static int ucv_score_full(pbn_ucv* s, const double* H, int is_diag, double* out) {
    std::vector<double> L;
    PBN_TRY(ucv_factor(s, H, is_diag, L));
    double S2, S1;
    PBN_TRY(ucv_sums_devices(s, L.data(), 0, 1, &S2, &S1));
    *out = ucv_combine(s, L.data(), S2, S1);
    return PBN_OK;
}

// ---- Nelder-Mead, NLopt-2.7.1 style (see header) ---------------------------------------
namespace {
struct NMResult {
    std::vector<double> x;
    double f;
    int evals;
};

inline bool relstop(double vold, double vnew, double reltol, double abstol) {
    if (std::isinf(vold)) return false;
    return fabs(vnew - vold) < abstol || fabs(vnew - vold) < reltol * (fabs(vnew) + fabs(vold)) * 0.5 ||
           (reltol > 0 && vnew == vold);
}

// xnew = c + scale * (c - xold); false if no progress is possible
inline bool reflectpt(int n, double* xnew, const double* c, double scale, const double* xold) {
    bool equalc = true, equalold = true;
    for (int i = 0; i < n; ++i) {
        double v = c[i] + scale * (c[i] - xold[i]);
        equalc = equalc && fabs(v - c[i]) <= 1e-13 * (fabs(v) + fabs(c[i]));
        equalold = equalold && fabs(v - xold[i]) <= 1e-13 * (fabs(v) + fabs(xold[i]));
        xnew[i] = v;
    }
    return !(equalc || equalold);
}

NMResult nelder_mead(const std::function<double(const double*)>& f, std::vector<double> x0, double ftol_rel,
                     double xtol_rel, int max_evals) {
    const int n = (int)x0.size();
    const double alpha = 1, beta = 0.5, gamm = 2, delta = 0.5;
    NMResult best;
    int evals = 0;
    auto eval = [&](const double* x) {
        double v = f(x);
        ++evals;
        if (best.x.empty() || v < best.f) { best.f = v; best.x.assign(x, x + n); }
        return v;
    };
    std::vector<std::vector<double>> pts(n + 1, std::vector<double>(n));
    std::vector<double> fv(n + 1);
    pts[0] = x0;
    fv[0] = eval(x0.data());
    for (int i = 0; i < n; ++i) {
        pts[i + 1] = x0;
        double step = fabs(x0[i]);
        if (step == 0 || std::isinf(step)) step = 1;
        pts[i + 1][i] += step;
        fv[i + 1] = eval(pts[i + 1].data());
    }
    std::vector<double> c(n), xcur(n, 0.0), tmp(n);
    bool have_xcur = false;
    while (evals < max_evals) {
        // order: (f, index) as NLopt's red-black tree does
        int lo = 0, hi = 0;
        for (int i = 1; i <= n; ++i) {
            if (fv[i] < fv[lo]) lo = i;
            if (fv[i] >= fv[hi]) hi = i;
        }
        int second = -1;  // predecessor of high
        for (int i = 0; i <= n; ++i) {
            if (i == hi) continue;
            if (second < 0 || fv[i] > fv[second] || (fv[i] == fv[second] && i > second)) second = i;
        }
        double fl = fv[lo], fh = fv[hi];
        if (relstop(fl, fh, ftol_rel, 0.0)) break;
        std::fill(c.begin(), c.end(), 0.0);
        for (int i = 0; i <= n; ++i)
            if (i != hi)
                for (int k = 0; k < n; ++k) c[k] += pts[i][k];
        for (int k = 0; k < n; ++k) c[k] /= n;
        if (have_xcur) {  // nlopt_stop_x: L1 norms
            double dn = 0, vn = 0;
            for (int k = 0; k < n; ++k) { dn += fabs(c[k] - xcur[k]); vn += fabs(c[k]); }
            if (dn <= xtol_rel * vn) break;
        }
        xcur = c;
        have_xcur = true;
        if (!reflectpt(n, xcur.data(), c.data(), alpha, pts[hi].data())) break;
        double fr = eval(xcur.data());
        if (fr < fl) {
            if (!reflectpt(n, tmp.data(), c.data(), gamm, pts[hi].data())) break;
            double fe = eval(tmp.data());
            if (fe >= fr) { pts[hi] = xcur; fv[hi] = fr; }
            else { pts[hi] = tmp; fv[hi] = fe; }
        } else if (fr < fv[second]) {
            pts[hi] = xcur;
            fv[hi] = fr;
        } else {
            if (!reflectpt(n, xcur.data(), c.data(), fh <= fr ? -beta : beta, pts[hi].data())) break;
            double fc = eval(xcur.data());
            if (fc < fr && fc < fh) {
                pts[hi] = xcur;
                fv[hi] = fc;
            } else {
                bool ok = true;
                for (int i = 0; i <= n && ok; ++i) {
                    if (i == lo) continue;
                    std::vector<double> old = pts[i];
                    if (!reflectpt(n, pts[i].data(), pts[lo].data(), -delta, old.data())) { ok = false; break; }
                    fv[i] = eval(pts[i].data());
                }
                if (!ok) break;
            }
        }
    }
    best.evals = evals;
    return best;
}
}  // namespace

extern "C" {

static int ucv_create_one(pbn_ctx* ctx, const pbn_table* tbl, const int* cols, int d, pbn_rows rows, pbn_ucv** out);

int pbn_ucv_create(pbn_ctx* ctx, const pbn_table* tbl, const int* cols, int d, pbn_rows rows, pbn_ucv** out) {
    if (!ctx || !out) return set_error(PBN_ERR_ARG, "null argument");
    if (!pbn_replicated(ctx, tbl)) return ucv_create_one(ctx, tbl, cols, d, rows, out);
    const int nd = pbn_num_devices(ctx);
    std::vector<pbn_ucv*> u(nd, nullptr);
    int rc = pbn_run_on_devices(nd, [&](int i) {
        return ucv_create_one(pbn_device_ctx(ctx, i), pbn_replica(const_cast<pbn_table*>(tbl), i), cols, d, rows, &u[i]);
    });
    if (rc != PBN_OK) {
        for (pbn_ucv* q : u)
            if (q) pbn_ucv_free(q);
        return rc;
    }
    u[0]->rep.assign(u.begin() + 1, u.end());
    *out = u[0];
    return PBN_OK;
}

static int ucv_create_one(pbn_ctx* ctx, const pbn_table* tbl, const int* cols, int d, pbn_rows rows, pbn_ucv** out) {
    if (!ctx || !out) return set_error(PBN_ERR_ARG, "null argument");
    PBN_TRY(check_cols(tbl, cols, d));
    PBN_TRY(check_rows(tbl, rows));
    if (d > 8) return set_error(PBN_ERR_UNSUPPORTED, "UCV scoring supports at most 8 variables");
    int64_t n = seg_count(rows);
    if (n < 2) return set_error(PBN_ERR_ARG, "UCV needs at least 2 instances");
    DevSetter ds(ctx->device);
    pbn_ucv* s = new pbn_ucv();  // value-initialised: every scalar member starts at zero
    s->ctx = ctx;
    s->tbl = tbl;
    s->d = d;
    s->dtype = tbl->dtype;
    s->rows = rows;
    s->n = n;
    for (int i = 0; i < d; ++i) s->cols[i] = cols[i];
    int rc = moments_impl(ctx, tbl, cols, d, rows, s->mu, nullptr);
    if (rc != PBN_OK) { delete s; return rc; }
    const bool f64 = s->dtype == PBN_F64;
    const int TILE = f64 ? pbn::pair_tile_f64(d) : pbn::pair_tile_f32(d);
    const int TB = f64 ? pbn::pair_tb_f64() : pbn::pair_tb_f32();
    s->n_row_tiles = (int)((n + TB - 1) / TB);
    std::vector<long long> prefix(s->n_row_tiles + 1, 0);
    for (int tt = 0; tt < s->n_row_tiles; ++tt) {
        long long row_hi = std::min<long long>(n, (long long)(tt + 1) * TB);
        prefix[tt + 1] = prefix[tt] + (row_hi + TILE - 1) / TILE;
    }
    s->total_units = prefix.back();
    s->max_grid = ctx->sm_count * 2;
    size_t es = elem_size(s->dtype);
    size_t ybytes = (((size_t)((n + TILE - 1) / TILE * TILE + 16) * d * es) + 255) / 256 * 256;
    cudaStream_t st = ctx->stream;
    const size_t nbytes = (f64 && d <= 8) ? (((size_t)((n + TILE - 1) / TILE * TILE + 16) * sizeof(double)) + 255) / 256 * 256 : 0;
    cudaError_t e = cudaMallocAsync(&s->y, ybytes + 256 + nbytes, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(s->y, 0, ybytes + 256 + nbytes, st);
    if (e == cudaSuccess) e = cudaMallocAsync(&s->d_prefix, prefix.size() * sizeof(long long), st);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(s->d_prefix, prefix.data(), prefix.size() * sizeof(long long), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMallocAsync(&s->d_partial, 2 * (size_t)s->max_grid * sizeof(double), st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        if (s->y) cudaFreeAsync(s->y, st);
        if (s->d_prefix) cudaFreeAsync(s->d_prefix, st);
        if (s->d_partial) cudaFreeAsync(s->d_partial, st);
        delete s;
        PBN_CUDA_TRY(e);
    }
    s->d_bound = reinterpret_cast<float*>(static_cast<char*>(s->y) + ybytes);
    s->nrm = nbytes ? reinterpret_cast<double*>(static_cast<char*>(s->y) + ybytes + 256) : nullptr;
    *out = s;
    return PBN_OK;
}

int pbn_ucv_free(pbn_ucv* s) {
    if (!s) return PBN_OK;
    for (pbn_ucv* r : s->rep) pbn_ucv_free(r);
    s->rep.clear();
    DevSetter ds(s->ctx->device);
    cudaStream_t st = s->ctx->stream;
    cudaFreeAsync(s->y, st);
    cudaFreeAsync(s->d_prefix, st);
    cudaFreeAsync(s->d_partial, st);
    delete s;
    return PBN_OK;
}

int pbn_ucv_score(pbn_ucv* s, const double* H_or_hdiag, int is_diag, double* out) {
    if (!s || !H_or_hdiag || !out) return set_error(PBN_ERR_ARG, "null argument");
    DevSetter ds(s->ctx->device);
    return ucv_score_full(s, H_or_hdiag, is_diag, out);
}

int pbn_ucv_pair_sums(pbn_ucv* s, const double* H_or_hdiag, int is_diag, int part, int nparts, double* S2, double* S1) {
    if (!s || !H_or_hdiag || !S2 || !S1 || nparts < 1 || part < 0 || part >= nparts)
        return set_error(PBN_ERR_ARG, "invalid argument");
    DevSetter ds(s->ctx->device);
    std::vector<double> L;
    PBN_TRY(ucv_factor(s, H_or_hdiag, is_diag, L));
    return ucv_sums_devices(s, L.data(), part, nparts, S2, S1);
}

int pbn_ucv_score_from_sums(pbn_ucv* s, const double* H_or_hdiag, int is_diag, double S2, double S1, double* out) {
    if (!s || !H_or_hdiag || !out) return set_error(PBN_ERR_ARG, "invalid argument");
    std::vector<double> L;
    PBN_TRY(ucv_factor(s, H_or_hdiag, is_diag, L));
    *out = ucv_combine(s, L.data(), S2, S1);
    return PBN_OK;
}

int64_t pbn_ucv_pairs(const pbn_ucv* s) { return s ? s->n * (s->n - 1) / 2 : 0; }

int pbn_ucv_bandwidth(pbn_ctx* ctx, const pbn_table* tbl, const int* cols, int d, pbn_rows rows, int diagonal,
                      double* H_out, int* n_evals) {
    if (!ctx || !H_out) return set_error(PBN_ERR_ARG, "null argument");
    const double machine_tol = 1.4901161193847656e-08;  // util/math_constants.hpp:30
    pbn_ucv* s = nullptr;
    if (diagonal) {
        std::vector<double> h0(d);
        PBN_TRY(pbn_diag_bandwidth(ctx, tbl, cols, d, rows, PBN_BW_NORMAL_REFERENCE, h0.data()));
        PBN_TRY(pbn_ucv_create(ctx, tbl, cols, d, rows, &s));
        double start_score;
        // (the reference scores this start point with score_unconstrained on a vector, UCV.cpp:460,
        //  which is only well defined for d == 1; the diagonal score is the same quantity)
        int rc = ucv_score_full(s, h0.data(), 1, &start_score);
        if (rc != PBN_OK) { pbn_ucv_free(s); return rc; }
        double start_det = 1;
        for (int i = 0; i < d; ++i) start_det *= h0[i];
        std::vector<double> x0(d);
        for (int i = 0; i < d; ++i) x0[i] = sqrt(h0[i]);
        int err = PBN_OK;
        auto obj = [&](const double* x) -> double {  // wrap_ucv_diag_optim, UCV.cpp:408-425
            double ds = 1;
            for (int i = 0; i < d; ++i) ds *= x[i];
            double det = ds * ds;
            if (det <= machine_tol || det < 1e-3 * start_det || det > 1e3 * start_det) return start_score + 10e-8;
            std::vector<double> h(d);
            for (int i = 0; i < d; ++i) h[i] = x[i] * x[i];
            double v;
            int rc2 = ucv_score_full(s, h.data(), 1, &v);
            if (rc2 != PBN_OK) { err = rc2; return start_score + 10e-8; }
            if (fabs(v) > 1e3 * fabs(start_score)) return start_score + 10e-8;
            return v;
        };
        NMResult r = nelder_mead(obj, x0, 1e-4, 1e-4, 100000);
        if (n_evals) *n_evals = r.evals;
        pbn_ucv_free(s);
        if (err != PBN_OK && err != PBN_ERR_SINGULAR) return err;
        for (int i = 0; i < d; ++i) H_out[i] = r.x[i] * r.x[i];
        return PBN_OK;
    }
    std::vector<double> H0(d * d), L0(d * d);
    PBN_TRY(pbn_bandwidth(ctx, tbl, cols, d, rows, PBN_BW_NORMAL_REFERENCE, H0.data()));
    PBN_TRY(pbn_ucv_create(ctx, tbl, cols, d, rows, &s));
    double start_score;
    int rc = ucv_score_full(s, H0.data(), 0, &start_score);
    if (rc != PBN_OK) { pbn_ucv_free(s); return rc; }
    chol_lower(H0.data(), d, L0.data());
    double start_det = 1;
    for (int i = 0; i < d; ++i) start_det *= L0[i + i * d] * L0[i + i * d];
    // vech: lower triangle stacked column by column (util/vech_ops.cpp:6-23)
    std::vector<double> x0;
    for (int j = 0; j < d; ++j)
        for (int i = j; i < d; ++i) x0.push_back(L0[i + j * d]);
    int err = PBN_OK;
    auto unpack = [&](const double* x, std::vector<double>& S) {
        S.assign(d * d, 0.0);
        int k = 0;
        for (int j = 0; j < d; ++j)
            for (int i = j; i < d; ++i) S[i + j * d] = x[k++];
    };
    auto obj = [&](const double* x) -> double {  // wrap_ucv_optim, UCV.cpp:427-450
        std::vector<double> S;
        unpack(x, S);
        double sl = 0;
        for (int i = 0; i < d; ++i) sl += log(S[i + i * d]);
        double det = exp(2 * sl);
        if (det <= machine_tol || det < 1e-3 * start_det || det > 1e3 * start_det || std::isnan(det))
            return start_score + 10e-8;
        std::vector<double> H(d * d, 0.0);
        for (int i = 0; i < d; ++i)
            for (int j = 0; j < d; ++j) {
                double a = 0;
                for (int k = 0; k < d; ++k) a += S[i + k * d] * S[j + k * d];
                H[i + j * d] = a;
            }
        double v;
        int rc2 = ucv_score_full(s, H.data(), 0, &v);
        if (rc2 != PBN_OK) { if (rc2 != PBN_ERR_SINGULAR) err = rc2; return start_score + 10e-8; }
        if (fabs(v) > 1e3 * fabs(start_score)) return start_score + 10e-8;
        return v;
    };
    NMResult r = nelder_mead(obj, x0, 1e-4, 1e-4, 100000);
    if (n_evals) *n_evals = r.evals;
    pbn_ucv_free(s);
    if (err != PBN_OK) return err;
    std::vector<double> S;
    unpack(r.x.data(), S);
    for (int i = 0; i < d; ++i)
        for (int j = 0; j < d; ++j) {
            double a = 0;
            for (int k = 0; k < d; ++k) a += S[i + k * d] * S[j + k * d];
            H_out[i + j * d] = a;
        }
    return PBN_OK;
}

}  // extern "C"
