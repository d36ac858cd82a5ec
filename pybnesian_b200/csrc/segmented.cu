// segmented.cu — row-subset ("slice") evaluation for hybrid factors: pbn_table_take, pbn_kde_logl_multi.
//
// The reference's DiscreteAdaptator (factors/discrete/DiscreteAdaptator.hpp:201-325: HCKDE, CLinearGaussianCPD)
// keeps one base factor per configuration of the discrete parents and, for every fit / logl / slogl call,
// builds one Arrow `Take` of the frame per configuration and runs the base factor on it — for a CKDE that is
// 2 uploads + 4 launches per test row per configuration.  Here the frame is gathered ONCE on the device into
// configuration-major order (pbn_table_take), every configuration is then a contiguous row range of that
// table, and the fitted per-configuration KDEs are evaluated against their ranges by ONE whitening launch
// and ONE multi-job launch of the pair kernel (pair_kernel.cuh), exactly like the (candidate, fold) jobs of
// a score batch in cv.cu.
#include <cuda_runtime.h>
#include <math.h>

#include <algorithm>
#include <limits>
#include <string>
#include <vector>

#include "internal.h"

using pbn::PairJob;

namespace {

#include "batch_kernels.cuh"

// out[c][r] = in[c][idx[r]]  (grid.y = column)
template <typename T>
__global__ void take_rows_kernel(const T* __restrict__ in, int64_t in_stride, const int32_t* __restrict__ idx, int64_t n,
                                 T* __restrict__ out, int64_t out_stride) {
    const T* src = in + (int64_t)blockIdx.y * in_stride;
    T* dst = out + (int64_t)blockIdx.y * out_stride;
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x)
        dst[r] = src[idx[r]];
}

struct SegWhitenJob {
    const void* cols[kMaxFast];               // column base pointers of the test table, internal variable order
    double W[kMaxFast * (kMaxFast + 1) / 2];  // packed lower triangle, row-major, includes the unit scale
    double mu[kMaxFast];
    long long b0, n0, b1, n;  // rows [b0, b0 + n0) ++ [b1, b1 + n - n0)
    void* y;                  // AoS [n][D]
    float* bound;             // max |coordinate|
    double* nrm;              // f64 only (else null): -sum_{c<dn} y_c^2 per row
    int dn;
};

template <typename T, int D>
__global__ void whiten_seg_kernel(const SegWhitenJob* __restrict__ jobs) {
    const SegWhitenJob& jb = jobs[blockIdx.y];
    float mx = 0.f;
    for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < jb.n; r += (long long)gridDim.x * blockDim.x) {
        const long long src = r < jb.n0 ? jb.b0 + r : jb.b1 + (r - jb.n0);
        double x[D];
#pragma unroll
        for (int c = 0; c < D; ++c) x[c] = static_cast<double>(static_cast<const T*>(jb.cols[c])[src]) - jb.mu[c];
        T* out = static_cast<T*>(jb.y) + r * D;
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
        if (sizeof(T) == 8 && jb.nrm) jb.nrm[r] = nn;
    }
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    // non-negative floats order like their bit patterns
    if ((threadIdx.x & 31) == 0 && mx > 0.f) {  // issued only when it can still raise the maximum (runtime.cu: whiten_kernel)
        const int v = __float_as_int(mx * 1.0001f);
        if (v > *reinterpret_cast<volatile int*>(jb.bound)) atomicMax(reinterpret_cast<int*>(jb.bound), v);
    }
}

template <typename T>
cudaError_t launch_whiten_seg(int d, const SegWhitenJob* jobs, int n_jobs, long long max_n, int sm_count, cudaStream_t st) {
    unsigned bx = (unsigned)std::max<long long>(1, std::min<long long>((max_n + 255) / 256, (long long)sm_count * 8));
    dim3 grid(bx, (unsigned)n_jobs);
    switch (d) {
        case 1: whiten_seg_kernel<T, 1><<<grid, 256, 0, st>>>(jobs); break;
        case 2: whiten_seg_kernel<T, 2><<<grid, 256, 0, st>>>(jobs); break;
        case 3: whiten_seg_kernel<T, 3><<<grid, 256, 0, st>>>(jobs); break;
        case 4: whiten_seg_kernel<T, 4><<<grid, 256, 0, st>>>(jobs); break;
        case 5: whiten_seg_kernel<T, 5><<<grid, 256, 0, st>>>(jobs); break;
        case 6: whiten_seg_kernel<T, 6><<<grid, 256, 0, st>>>(jobs); break;
        case 7: whiten_seg_kernel<T, 7><<<grid, 256, 0, st>>>(jobs); break;
        case 8: whiten_seg_kernel<T, 8><<<grid, 256, 0, st>>>(jobs); break;
        case 9: whiten_seg_kernel<T, 9><<<grid, 256, 0, st>>>(jobs); break;
        case 10: whiten_seg_kernel<T, 10><<<grid, 256, 0, st>>>(jobs); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

__global__ void fill_nan_kernel(double* __restrict__ out, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = __longlong_as_double(0x7ff8000000000000LL);
}

}  // namespace

extern "C" {

// DataFrame::take (dataset/dataset.hpp: arrow::compute::Take of every column) on a resident table.
static int table_take_one(pbn_ctx* ctx, const pbn_table* tbl, const int32_t* indices, int64_t n, pbn_table** out);

int pbn_table_take(pbn_ctx* ctx, const pbn_table* tbl, const int32_t* indices, int64_t n, pbn_table** out) {
    if (!ctx || !tbl || !out || (n > 0 && !indices)) return set_error(PBN_ERR_ARG, "null argument");
    if (!pbn_replicated(ctx, tbl)) return table_take_one(ctx, tbl, indices, n, out);
    const int nd = pbn_num_devices(ctx);  // multi-device context: every replica is gathered on its own device
    std::vector<pbn_table*> t(nd, nullptr);
    int rc = pbn_run_on_devices(nd, [&](int i) {
        return table_take_one(pbn_device_ctx(ctx, i), pbn_replica(const_cast<pbn_table*>(tbl), i), indices, n, &t[i]);
    });
    if (rc != PBN_OK) {
        for (pbn_table* q : t)
            if (q) pbn_table_free(q);
        return rc;
    }
    t[0]->rep.assign(t.begin() + 1, t.end());
    *out = t[0];
    return PBN_OK;
}

static int table_take_one(pbn_ctx* ctx, const pbn_table* tbl, const int32_t* indices, int64_t n, pbn_table** out) {
    if (!ctx || !tbl || !out || (n > 0 && !indices)) return set_error(PBN_ERR_ARG, "null argument");
    if (n < 0) return set_error(PBN_ERR_ARG, "negative row count");
    for (int64_t i = 0; i < n; ++i)
        if (indices[i] < 0 || indices[i] >= tbl->nrows) return set_error(PBN_ERR_ARG, "take index out of range");
    DevSetter ds(ctx->device);
    cudaStream_t st = ctx->stream;
    pbn_table* t = new pbn_table();
    t->ctx = ctx;
    t->ncols = tbl->ncols;
    t->nrows = n;
    t->dtype = tbl->dtype;
    t->stride = (n + 63) / 64 * 64 + 64;
    const size_t es = elem_size(tbl->dtype);
    cudaError_t e = cudaMallocAsync(&t->data, (size_t)t->ncols * t->stride * es, st);
    if (e != cudaSuccess) { delete t; PBN_CUDA_TRY(e); }
    if (n > 0) {
        int32_t* d_idx = nullptr;
        e = cudaMallocAsync(&d_idx, (size_t)n * sizeof(int32_t), st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_idx, indices, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) { cudaFreeAsync(t->data, st); delete t; PBN_CUDA_TRY(e); }
        ctx->h2d += n * 4;
        dim3 grid((unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * 8), (unsigned)t->ncols);
        if (tbl->dtype == PBN_F64)
            take_rows_kernel<double><<<grid, 256, 0, st>>>(static_cast<const double*>(tbl->data), tbl->stride, d_idx, n,
                                                           static_cast<double*>(t->data), t->stride);
        else
            take_rows_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(tbl->data), tbl->stride, d_idx, n,
                                                          static_cast<float*>(t->data), t->stride);
        ctx->launches++;
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);  // `indices` may be released by the caller
        cudaFreeAsync(d_idx, st);
        if (e != cudaSuccess) { cudaFreeAsync(t->data, st); delete t; PBN_CUDA_TRY(e); }
    }
    *out = t;
    return PBN_OK;
}

int pbn_kde_logl_multi(pbn_ctx* ctx, const pbn_kde* const* kdes, int n_jobs, const pbn_table* test, const int* cols,
                       const pbn_rows* rows, double* out_logl, double* out_slogl) {
    if (!ctx || !kdes || !test || !cols || !rows) return set_error(PBN_ERR_ARG, "null argument");
    if (n_jobs < 0) return set_error(PBN_ERR_ARG, "negative job count");
    const pbn_kde* first = nullptr;
    std::vector<long long> off(n_jobs + 1, 0);
    for (int j = 0; j < n_jobs; ++j) {
        PBN_TRY(check_rows(test, rows[j]));
        off[j + 1] = off[j] + seg_count(rows[j]);
        if (out_slogl) out_slogl[j] = 0.0;
        const pbn_kde* k = kdes[j];
        if (!k) continue;
        if (!first) first = k;
        if (k->d != first->d || k->dtype != first->dtype || k->ckde != first->ckde)
            return set_error(PBN_ERR_ARG, "all KDEs of a multi-job evaluation must share dimension, data type and kind");
    }
    const long long total = off[n_jobs];
    const double qnan = std::numeric_limits<double>::quiet_NaN();
    ctx->last_fallback_rows = 0;
    if (!first) {
        if (out_logl) std::fill(out_logl, out_logl + total, qnan);
        return PBN_OK;
    }
    PBN_TRY(check_cols(test, cols, first->d));
    if (test->dtype != first->dtype)
        return set_error(PBN_ERR_ARG, "Data type of training and test datasets is different.");
    const int d = first->d;
    if (d > kMaxFast) {
        // wide families: one generic evaluation per slice (row_kernel in runtime.cu)
        for (int j = 0; j < n_jobs; ++j) {
            const long long m = off[j + 1] - off[j];
            if (!kdes[j]) {
                if (out_logl) std::fill(out_logl + off[j], out_logl + off[j + 1], qnan);
                continue;
            }
            if (m == 0) continue;
            PBN_TRY(pbn_logl_impl(ctx, kdes[j], test, cols, rows[j], nullptr, nullptr, out_logl ? out_logl + off[j] : nullptr,
                                  out_slogl ? out_slogl + j : nullptr));
        }
        return PBN_OK;
    }

    DevSetter ds(ctx->device);
    cudaStream_t st = ctx->stream;
    const bool f64 = first->dtype == PBN_F64;
    const bool ckde = first->ckde;
    const size_t es = elem_size(first->dtype);
    const int TILE = f64 ? pbn::pair_tile_f64(d) : pbn::pair_tile_f32(d);
    const int TB = f64 ? pbn::pair_tb_for_f64(d, ckde) : pbn::pair_tb_for_f32(d, ckde);
    const double unit = unit_scale(first->dtype);

    std::vector<int> live;  // jobs with a fitted KDE and at least one row
    for (int j = 0; j < n_jobs; ++j)
        if (kdes[j] && off[j + 1] > off[j]) live.push_back(j);
    const int J = (int)live.size();

    double* d_out = nullptr;
    if (total > 0) {
        PBN_CUDA_TRY(cudaMallocAsync(&d_out, (size_t)total * sizeof(double), st));
        if (J < n_jobs) {  // slices without a fitted factor evaluate to NaN (DiscreteAdaptator.hpp:273-277)
            fill_nan_kernel<<<(int)std::min<long long>((total + 255) / 256, (long long)ctx->sm_count * 8), 256, 0, st>>>(d_out, total);
            ctx->launches++;
            PBN_CUDA_TRY(cudaGetLastError());
        }
    }
    std::vector<double> sums(std::max(J, 1), 0.0);
    if (J > 0) {
        std::vector<SegWhitenJob> wj(J);
        std::vector<PairJob> pj(J);
        std::vector<FinJob> fj(J);
        std::vector<size_t> y_off(J), nrm_off(J), part_off(J);
        size_t ybytes = 0;
        long long U = 0, max_m = 0;
        for (int q = 0; q < J; ++q) {
            const int j = live[q];
            const pbn_kde* k = kdes[j];
            const long long m = off[j + 1] - off[j];
            memset(&wj[q], 0, sizeof(SegWhitenJob));
            memset(&pj[q], 0, sizeof(PairJob));
            for (int i = 0; i < d; ++i) wj[q].cols[i] = col_ptr(test, cols[k->perm[i]]);
            int w = 0;
            for (int i = 0; i < d; ++i)
                for (int c = 0; c <= i; ++c) wj[q].W[w++] = k->W[i * d + c];
            for (int i = 0; i < d; ++i) wj[q].mu[i] = k->mu[i];
            wj[q].b0 = rows[j].b0;
            wj[q].n0 = rows[j].e0 - rows[j].b0;
            wj[q].b1 = rows[j].b1;
            wj[q].n = m;
            wj[q].dn = ckde ? d - 1 : d;
            y_off[q] = ybytes;
            ybytes += ((size_t)m * d * es + 16 + 255) / 256 * 256;
            nrm_off[q] = ybytes;
            if (k->nrm) ybytes += ((size_t)m * 8 + 255) / 256 * 256;
            pj[q].n_train = k->n;
            pj[q].m = m;
            pj[q].n_train_tiles = (int)((k->n + TILE - 1) / TILE);
            pj[q].n_test_tiles = (int)((m + TB - 1) / TB);
            pj[q].unit_begin = U;
            U += (long long)pj[q].n_train_tiles * pj[q].n_test_tiles;
            max_m = std::max(max_m, m);
            fj[q].lognorm_joint = k->lognorm_joint;
            fj[q].lognorm_marg = k->lognorm_marg;
            fj[q].out_off = off[j];
        }
        int grid = (int)std::min<long long>(U, (long long)ctx->sm_count * (f64 ? pbn::pair_ctas_per_sm_f64() : pbn::pair_ctas_per_sm_f32()));
        long long upb = (U + grid - 1) / grid;
        grid = (int)((U + upb - 1) / upb);
        const int n_acc = ckde ? 2 : 1;
        size_t part_elems = 0;
        for (int q = 0; q < J; ++q) {
            pj[q].slots = (int)std::min<long long>((pj[q].n_train_tiles + upb - 1) / upb + 1, grid);
            pj[q].m_pad = (pj[q].m + 31) / 32 * 32;
            part_off[q] = part_elems;
            part_elems += (size_t)n_acc * pj[q].slots * pj[q].m_pad;
        }
        size_t o = 0;
        auto carve = [&](size_t bytes) {
            size_t at = o;
            o += (bytes + 255) / 256 * 256;
            return at;
        };
        const size_t o_y = carve(ybytes), o_bound = carve((size_t)J * sizeof(float)), o_wj = carve((size_t)J * sizeof(SegWhitenJob));
        const size_t o_pj = carve((size_t)J * sizeof(PairJob)), o_fj = carve((size_t)J * sizeof(FinJob));
        const size_t o_part = carve(part_elems * sizeof(double)), o_flag = carve((size_t)total * sizeof(int2));
        const size_t o_nflag = carve(256), o_sums = carve((size_t)J * sizeof(double));
        char* base = nullptr;
        PBN_CUDA_TRY(cudaMallocAsync(&base, o, st));
        for (int q = 0; q < J; ++q) {
            const pbn_kde* k = kdes[live[q]];
            wj[q].y = base + o_y + y_off[q];
            wj[q].bound = reinterpret_cast<float*>(base + o_bound) + q;
            wj[q].nrm = k->nrm ? reinterpret_cast<double*>(base + o_y + nrm_off[q]) : nullptr;
            pj[q].train = k->y;
            pj[q].test = wj[q].y;
            pj[q].bound_train = k->d_bound;
            pj[q].bound_test = wj[q].bound;
            pj[q].train_nrm = k->nrm;
            pj[q].test_nrm = wj[q].nrm;
            pj[q].part = reinterpret_cast<double*>(base + o_part) + part_off[q];
        }
        PBN_CUDA_TRY(cudaMemsetAsync(base + o_bound, 0, (size_t)J * sizeof(float), st));
        PBN_CUDA_TRY(cudaMemsetAsync(base + o_nflag, 0, 256, st));
        PBN_CUDA_TRY(cudaMemcpyAsync(base + o_wj, wj.data(), (size_t)J * sizeof(SegWhitenJob), cudaMemcpyHostToDevice, st));
        PBN_CUDA_TRY(cudaMemcpyAsync(base + o_pj, pj.data(), (size_t)J * sizeof(PairJob), cudaMemcpyHostToDevice, st));
        PBN_CUDA_TRY(cudaMemcpyAsync(base + o_fj, fj.data(), (size_t)J * sizeof(FinJob), cudaMemcpyHostToDevice, st));
        ctx->h2d += (int64_t)J * (sizeof(SegWhitenJob) + sizeof(PairJob) + sizeof(FinJob));
        const SegWhitenJob* d_wj = reinterpret_cast<const SegWhitenJob*>(base + o_wj);
        const PairJob* d_pj = reinterpret_cast<const PairJob*>(base + o_pj);
        const FinJob* d_fj = reinterpret_cast<const FinJob*>(base + o_fj);
        int2* d_flag = reinterpret_cast<int2*>(base + o_flag);
        int* d_nflag = reinterpret_cast<int*>(base + o_nflag);
        double* d_sums = reinterpret_cast<double*>(base + o_sums);

        PBN_CUDA_TRY(f64 ? launch_whiten_seg<double>(d, d_wj, J, max_m, ctx->sm_count, st)
                         : launch_whiten_seg<float>(d, d_wj, J, max_m, ctx->sm_count, st));
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
            for (int q = 0; q < J; ++q) ctx->pair_units += pj[q].n_train * pj[q].m * (ckde ? 2 : 1);
        }
        const double thresh = f64 ? ldexp(1.0, -900) : ldexp(1.0, -64);
        dim3 fgrid((unsigned)std::min<long long>((max_m + 255) / 256, 64), (unsigned)J);
        finalize_batch_kernel<<<fgrid, 256, 0, st>>>(d_pj, d_fj, upb, TB, ckde ? 1 : 0, thresh, d_out, d_flag, d_nflag);
        ctx->launches++;
        PBN_CUDA_TRY(cudaGetLastError());
        const int rgrid = ctx->sm_count * 4;
        if (f64) row_batch_kernel<double><<<rgrid, 256, 0, st>>>(d_pj, d_fj, d, ckde ? 1 : 0, 1.0 / unit, d_flag, d_nflag, d_out);
        else row_batch_kernel<float><<<rgrid, 256, 0, st>>>(d_pj, d_fj, d, ckde ? 1 : 0, 1.0 / unit, d_flag, d_nflag, d_out);
        ctx->launches++;
        PBN_CUDA_TRY(cudaGetLastError());
        if (out_slogl) {
            segsum_kernel<<<J, 256, 0, st>>>(d_pj, d_fj, d_out, d_sums);
            ctx->launches++;
            PBN_CUDA_TRY(cudaGetLastError());
            PBN_CUDA_TRY(cudaMemcpyAsync(sums.data(), d_sums, (size_t)J * sizeof(double), cudaMemcpyDeviceToHost, st));
            ctx->d2h += (int64_t)J * 8;
        }
        int nflag = 0;
        PBN_CUDA_TRY(cudaMemcpyAsync(&nflag, d_nflag, sizeof(int), cudaMemcpyDeviceToHost, st));
        ctx->d2h += 4;
        if (out_logl) {
            PBN_CUDA_TRY(cudaMemcpyAsync(out_logl, d_out, (size_t)total * sizeof(double), cudaMemcpyDeviceToHost, st));
            ctx->d2h += total * 8;
        }
        PBN_CUDA_TRY(cudaStreamSynchronize(st));
        ctx->last_fallback_rows = nflag;
        PBN_CUDA_TRY(cudaFreeAsync(base, st));
        if (out_slogl)
            for (int q = 0; q < J; ++q) out_slogl[live[q]] = sums[q];
    } else if (out_logl && total > 0) {
        PBN_CUDA_TRY(cudaMemcpyAsync(out_logl, d_out, (size_t)total * sizeof(double), cudaMemcpyDeviceToHost, st));
        PBN_CUDA_TRY(cudaStreamSynchronize(st));
        ctx->d2h += total * 8;
    }
    if (d_out) PBN_CUDA_TRY(cudaFreeAsync(d_out, st));
    return PBN_OK;
}

}  // extern "C"
