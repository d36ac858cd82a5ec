// runtime.cu — libpbn_cuda.so: context, resident column store, moments / whitening /
// finalize kernels and the C ABI declared in include/pbn_cuda.h.
//
// This is the layer that sits where opencl::OpenCLConfig sits in the reference
// (opencl/opencl_config.{hpp,cpp}); the per-function reference citations are in
// include/pbn_cuda.h.  There is no CPU fallback: every compute entry point launches
// CUDA kernels and fails with PBN_ERR_CUDA if that is impossible.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "internal.h"

using pbn::PairJob;

static thread_local std::string g_last_error;

int pbn_set_error(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}

// ------------------------------------------------------------------------------------
// small host linear algebra (d <= 32)
// ------------------------------------------------------------------------------------
bool chol_lower(const double* H, int d, double* L) {  // column-major
    std::fill(L, L + d * d, 0.0);
    for (int j = 0; j < d; ++j) {
        double s = H[j + j * d];
        for (int k = 0; k < j; ++k) s -= L[j + k * d] * L[j + k * d];
        if (!(s > 0.0) || !std::isfinite(s)) return false;
        double ljj = sqrt(s);
        L[j + j * d] = ljj;
        for (int i = j + 1; i < d; ++i) {
            double t = H[i + j * d];
            for (int k = 0; k < j; ++k) t -= L[i + k * d] * L[j + k * d];
            L[i + j * d] = t / ljj;
        }
    }
    return true;
}

// inverse of a lower-triangular column-major matrix, result row-major lower
void tri_inverse_rowmajor(const double* L, int d, double* Winv) {
    std::fill(Winv, Winv + d * d, 0.0);
    for (int j = 0; j < d; ++j) {
        // solve L x = e_j
        for (int i = j; i < d; ++i) {
            double s = (i == j) ? 1.0 : 0.0;
            for (int k = j; k < i; ++k) s -= L[i + k * d] * Winv[k * d + j];
            Winv[i * d + j] = s / L[i + i * d];
        }
    }
}

static void jacobi_eigenvalues(std::vector<double> a, int d, std::vector<double>& ev) {
    for (int sweep = 0; sweep < 100; ++sweep) {
        double off = 0;
        for (int p = 0; p < d; ++p)
            for (int q = p + 1; q < d; ++q) off += a[p + q * d] * a[p + q * d];
        if (off < 1e-300) break;
        for (int p = 0; p < d; ++p)
            for (int q = p + 1; q < d; ++q) {
                double apq = a[p + q * d];
                if (apq == 0.0) continue;
                double theta = (a[q + q * d] - a[p + p * d]) / (2.0 * apq);
                double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
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

// util/basic_eigen_ops.hpp:136-147 (eps of the data type)
bool is_psd(const double* cov, int d, int dtype) {
    std::vector<double> a(cov, cov + d * d), ev;
    jacobi_eigenvalues(a, d, ev);
    double mx = *std::max_element(ev.begin(), ev.end());
    double mn = *std::min_element(ev.begin(), ev.end());
    double eps = dtype == PBN_F64 ? 2.220446049250313e-16 : 1.1920928955078125e-07;
    return !(mn < mx * d * eps);
}

// ------------------------------------------------------------------------------------
// device kernels: moments
// ------------------------------------------------------------------------------------

__device__ __forceinline__ int64_t map_row(int64_t r, int64_t b0, int64_t n0, int64_t b1) {
    return r < n0 ? b0 + r : b1 + (r - n0);
}

__device__ __forceinline__ double block_reduce_sum(double v, double* sh) {
    // deterministic: warp shuffle tree, then warp 0 adds the warp partials in order
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    double r = 0;
    if (threadIdx.x == 0) {
        int nw = (blockDim.x + 31) >> 5;
        for (int i = 0; i < nw; ++i) r += sh[i];
    }
    return r;  // valid in thread 0
}

// column sums: partial[block][d]
template <typename T>
__global__ void colsum_kernel(ColPtrs cols, int d, int64_t b0, int64_t n0, int64_t b1, int64_t n, double* partial) {
    __shared__ double sh[32];
    for (int c = 0; c < d; ++c) {
        const T* x = static_cast<const T*>(cols.p[c]);
        double s = 0;
        for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x)
            s += static_cast<double>(x[map_row(r, b0, n0, b1)]);
        double tot = block_reduce_sum(s, sh);
        if (threadIdx.x == 0) partial[blockIdx.x * d + c] = tot;
    }
}

// centred cross products for a 4x4 tile of the covariance matrix (blockIdx.y = tile id):
// partial[(tile * gridDim.x + block) * 16 + a*4+b]
template <typename T>
__global__ void cov_tile_kernel(ColPtrs cols, Vec32 mean, int d, int ntile_side, int64_t b0, int64_t n0, int64_t b1,
                                int64_t n, double* partial) {
    __shared__ double sh[32];
    int ti = blockIdx.y / ntile_side, tj = blockIdx.y % ntile_side;
    if (tj < ti) return;
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0;
    const T* xa[4];
    const T* xb[4];
    double ma[4], mb[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        int ca = ti * 4 + a, cb = tj * 4 + a;
        xa[a] = ca < d ? static_cast<const T*>(cols.p[ca]) : nullptr;
        xb[a] = cb < d ? static_cast<const T*>(cols.p[cb]) : nullptr;
        ma[a] = ca < d ? mean.v[ca] : 0;
        mb[a] = cb < d ? mean.v[cb] : 0;
    }
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
        int64_t rr = map_row(r, b0, n0, b1);
        double va[4], vb[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            va[a] = xa[a] ? static_cast<double>(xa[a][rr]) - ma[a] : 0.0;
            vb[a] = xb[a] ? static_cast<double>(xb[a][rr]) - mb[a] : 0.0;
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a * 4 + b] = fma(va[a], vb[b], acc[a * 4 + b]);
    }
    for (int i = 0; i < 16; ++i) {
        double tot = block_reduce_sum(acc[i], sh);
        if (threadIdx.x == 0) partial[((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * 16 + i] = tot;
    }
}

// ------------------------------------------------------------------------------------
// device kernels: whitening  y = W (x - mu), AoS output
// ------------------------------------------------------------------------------------
// DT = compile-time number of variables (1..8: everything stays in registers), 0 = run-time d up to PBN_MAX_DIM
template <typename T, int DT>
__global__ void whiten_kernel(const __grid_constant__ WhitenParams P, T* __restrict__ out, float* __restrict__ bound,
                              double* __restrict__ nrm) {
    // the block's rows form one contiguous chunk of the AoS output: stage them in shared memory and write the chunk
    // with coalesced stores (a thread's own d values are 8 d bytes apart from its neighbour's)
    extern __shared__ __align__(16) unsigned char whiten_smem[];
    T* tile = reinterpret_cast<T*>(whiten_smem);
    constexpr int DA = DT ? DT : PBN_MAX_DIM;
    const int64_t base = blockIdx.x * (int64_t)blockDim.x;
    int64_t r = base + threadIdx.x;
    float mx = 0.f;
    const int d = DT ? DT : P.d;
    if (r < P.n) {
        int64_t rr = map_row(r, P.b0, P.n0, P.b1);
        double x[DA];
#pragma unroll
        for (int c = 0; c < DA; ++c)
            if (c < d) x[c] = static_cast<double>(static_cast<const T*>(P.cols.p[c])[rr]) - P.mu[c];
        double nn = 0;
#pragma unroll
        for (int i = 0; i < DA; ++i) {
            if (i < d) {
                double s = 0;
#pragma unroll
                for (int k = 0; k < DA; ++k)
                    if (k <= i) s = fma(P.W[i * (i + 1) / 2 + k], x[k], s);
                tile[threadIdx.x * d + i] = static_cast<T>(s);
                if (i < P.dn) nn = fma(-s, s, nn);
                float a = fabsf(static_cast<float>(s));
                mx = (a > mx || a != a) ? (a != a ? INFINITY : a) : mx;  // NaN counts as unbounded
            }
        }
        if (nrm) nrm[r] = nn;
    }
    __syncthreads();
    {
        int64_t rows_here = P.n - base;
        if (rows_here > blockDim.x) rows_here = blockDim.x;
        const int64_t cnt = rows_here * d;
        T* dst = out + base * d;
        for (int64_t j = threadIdx.x; j < cnt; j += blockDim.x) dst[j] = tile[j];
    }
    // max |coordinate| of the launch (non-negative floats order like their bit patterns).  One atomic per warp on a
    // single address serialises in L2 (it was the whole run time of this kernel at 8M rows), so a warp only issues it
    // when its value can still raise the current maximum (the plain read may be stale: the atomic stays correct).
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (bound && (threadIdx.x & 31) == 0 && mx > 0.f) {
        const int v = __float_as_int(mx * 1.0001f);
        if (v > *reinterpret_cast<volatile int*>(bound)) atomicMax(reinterpret_cast<int*>(bound), v);
    }
}

// ------------------------------------------------------------------------------------
// device kernels: finalize / fallback / reduce
// ------------------------------------------------------------------------------------
struct FinalizeParams {
    const PairJob* job;  // single job in device memory
    long long upb;
    int tb;              // test rows per tile
    int ckde;
    double lognorm_joint, lognorm_marg;
    double u2ln;         // kernel exponent unit -> natural log
    double thresh;       // sums below this are re-evaluated with a shift
    double* out;         // [m]
    int* flagged;        // [m] row ids needing the shifted path
    int* n_flagged;
};

__global__ void finalize_kernel(FinalizeParams P) {
    const PairJob jb = *P.job;
    long long row = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (row >= jb.m) return;
    long long tt = row / P.tb;
    long long ustart = jb.unit_begin + tt * jb.n_train_tiles;
    int first = (int)(ustart / P.upb);
    int last = (int)((ustart + jb.n_train_tiles - 1) / P.upb);
    int ns = last - first + 1;
    double sj = 0, sm = 0;
    for (int s = 0; s < ns; ++s) {
        sj += jb.part[(long long)s * jb.m_pad + row];
        if (P.ckde) sm += jb.part[((long long)jb.slots + s) * jb.m_pad + row];
    }
    bool bad = !(sj >= P.thresh) || (P.ckde && !(sm >= P.thresh));
    // NaN sums (NaN inputs) are not "underflow": propagate them
    if (sj != sj || (P.ckde && sm != sm)) bad = false;
    if (bad) {
        int slot = atomicAdd(P.n_flagged, 1);
        P.flagged[slot] = (int)row;
        return;
    }
    double v = P.lognorm_joint + log(sj);
    if (P.ckde) v -= P.lognorm_marg + log(sm);
    P.out[row] = v;
}

// Robust per-row evaluation (max-shifted two-pass log-sum-exp, one CTA per test row).
// Used for rows flagged by finalize_kernel and as the generic path for d > 8.
struct RowParams {
    const void* train;   // whitened AoS [n][d]
    const void* test;    // whitened AoS [m][d]
    long long n;
    int d;
    int ckde;
    double lognorm_joint, lognorm_marg, u2ln;
    const int* rows;     // row ids (null = all rows 0..count)
    const int* count_ptr;  // device count (null = use count)
    long long count;
    double* out;
};

template <typename T>
__global__ void row_kernel(RowParams P) {
    __shared__ double sh[32];
    __shared__ double bc[2];
    __shared__ double yt[PBN_MAX_DIM];
    long long cnt = P.count_ptr ? (long long)(*P.count_ptr) : P.count;
    const T* tr = static_cast<const T*>(P.train);
    const T* te = static_cast<const T*>(P.test);
    const int d = P.d;
    for (long long f = blockIdx.x; f < cnt; f += gridDim.x) {
        long long row = P.rows ? P.rows[f] : f;
        __syncthreads();
        if (threadIdx.x < d) yt[threadIdx.x] = static_cast<double>(te[row * d + threadIdx.x]);
        __syncthreads();
        // pass 1: minima of the joint / marginal squared distances (kernel units)
        double mnj = INFINITY, mnm = INFINITY;
        for (long long i = threadIdx.x; i < P.n; i += blockDim.x) {
            double s = 0, sm = 0;
            for (int c = 0; c < d; ++c) {
                double dl = yt[c] - static_cast<double>(tr[i * d + c]);
                s = fma(dl, dl, s);
                if (c == d - 2) sm = s;
            }
            mnj = fmin(mnj, s);
            mnm = fmin(mnm, sm);
        }
        for (int o = 16; o > 0; o >>= 1) {
            mnj = fmin(mnj, __shfl_down_sync(0xffffffffu, mnj, o));
            mnm = fmin(mnm, __shfl_down_sync(0xffffffffu, mnm, o));
        }
        __syncthreads();
        if ((threadIdx.x & 31) == 0) { sh[threadIdx.x >> 5] = mnj; sh[16 + (threadIdx.x >> 5)] = mnm; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double a = INFINITY, b = INFINITY;
            for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { a = fmin(a, sh[i]); b = fmin(b, sh[16 + i]); }
            bc[0] = a; bc[1] = b;
        }
        __syncthreads();
        mnj = bc[0]; mnm = bc[1];
        // pass 2: shifted sums
        const double u2 = P.u2ln;  // exp(-(s - min) * u2ln)
        double aj = 0, am = 0;
        for (long long i = threadIdx.x; i < P.n; i += blockDim.x) {
            double s = 0, sm = 0;
            for (int c = 0; c < d; ++c) {
                double dl = yt[c] - static_cast<double>(tr[i * d + c]);
                s = fma(dl, dl, s);
                if (c == d - 2) sm = s;
            }
            aj += exp(-(s - mnj) * u2);
            if (P.ckde) am += exp(-(sm - mnm) * u2);
        }
        double tj = block_reduce_sum(aj, sh);
        double tm = 0;
        if (P.ckde) tm = block_reduce_sum(am, sh);
        if (threadIdx.x == 0) {
            double v = P.lognorm_joint + log(tj) - mnj * u2;
            if (P.ckde) v -= P.lognorm_marg + log(tm) - mnm * u2;
            P.out[row] = v;
        }
    }
}

__global__ void sum_partial_kernel(const double* __restrict__ x, long long n, double* partial) {
    __shared__ double sh[32];
    double s = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        s += x[i];
    double tot = block_reduce_sum(s, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = tot;
}
__global__ void sum_final_kernel(const double* __restrict__ partial, int n, double* out) {
    __shared__ double sh[32];
    double s = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += partial[i];
    double tot = block_reduce_sum(s, sh);
    if (threadIdx.x == 0) *out = tot;
}

__global__ void write_job_kernel(PairJob j, PairJob* dst, int* zero_counter) {
    *dst = j;
    if (zero_counter) *zero_counter = 0;
}

// ------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------
int check_cols(const pbn_table* tbl, const int* cols, int d) {
    if (!tbl || !cols) return set_error(PBN_ERR_ARG, "null table / column list");
    if (d < 1 || d > PBN_MAX_DIM) return set_error(PBN_ERR_UNSUPPORTED, "number of variables must be in [1, 32]");
    for (int i = 0; i < d; ++i)
        if (cols[i] < 0 || cols[i] >= tbl->ncols) return set_error(PBN_ERR_ARG, "column index out of range");
    return PBN_OK;
}
int check_rows(const pbn_table* tbl, const pbn_rows& r) {
    if (r.b0 < 0 || r.e0 < r.b0 || r.e0 > tbl->nrows || r.b1 < 0 || r.e1 < r.b1 || r.e1 > tbl->nrows)
        return set_error(PBN_ERR_ARG, "row range out of bounds");
    return PBN_OK;
}
const void* col_ptr(const pbn_table* t, int c) {
    return static_cast<const char*>(t->data) + (size_t)c * t->stride * elem_size(t->dtype);
}

int moments_impl(pbn_ctx* ctx, const pbn_table* tbl, const int* cols, int d, pbn_rows rows, double* mean_out,
                        double* cov_out) {
    int64_t n = seg_count(rows);
    if (n <= 0) return set_error(PBN_ERR_ARG, "empty row range");
    ColPtrs cp;
    for (int i = 0; i < d; ++i) cp.p[i] = col_ptr(tbl, cols[i]);
    int64_t n0 = rows.e0 - rows.b0;
    const int threads = 256;
    int blocks = (int)std::min<int64_t>((n + threads - 1) / threads, (int64_t)ctx->sm_count * 4);
    int side = (d + 3) / 4;
    size_t pbytes = std::max<size_t>((size_t)blocks * d, cov_out ? (size_t)side * side * blocks * 16 : 0) * sizeof(double);
    double* d_part = nullptr;
    PBN_CUDA_TRY(cudaMallocAsync(&d_part, pbytes, ctx->stream));
    std::vector<double> h((size_t)pbytes / sizeof(double));
    if (tbl->dtype == PBN_F64)
        colsum_kernel<double><<<blocks, threads, 0, ctx->stream>>>(cp, d, rows.b0, n0, rows.b1, n, d_part);
    else
        colsum_kernel<float><<<blocks, threads, 0, ctx->stream>>>(cp, d, rows.b0, n0, rows.b1, n, d_part);
    ctx->launches++;
    PBN_CUDA_TRY(cudaGetLastError());
    PBN_CUDA_TRY(cudaMemcpyAsync(h.data(), d_part, (size_t)blocks * d * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    PBN_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    ctx->d2h += (int64_t)blocks * d * 8;
    Vec32 mean;
    for (int c = 0; c < d; ++c) {
        double s = 0;
        for (int b = 0; b < blocks; ++b) s += h[(size_t)b * d + c];
        mean.v[c] = s / (double)n;
        if (mean_out) mean_out[c] = mean.v[c];
    }
    if (cov_out) {
        dim3 grid(blocks, side * side);
        if (tbl->dtype == PBN_F64)
            cov_tile_kernel<double><<<grid, threads, 0, ctx->stream>>>(cp, mean, d, side, rows.b0, n0, rows.b1, n, d_part);
        else
            cov_tile_kernel<float><<<grid, threads, 0, ctx->stream>>>(cp, mean, d, side, rows.b0, n0, rows.b1, n, d_part);
        ctx->launches++;
        PBN_CUDA_TRY(cudaGetLastError());
        size_t cb = (size_t)side * side * blocks * 16 * sizeof(double);
        PBN_CUDA_TRY(cudaMemcpyAsync(h.data(), d_part, cb, cudaMemcpyDeviceToHost, ctx->stream));
        PBN_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        ctx->d2h += (int64_t)cb;
        double inv = 1.0 / (double)(n - 1);
        for (int ti = 0; ti < side; ++ti)
            for (int tj = ti; tj < side; ++tj) {
                int tile = ti * side + tj;
                for (int a = 0; a < 4; ++a)
                    for (int b = 0; b < 4; ++b) {
                        int i = ti * 4 + a, j = tj * 4 + b;
                        if (i >= d || j >= d) continue;
                        double s = 0;
                        for (int blk = 0; blk < blocks; ++blk) s += h[((size_t)tile * blocks + blk) * 16 + a * 4 + b];
                        cov_out[i + j * d] = cov_out[j + i * d] = s * inv;
                    }
            }
    }
    PBN_CUDA_TRY(cudaFreeAsync(d_part, ctx->stream));
    return PBN_OK;
}

std::string var_list(const int* cols, int d) {
    std::string s = "[";
    for (int i = 0; i < d; ++i) s += (i ? ", " : "") + std::to_string(cols[i]);
    return s + "]";
}

int whiten_raw_launch(pbn_ctx* ctx, const pbn_table* tbl, const int* cols_io, int d, pbn_rows rows, const double* Wfull,
                      const double* mu, void* out, float* bound, double* nrm, int dn) {
    WhitenParams P;
    memset(&P, 0, sizeof(P));
    for (int i = 0; i < d; ++i) P.cols.p[i] = col_ptr(tbl, cols_io[i]);
    int w = 0;
    for (int i = 0; i < d; ++i)
        for (int j = 0; j <= i; ++j) P.W[w++] = Wfull[i * d + j];
    for (int i = 0; i < d; ++i) P.mu[i] = mu[i];
    P.d = d;
    P.dn = nrm ? dn : 0;
    P.b0 = rows.b0;
    P.n0 = rows.e0 - rows.b0;
    P.b1 = rows.b1;
    P.n = seg_count(rows);
    if (P.n == 0) return PBN_OK;
    const int threads = 256;
    int blocks = (int)((P.n + threads - 1) / threads);
    const size_t smem = (size_t)threads * d * elem_size(tbl->dtype);
    const bool f64 = tbl->dtype == PBN_F64;
#define PBN_WHITEN_CASE(DT)                                                                                              \
    case DT:                                                                                                             \
        if (f64) whiten_kernel<double, DT><<<blocks, threads, smem, ctx->stream>>>(P, static_cast<double*>(out), bound, nrm); \
        else whiten_kernel<float, DT><<<blocks, threads, smem, ctx->stream>>>(P, static_cast<float*>(out), bound, nullptr);   \
        break;
    switch (d) {
        PBN_WHITEN_CASE(1) PBN_WHITEN_CASE(2) PBN_WHITEN_CASE(3) PBN_WHITEN_CASE(4)
        PBN_WHITEN_CASE(5) PBN_WHITEN_CASE(6) PBN_WHITEN_CASE(7) PBN_WHITEN_CASE(8)
        default:
            if (smem > 48 * 1024) {  // d > 24 doubles per row: opt in to the larger carve-out
                if (f64)
                    PBN_CUDA_TRY(cudaFuncSetAttribute(whiten_kernel<double, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                else
                    PBN_CUDA_TRY(cudaFuncSetAttribute(whiten_kernel<float, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            }
            if (f64) whiten_kernel<double, 0><<<blocks, threads, smem, ctx->stream>>>(P, static_cast<double*>(out), bound, nrm);
            else whiten_kernel<float, 0><<<blocks, threads, smem, ctx->stream>>>(P, static_cast<float*>(out), bound, nullptr);
    }
#undef PBN_WHITEN_CASE
    ctx->launches++;
    PBN_CUDA_TRY(cudaGetLastError());
    return PBN_OK;
}

int pbn_whiten_kde(pbn_ctx* ctx, const pbn_kde* k, const pbn_table* tbl, const int* cols, pbn_rows rows, void* out,
                         float* bound, double* nrm) {
    int pc[PBN_MAX_DIM];
    for (int i = 0; i < k->d; ++i) pc[i] = cols[k->perm[i]];
    return whiten_raw_launch(ctx, tbl, pc, k->d, rows, k->W, k->mu, out, bound, nrm, k->ckde ? k->d - 1 : k->d);
}

double unit_scale(int dtype) {  // kernel exponent units per natural-log unit of -s/2
    const double log2e = 1.4426950408889634074;
    return dtype == PBN_F64 ? (double)pbn::kExpTab * log2e : log2e;
}

static int fit_impl(pbn_ctx* ctx, const pbn_table* tbl, const int* cols, int d, pbn_rows rows, const double* H,
                    bool ckde, pbn_kde** out) {
    if (!ctx || !out || !H) return set_error(PBN_ERR_ARG, "null argument");
    PBN_TRY(check_cols(tbl, cols, d));
    PBN_TRY(check_rows(tbl, rows));
    DevSetter ds(ctx->device);
    int64_t n = seg_count(rows);
    if (n <= 0) return set_error(PBN_ERR_ARG, "cannot fit a KDE with 0 instances");
    pbn_kde* k = new pbn_kde();
    k->ctx = ctx;
    k->d = d;
    k->dtype = tbl->dtype;
    k->n = n;
    k->ckde = ckde && d >= 2;
    k->y = nullptr;
    // internal column order: CKDE stores the variable last so that the marginal's whitened
    // coordinates are a prefix of the joint's (chol(H_perm)[:p,:p] == chol(H[1:,1:])).
    for (int i = 0; i < d; ++i) k->perm[i] = k->ckde ? (i + 1) % d : i;
    std::vector<double> Hp(d * d), L(d * d), Winv(d * d);
    for (int i = 0; i < d; ++i)
        for (int j = 0; j < d; ++j) Hp[i + j * d] = H[k->perm[i] + k->perm[j] * d];
    if (!chol_lower(Hp.data(), d, L.data())) {
        delete k;
        return set_error(PBN_ERR_SINGULAR, "bandwidth matrix is not positive definite");
    }
    tri_inverse_rowmajor(L.data(), d, Winv.data());
    double c = sqrt(0.5 * unit_scale(k->dtype));
    for (int i = 0; i < d * d; ++i) k->W[i] = c * Winv[i];
    double slog = 0, slog_m = 0;
    for (int i = 0; i < d; ++i) {
        slog += log(L[i + i * d]);
        if (i < d - 1) slog_m += log(L[i + i * d]);
    }
    const double log2pi = 1.8378770664093454836;
    k->lognorm_joint = -slog - 0.5 * d * log2pi - log((double)n);
    k->lognorm_marg = -slog_m - 0.5 * (d - 1) * log2pi - log((double)n);
    // centre on the training mean (keeps whitened coordinates small; the kernel only sees differences)
    std::vector<int> pc(d);
    for (int i = 0; i < d; ++i) pc[i] = cols[k->perm[i]];
    int rc = moments_impl(ctx, tbl, pc.data(), d, rows, k->mu, nullptr);
    if (rc != PBN_OK) { delete k; return rc; }
    int tile = k->dtype == PBN_F64 ? pbn::pair_tile_f64(d) : pbn::pair_tile_f32(d);
    int64_t n_pad = ((n + tile - 1) / tile) * tile + 16;
    size_t ybytes = ((size_t)n_pad * d * elem_size(k->dtype) + 255) / 256 * 256;
    // row norms for the dot-product form of the pair kernel (f64 fast path only)
    size_t nbytes = (k->dtype == PBN_F64 && d <= 8) ? ((size_t)n_pad * sizeof(double) + 255) / 256 * 256 : 0;
    cudaError_t e = cudaMallocAsync(&k->y, ybytes + 256 + nbytes, ctx->stream);
    if (e != cudaSuccess) { delete k; PBN_CUDA_TRY(e); }
    e = cudaMemsetAsync(k->y, 0, ybytes + 256 + nbytes, ctx->stream);
    if (e != cudaSuccess) { cudaFreeAsync(k->y, ctx->stream); delete k; PBN_CUDA_TRY(e); }
    k->d_bound = reinterpret_cast<float*>(static_cast<char*>(k->y) + ybytes);
    k->nrm = nbytes ? reinterpret_cast<double*>(static_cast<char*>(k->y) + ybytes + 256) : nullptr;
    rc = pbn_whiten_kde(ctx, k, tbl, cols, rows, k->y, k->d_bound, k->nrm);
    if (rc != PBN_OK) { cudaFreeAsync(k->y, ctx->stream); delete k; return rc; }
    *out = k;
    return PBN_OK;
}

int pbn_logl_impl(pbn_ctx* ctx, const pbn_kde* k, const pbn_table* test, const int* cols, pbn_rows rows,
                     double* d_out_logl, double* d_out_slogl, double* h_out_logl, double* h_out_slogl) {
    if (!ctx || !k) return set_error(PBN_ERR_ARG, "null argument");
    PBN_TRY(check_cols(test, cols, k->d));
    PBN_TRY(check_rows(test, rows));
    if (test->dtype != k->dtype)
        return set_error(PBN_ERR_ARG, "Data type of training and test datasets is different.");
    DevSetter ds(ctx->device);
    cudaStream_t st = ctx->stream;
    const int d = k->d;
    const int64_t m = seg_count(rows);
    ctx->last_fallback_rows = 0;
    if (m == 0) {
        if (h_out_slogl) *h_out_slogl = 0.0;
        if (d_out_slogl) PBN_CUDA_TRY(cudaMemsetAsync(d_out_slogl, 0, 8, st));
        return PBN_OK;
    }
    const bool f64 = k->dtype == PBN_F64;
    const size_t es = elem_size(k->dtype);
    const int TILE = f64 ? pbn::pair_tile_f64(d) : pbn::pair_tile_f32(d);
    const int TB = f64 ? pbn::pair_tb_for_f64(d, k->ckde) : pbn::pair_tb_for_f32(d, k->ckde);
    const bool fast = d <= 8;

    void* ytest = nullptr;
    size_t ytbytes = ((size_t)m * d * es + 255) / 256 * 256;
    const size_t tnbytes = k->nrm ? ((size_t)m * sizeof(double) + 255) / 256 * 256 : 0;
    PBN_CUDA_TRY(cudaMallocAsync(&ytest, ytbytes + 256 + tnbytes, st));
    float* bound_test = reinterpret_cast<float*>(static_cast<char*>(ytest) + ytbytes);
    double* nrm_test = tnbytes ? reinterpret_cast<double*>(static_cast<char*>(ytest) + ytbytes + 256) : nullptr;
    PBN_CUDA_TRY(cudaMemsetAsync(bound_test, 0, 256, st));
    PBN_TRY(pbn_whiten_kde(ctx, k, test, cols, rows, ytest, bound_test, nrm_test));

    double* out = d_out_logl;
    bool own_out = false;
    if (!out) {
        PBN_CUDA_TRY(cudaMallocAsync(&out, (size_t)m * sizeof(double), st));
        own_out = true;
    }
    const double u2ln = 1.0 / unit_scale(k->dtype);  // (kernel units of s'/ -t) -> natural log

    if (fast) {
        int n_test_tiles = (int)((m + TB - 1) / TB);
        int n_train_tiles = (int)((k->n + TILE - 1) / TILE);
        long long U = (long long)n_test_tiles * n_train_tiles;
        int max_grid = ctx->sm_count * 2;
        int grid = (int)std::min<long long>(U, max_grid);
        long long upb = (U + grid - 1) / grid;
        grid = (int)((U + upb - 1) / upb);
        int slots = (int)std::min<long long>((n_train_tiles + upb - 1) / upb + 1, grid);
        int n_acc = k->ckde ? 2 : 1;
        long long m_pad = (m + 31) / 32 * 32;
        double* part = nullptr;
        PairJob* d_job = nullptr;
        int* flagged = nullptr;
        PBN_CUDA_TRY(cudaMallocAsync(&part, (size_t)n_acc * slots * m_pad * sizeof(double), st));
        PBN_CUDA_TRY(cudaMallocAsync(&d_job, sizeof(PairJob) + 16, st));
        PBN_CUDA_TRY(cudaMallocAsync(&flagged, ((size_t)m + 1) * sizeof(int), st));
        int* n_flagged = flagged + m;
        PairJob job;
        memset(&job, 0, sizeof(job));
        job.train = k->y;
        job.test = ytest;
        job.part = part;
        job.bound_train = k->d_bound;
        job.bound_test = bound_test;
        job.train_nrm = k->nrm;
        job.test_nrm = nrm_test;
        job.n_train = k->n;
        job.m = m;
        job.m_pad = m_pad;
        job.unit_begin = 0;
        job.n_train_tiles = n_train_tiles;
        job.n_test_tiles = n_test_tiles;
        job.slots = slots;
        write_job_kernel<<<1, 1, 0, st>>>(job, d_job, n_flagged);
        ctx->launches++;
        PBN_CUDA_TRY(cudaGetLastError());
        cudaEvent_t ev0 = nullptr, ev1 = nullptr;
        if (ctx->timing) {
            PBN_CUDA_TRY(cudaEventCreate(&ev0));
            PBN_CUDA_TRY(cudaEventCreate(&ev1));
            PBN_CUDA_TRY(cudaEventRecord(ev0, st));
        }
        cudaError_t e = f64 ? pbn::launch_pair_f64(d, k->ckde, d_job, 1, U, upb, grid, ctx->d_exp_tab, st)
                            : pbn::launch_pair_f32(d, k->ckde, d_job, 1, U, upb, grid, ctx->d_exp_tab, st);
        ctx->launches++;
        PBN_CUDA_TRY(e);
        if (ctx->timing) {
            PBN_CUDA_TRY(cudaEventRecord(ev1, st));
            ctx->timed.emplace_back(ev0, ev1);
            ctx->pair_units += (int64_t)k->n * m * (k->ckde ? 2 : 1);
        }
        FinalizeParams F;
        F.job = d_job;
        F.upb = upb;
        F.tb = TB;
        F.ckde = k->ckde ? 1 : 0;
        F.lognorm_joint = k->lognorm_joint;
        F.lognorm_marg = k->lognorm_marg;
        F.u2ln = u2ln;
        F.thresh = f64 ? ldexp(1.0, -900) : ldexp(1.0, -64);
        F.out = out;
        F.flagged = flagged;
        F.n_flagged = n_flagged;
        finalize_kernel<<<(int)((m + 255) / 256), 256, 0, st>>>(F);
        ctx->launches++;
        PBN_CUDA_TRY(cudaGetLastError());
        // rows whose unshifted sums underflowed: exact max-shifted evaluation (count read on device)
        RowParams R;
        R.train = k->y;
        R.test = ytest;
        R.n = k->n;
        R.d = d;
        R.ckde = k->ckde ? 1 : 0;
        R.lognorm_joint = k->lognorm_joint;
        R.lognorm_marg = k->lognorm_marg;
        R.u2ln = u2ln;
        R.rows = flagged;
        R.count_ptr = n_flagged;
        R.count = 0;
        R.out = out;
        int rgrid = (int)std::min<int64_t>(m, (int64_t)ctx->sm_count * 8);
        if (f64) row_kernel<double><<<rgrid, 256, 0, st>>>(R);
        else row_kernel<float><<<rgrid, 256, 0, st>>>(R);
        ctx->launches++;
        PBN_CUDA_TRY(cudaGetLastError());
        if (h_out_logl || h_out_slogl) {
            int nf = 0;
            PBN_CUDA_TRY(cudaMemcpyAsync(&nf, n_flagged, sizeof(int), cudaMemcpyDeviceToHost, st));
            PBN_CUDA_TRY(cudaStreamSynchronize(st));
            ctx->last_fallback_rows = nf;
            ctx->d2h += 4;
        }
        PBN_CUDA_TRY(cudaFreeAsync(part, st));
        PBN_CUDA_TRY(cudaFreeAsync(d_job, st));
        PBN_CUDA_TRY(cudaFreeAsync(flagged, st));
    } else {
        RowParams R;
        R.train = k->y;
        R.test = ytest;
        R.n = k->n;
        R.d = d;
        R.ckde = k->ckde ? 1 : 0;
        R.lognorm_joint = k->lognorm_joint;
        R.lognorm_marg = k->lognorm_marg;
        R.u2ln = u2ln;
        R.rows = nullptr;
        R.count_ptr = nullptr;
        R.count = m;
        R.out = out;
        int rgrid = (int)std::min<int64_t>(m, (int64_t)ctx->sm_count * 8);
        if (f64) row_kernel<double><<<rgrid, 256, 0, st>>>(R);
        else row_kernel<float><<<rgrid, 256, 0, st>>>(R);
        ctx->launches++;
        PBN_CUDA_TRY(cudaGetLastError());
    }

    double* d_sum = d_out_slogl;
    if (h_out_slogl || d_out_slogl) {
        int sb = (int)std::min<int64_t>((m + 255) / 256, (int64_t)ctx->sm_count * 4);
        double* partial = nullptr;
        PBN_CUDA_TRY(cudaMallocAsync(&partial, (size_t)(sb + 1) * sizeof(double), st));
        if (!d_sum) d_sum = partial + sb;
        sum_partial_kernel<<<sb, 256, 0, st>>>(out, m, partial);
        sum_final_kernel<<<1, 256, 0, st>>>(partial, sb, d_sum);
        ctx->launches += 2;
        PBN_CUDA_TRY(cudaGetLastError());
        if (h_out_slogl) {
            PBN_CUDA_TRY(cudaMemcpyAsync(h_out_slogl, d_sum, sizeof(double), cudaMemcpyDeviceToHost, st));
            ctx->d2h += 8;
        }
        if (h_out_slogl) PBN_CUDA_TRY(cudaStreamSynchronize(st));
        PBN_CUDA_TRY(cudaFreeAsync(partial, st));
    }
    if (h_out_logl) {
        PBN_CUDA_TRY(cudaMemcpyAsync(h_out_logl, out, (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, st));
        PBN_CUDA_TRY(cudaStreamSynchronize(st));
        ctx->d2h += m * 8;
    }
    PBN_CUDA_TRY(cudaFreeAsync(ytest, st));
    if (own_out) PBN_CUDA_TRY(cudaFreeAsync(out, st));
    return PBN_OK;
}

// ------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------
extern "C" {

const char* pbn_last_error(void) { return g_last_error.c_str(); }
const char* pbn_version(void) { return "pybnesian_b200 0.1 (sm_100a)"; }

int pbn_device_count(int* out) {
    if (!out) return set_error(PBN_ERR_ARG, "null argument");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *out = 0;
        return set_error(PBN_ERR_CUDA, std::string("CUDA error ") + cudaGetErrorName(e) + ": no usable GPU");
    }
    *out = n;
    return PBN_OK;
}

int pbn_ctx_create(int device, pbn_ctx** out) {
    if (!out) return set_error(PBN_ERR_ARG, "null argument");
    int n = 0;
    PBN_CUDA_TRY(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n) return set_error(PBN_ERR_ARG, "invalid CUDA device index");
    DevSetter ds(device);
    pbn_ctx* c = new pbn_ctx();
    c->device = device;
    cudaDeviceProp prop;
    PBN_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        delete c;
        return set_error(PBN_ERR_UNSUPPORTED, "libpbn_cuda is built for sm_100a (Blackwell B200) only");
    }
    c->sm_count = prop.multiProcessorCount;
    PBN_CUDA_TRY(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    std::vector<double> tab(pbn::kExpTab);
    for (int j = 0; j < pbn::kExpTab; ++j) {
        // T'[j]: 2^(j/K) with (j << (20 - log2 K)) subtracted from the high word (see exp2_tab)
        double v = (double)exp2l((long double)j / pbn::kExpTab);
        uint64_t bits;
        memcpy(&bits, &v, 8);
        uint32_t hi = (uint32_t)(bits >> 32) - ((uint32_t)j << (20 - pbn::kExpTabBits));
        bits = ((uint64_t)hi << 32) | (bits & 0xffffffffull);
        memcpy(&tab[j], &bits, 8);
    }
    PBN_CUDA_TRY(cudaMalloc(&c->d_exp_tab, tab.size() * sizeof(double)));
    PBN_CUDA_TRY(cudaMemcpy(c->d_exp_tab, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice));
    // keep freed stream-ordered allocations cached instead of returning them to the OS
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t thr = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    *out = c;
    return PBN_OK;
}

int pbn_ctx_destroy(pbn_ctx* ctx) {
    if (!ctx) return PBN_OK;
    DevSetter ds(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->d_exp_tab) cudaFree(ctx->d_exp_tab);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return PBN_OK;
}

int pbn_ctx_set_stream(pbn_ctx* ctx, void* s) {
    if (!ctx) return set_error(PBN_ERR_ARG, "null context");
    ctx->stream = s ? static_cast<cudaStream_t>(s) : ctx->own_stream;
    return PBN_OK;
}
void* pbn_ctx_stream(pbn_ctx* ctx) { return ctx ? ctx->stream : nullptr; }
int pbn_ctx_synchronize(pbn_ctx* ctx) {
    if (!ctx) return set_error(PBN_ERR_ARG, "null context");
    DevSetter ds(ctx->device);
    PBN_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return PBN_OK;
}
int pbn_ctx_sm_count(pbn_ctx* ctx) { return ctx ? ctx->sm_count : 0; }
int pbn_ctx_counters(pbn_ctx* ctx, int64_t* launches, int64_t* h2d, int64_t* d2h) {
    if (!ctx) return set_error(PBN_ERR_ARG, "null context");
    if (launches) *launches = ctx->launches;
    if (h2d) *h2d = ctx->h2d;
    if (d2h) *d2h = ctx->d2h;
    return PBN_OK;
}
int pbn_ctx_set_timing(pbn_ctx* ctx, int on) {
    if (!ctx) return set_error(PBN_ERR_ARG, "null context");
    ctx->timing = on != 0;
    return PBN_OK;
}
int pbn_ctx_pair_kernel_time(pbn_ctx* ctx, double* total_ms, int64_t* n_launches, int64_t* pair_evals, int reset) {
    if (!ctx) return set_error(PBN_ERR_ARG, "null context");
    DevSetter ds(ctx->device);
    for (auto& pr : ctx->timed) {
        PBN_CUDA_TRY(cudaEventSynchronize(pr.second));
        float ms = 0;
        PBN_CUDA_TRY(cudaEventElapsedTime(&ms, pr.first, pr.second));
        ctx->pair_ms += ms;
        ctx->pair_launches++;
        cudaEventDestroy(pr.first);
        cudaEventDestroy(pr.second);
    }
    ctx->timed.clear();
    if (total_ms) *total_ms = ctx->pair_ms;
    if (n_launches) *n_launches = ctx->pair_launches;
    if (pair_evals) *pair_evals = ctx->pair_units;
    if (reset) { ctx->pair_ms = 0; ctx->pair_launches = 0; ctx->pair_units = 0; }
    return PBN_OK;
}
int pbn_ctx_last_fallback_rows(pbn_ctx* ctx, int64_t* out) {
    if (!ctx || !out) return set_error(PBN_ERR_ARG, "null argument");
    *out = ctx->last_fallback_rows;
    return PBN_OK;
}

int pbn_table_upload(pbn_ctx* ctx, const void* const* col_ptrs, int ncols, int64_t nrows, int dtype, pbn_table** out) {
    if (!ctx || !col_ptrs || !out) return set_error(PBN_ERR_ARG, "null argument");
    if (ncols < 1 || nrows < 0) return set_error(PBN_ERR_ARG, "invalid table shape");
    if (dtype != PBN_F64 && dtype != PBN_F32)
        return set_error(PBN_ERR_ARG, "Wrong data type. [double] or [float] data is expected.");
    DevSetter ds(ctx->device);
    pbn_table* t = new pbn_table();
    t->ctx = ctx;
    t->ncols = ncols;
    t->nrows = nrows;
    t->dtype = dtype;
    t->stride = (nrows + 63) / 64 * 64 + 64;
    size_t es = elem_size(dtype);
    cudaError_t e = cudaMallocAsync(&t->data, (size_t)ncols * t->stride * es, ctx->stream);
    if (e != cudaSuccess) { delete t; PBN_CUDA_TRY(e); }
    for (int c = 0; c < ncols; ++c) {
        if (nrows == 0) break;
        e = cudaMemcpyAsync(static_cast<char*>(t->data) + (size_t)c * t->stride * es, col_ptrs[c], (size_t)nrows * es,
                            cudaMemcpyHostToDevice, ctx->stream);
        if (e != cudaSuccess) { cudaFreeAsync(t->data, ctx->stream); delete t; PBN_CUDA_TRY(e); }
    }
    ctx->h2d += (int64_t)ncols * nrows * es;
    // the caller's buffers may be released right after this call returns
    PBN_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    *out = t;
    return PBN_OK;
}

int pbn_table_free(pbn_table* t) {
    if (!t) return PBN_OK;
    DevSetter ds(t->ctx->device);
    cudaFreeAsync(t->data, t->ctx->stream);
    delete t;
    return PBN_OK;
}
int64_t pbn_table_rows(const pbn_table* t) { return t ? t->nrows : 0; }
int pbn_table_cols(const pbn_table* t) { return t ? t->ncols : 0; }

int pbn_table_download(pbn_ctx* ctx, const pbn_table* tbl, int col, pbn_rows rows, void* out) {
    if (!ctx || !tbl || !out) return set_error(PBN_ERR_ARG, "null argument");
    if (col < 0 || col >= tbl->ncols) return set_error(PBN_ERR_ARG, "column index out of range");
    PBN_TRY(check_rows(tbl, rows));
    DevSetter ds(ctx->device);
    size_t es = elem_size(tbl->dtype);
    const char* src = static_cast<const char*>(col_ptr(tbl, col));
    int64_t n0 = rows.e0 - rows.b0, n1 = rows.e1 - rows.b1;
    if (n0 > 0) PBN_CUDA_TRY(cudaMemcpyAsync(out, src + rows.b0 * es, n0 * es, cudaMemcpyDeviceToHost, ctx->stream));
    if (n1 > 0)
        PBN_CUDA_TRY(cudaMemcpyAsync(static_cast<char*>(out) + n0 * es, src + rows.b1 * es, n1 * es,
                                     cudaMemcpyDeviceToHost, ctx->stream));
    PBN_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    ctx->d2h += (n0 + n1) * es;
    return PBN_OK;
}

int pbn_table_moments(pbn_ctx* ctx, const pbn_table* tbl, const int* cols, int d, pbn_rows rows, double* mean_out,
                      double* cov_out) {
    if (!ctx) return set_error(PBN_ERR_ARG, "null context");
    PBN_TRY(check_cols(tbl, cols, d));
    PBN_TRY(check_rows(tbl, rows));
    DevSetter ds(ctx->device);
    if (cov_out && seg_count(rows) < 2) return set_error(PBN_ERR_ARG, "covariance needs at least 2 rows");
    return moments_impl(ctx, tbl, cols, d, rows, mean_out, cov_out);
}

int pbn_bandwidth(pbn_ctx* ctx, const pbn_table* tbl, const int* cols, int d, pbn_rows rows, int rule, double* H_out) {
    if (!ctx || !H_out) return set_error(PBN_ERR_ARG, "null argument");
    PBN_TRY(check_cols(tbl, cols, d));
    PBN_TRY(check_rows(tbl, rows));
    int64_t n = seg_count(rows);
    if (n <= d)
        return set_error(PBN_ERR_SINGULAR, "Bandwidth matrix of " + std::to_string(d) + " variables " + var_list(cols, d) +
                                               " cannot be estimated with " + std::to_string(n) + " instances");
    DevSetter ds(ctx->device);
    std::vector<double> cov(d * d), mean(d);
    PBN_TRY(moments_impl(ctx, tbl, cols, d, rows, mean.data(), cov.data()));
    if (tbl->dtype == PBN_F32)
        for (auto& v : cov) v = (double)(float)v;  // the reference holds the covariance in the data type
    if (!is_psd(cov.data(), d, tbl->dtype))
        return set_error(PBN_ERR_SINGULAR, "Covariance matrix for variables " + var_list(cols, d) + " is not positive-definite.");
    double N = (double)n, dd = (double)d, kfac;
    if (tbl->dtype == PBN_F32) N = (double)(float)N;
    if (rule == PBN_BW_NORMAL_REFERENCE)
        kfac = pow(4. / (N * (dd + 2.)), 2. / (dd + 4.));
    else if (rule == PBN_BW_SCOTT)
        kfac = pow(N, -2. / (dd + 4.));
    else
        return set_error(PBN_ERR_ARG, "unknown bandwidth rule");
    for (int i = 0; i < d * d; ++i) H_out[i] = kfac * cov[i];
    return PBN_OK;
}

int pbn_diag_bandwidth(pbn_ctx* ctx, const pbn_table* tbl, const int* cols, int d, pbn_rows rows, int rule, double* h_out) {
    if (!ctx || !h_out) return set_error(PBN_ERR_ARG, "null argument");
    PBN_TRY(check_cols(tbl, cols, d));
    PBN_TRY(check_rows(tbl, rows));
    int64_t n = seg_count(rows);
    int64_t need = rule == PBN_BW_SCOTT ? 1 : d;
    if (n <= need)
        return set_error(PBN_ERR_SINGULAR, "Diagonal bandwidth matrix of " + std::to_string(d) + " variables " +
                                               var_list(cols, d) + " cannot be estimated with " + std::to_string(n) +
                                               " instances");
    DevSetter ds(ctx->device);
    std::vector<double> cov(d * d), mean(d);
    PBN_TRY(moments_impl(ctx, tbl, cols, d, rows, mean.data(), cov.data()));
    if (tbl->dtype == PBN_F32)
        for (auto& v : cov) v = (double)(float)v;
    double N = (double)n, dd = (double)d;
    if (rule == PBN_BW_SCOTT) {
        double kfac = pow(N, -2. / (dd + 4.));
        for (int i = 0; i < d; ++i) h_out[i] = kfac * cov[i + i * d];
        return PBN_OK;
    }
    if (!is_psd(cov.data(), d, tbl->dtype))
        return set_error(PBN_ERR_SINGULAR, "Covariance matrix for variables " + var_list(cols, d) + " is not positive-definite.");
    // NormalReferenceRule::diag_bandwidth, eq. (3.4) of Chacon & Duong (2018): kde/NormalReferenceRule.hpp:88-105
    std::vector<double> delta(d * d), L(d * d), Li(d * d), dinv(d * d, 0.0);
    for (int i = 0; i < d; ++i)
        for (int j = 0; j < d; ++j) delta[i + j * d] = cov[i + j * d] / cov[i + i * d];
    // delta is not symmetric: invert with Gauss-Jordan (partial pivoting)
    std::vector<double> A(delta), Inv(d * d, 0.0);
    for (int i = 0; i < d; ++i) Inv[i + i * d] = 1.0;
    double det = 1.0;
    for (int c = 0; c < d; ++c) {
        int piv = c;
        for (int r = c + 1; r < d; ++r)
            if (fabs(A[r + c * d]) > fabs(A[piv + c * d])) piv = r;
        if (A[piv + c * d] == 0.0) return set_error(PBN_ERR_SINGULAR, "singular correlation structure");
        if (piv != c) {
            for (int j = 0; j < d; ++j) {
                std::swap(A[c + j * d], A[piv + j * d]);
                std::swap(Inv[c + j * d], Inv[piv + j * d]);
            }
            det = -det;
        }
        double pv = A[c + c * d];
        det *= pv;
        for (int j = 0; j < d; ++j) { A[c + j * d] /= pv; Inv[c + j * d] /= pv; }
        for (int r = 0; r < d; ++r) {
            if (r == c) continue;
            double f = A[r + c * d];
            if (f == 0.0) continue;
            for (int j = 0; j < d; ++j) { A[r + j * d] -= f * A[c + j * d]; Inv[r + j * d] -= f * Inv[c + j * d]; }
        }
    }
    double tr = 0, tr2 = 0;
    for (int i = 0; i < d; ++i) {
        tr += Inv[i + i * d];
        for (int j = 0; j < d; ++j) tr2 += Inv[i + j * d] * Inv[j + i * d];
    }
    double kk = 4 * dd * sqrt(det) / (2 * tr2 + tr * tr);
    double f = pow(kk / N, 2. / (dd + 4.));
    for (int i = 0; i < d; ++i) h_out[i] = f * cov[i + i * d];
    (void)L; (void)Li; (void)dinv;
    return PBN_OK;
}

int pbn_kde_fit(pbn_ctx* ctx, const pbn_table* tbl, const int* cols, int d, pbn_rows rows, const double* H, pbn_kde** out) {
    return fit_impl(ctx, tbl, cols, d, rows, H, false, out);
}
int pbn_ckde_fit(pbn_ctx* ctx, const pbn_table* tbl, const int* cols, int d, pbn_rows rows, const double* H, pbn_kde** out) {
    return fit_impl(ctx, tbl, cols, d, rows, H, true, out);
}
int pbn_product_kde_fit(pbn_ctx* ctx, const pbn_table* tbl, const int* cols, int d, pbn_rows rows, const double* h,
                        pbn_kde** out) {
    if (!h || d <= 0 || d > PBN_MAX_DIM) return set_error(PBN_ERR_ARG, "invalid diagonal bandwidth");
    std::vector<double> H((size_t)d * d, 0.0);
    for (int i = 0; i < d; ++i) {
        if (!(h[i] > 0.0) || !std::isfinite(h[i]))
            return set_error(PBN_ERR_SINGULAR, "diagonal bandwidth entries must be positive");
        H[i + (size_t)i * d] = h[i];
    }
    return fit_impl(ctx, tbl, cols, d, rows, H.data(), false, out);
}
int pbn_kde_free(pbn_kde* k) {
    if (!k) return PBN_OK;
    DevSetter ds(k->ctx->device);
    if (k->y) cudaFreeAsync(k->y, k->ctx->stream);
    delete k;
    return PBN_OK;
}
int64_t pbn_kde_num_instances(const pbn_kde* k) { return k ? k->n : 0; }
double pbn_kde_lognorm(const pbn_kde* k) { return k ? k->lognorm_joint : 0.0; }

int pbn_kde_logl(pbn_ctx* ctx, const pbn_kde* kde, const pbn_table* test, const int* cols, pbn_rows rows, double* out_logl,
                 double* out_slogl) {
    return pbn_logl_impl(ctx, kde, test, cols, rows, nullptr, nullptr, out_logl, out_slogl);
}
int pbn_kde_logl_device(pbn_ctx* ctx, const pbn_kde* kde, const pbn_table* test, const int* cols, pbn_rows rows,
                        double* d_out_logl, double* d_out_slogl) {
    return pbn_logl_impl(ctx, kde, test, cols, rows, d_out_logl, d_out_slogl, nullptr, nullptr);
}

int pbn_device_alloc(pbn_ctx* ctx, int64_t bytes, void** out) {
    if (!ctx || !out || bytes < 0) return set_error(PBN_ERR_ARG, "invalid argument");
    DevSetter ds(ctx->device);
    PBN_CUDA_TRY(cudaMallocAsync(out, (size_t)std::max<int64_t>(bytes, 8), ctx->stream));
    return PBN_OK;
}
int pbn_device_free(pbn_ctx* ctx, void* p) {
    if (!ctx) return set_error(PBN_ERR_ARG, "null context");
    DevSetter ds(ctx->device);
    if (p) PBN_CUDA_TRY(cudaFreeAsync(p, ctx->stream));
    return PBN_OK;
}
int pbn_device_read(pbn_ctx* ctx, const void* dptr, int64_t bytes, void* host_out) {
    if (!ctx || !dptr || !host_out) return set_error(PBN_ERR_ARG, "null argument");
    DevSetter ds(ctx->device);
    PBN_CUDA_TRY(cudaMemcpyAsync(host_out, dptr, (size_t)bytes, cudaMemcpyDeviceToHost, ctx->stream));
    PBN_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    ctx->d2h += bytes;
    return PBN_OK;
}

}  // extern "C"
