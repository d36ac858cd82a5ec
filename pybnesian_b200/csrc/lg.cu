// lg.cu — LinearGaussianCPD fit / logl / slogl on resident tables (pbn_lg_*).
//
// Replaces factors/continuous/LinearGaussianCPD.cpp:82-149, 251-292 and
// learning/parameters/mle_LinearGaussianCPD.hpp:11-221: the fit is O(d^2) host arithmetic on the
// device-computed mean / covariance of the selected rows (lg_fit_from_moments, cv.cu); logl is one
// streaming kernel over the (p + 1) columns (HBM-bound: 8 (p + 1) bytes read + 8 written per row).
#include <cuda_runtime.h>
#include <math.h>

#include <algorithm>
#include <string>
#include <vector>

#include "internal.h"

namespace {

struct LgParams {
    ColPtrs cols;  // variable first
    double beta[PBN_MAX_DIM];
    int p;
    double inv_std, cst;
    int64_t b0, n0, b1, n;
    double* out;       // [n] or null
    double* partial;   // [gridDim.x] block sums
};

template <typename T>
__global__ void lg_logl_kernel(const __grid_constant__ LgParams P) {
    __shared__ double sh[32];
    double s = 0;
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < P.n; r += (int64_t)gridDim.x * blockDim.x) {
        int64_t rr = r < P.n0 ? P.b0 + r : P.b1 + (r - P.n0);
        double mean = P.beta[0];
        for (int j = 1; j <= P.p; ++j) mean = fma(P.beta[j], static_cast<double>(static_cast<const T*>(P.cols.p[j])[rr]), mean);
        double z = P.inv_std * (static_cast<double>(static_cast<const T*>(P.cols.p[0])[rr]) - mean);
        double l = -0.5 * z * z + P.cst;
        if (P.out) P.out[r] = l;
        s += l;
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) sh[w] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sh[i];
        P.partial[blockIdx.x] = t;
    }
}

}  // namespace

extern "C" {

int pbn_lg_fit(pbn_ctx* ctx, const pbn_table* tbl, const int* cols, int d, pbn_rows rows, double* beta_out, double* variance_out) {
    if (!ctx || !beta_out || !variance_out) return set_error(PBN_ERR_ARG, "null argument");
    PBN_TRY(check_cols(tbl, cols, d));
    PBN_TRY(check_rows(tbl, rows));
    int64_t n = seg_count(rows);
    if (n < 1) return set_error(PBN_ERR_ARG, "cannot fit a LinearGaussianCPD with 0 instances");
    DevSetter ds(ctx->device);
    std::vector<double> mean(d), cov((size_t)d * d, 0.0);
    PBN_TRY(moments_impl(ctx, tbl, cols, d, rows, mean.data(), n >= 2 ? cov.data() : nullptr));
    for (auto& v : cov) v *= (double)(n - 1);  // back to centred sums of products
    *variance_out = lg_fit_from_moments(n, d - 1, mean.data(), cov.data(), beta_out);
    return PBN_OK;
}

int pbn_lg_logl(pbn_ctx* ctx, const pbn_table* tbl, const int* cols, int d, pbn_rows rows, const double* beta, double variance,
                double* out_logl, double* out_slogl) {
    if (!ctx || !beta) return set_error(PBN_ERR_ARG, "null argument");
    PBN_TRY(check_cols(tbl, cols, d));
    PBN_TRY(check_rows(tbl, rows));
    DevSetter ds(ctx->device);
    cudaStream_t st = ctx->stream;
    int64_t n = seg_count(rows);
    if (n == 0) {
        if (out_slogl) *out_slogl = 0.0;
        return PBN_OK;
    }
    LgParams P;
    memset(&P, 0, sizeof(P));
    for (int i = 0; i < d; ++i) {
        P.cols.p[i] = col_ptr(tbl, cols[i]);
        P.beta[i] = beta[i];
    }
    P.p = d - 1;
    P.inv_std = 1.0 / sqrt(variance);
    P.cst = -0.5 * log(variance) - 0.5 * 1.8378770664093454836;
    P.b0 = rows.b0;
    P.n0 = rows.e0 - rows.b0;
    P.b1 = rows.b1;
    P.n = n;
    int blocks = (int)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * 8);
    double* d_out = nullptr;
    double* d_part = nullptr;
    if (out_logl) PBN_CUDA_TRY(cudaMallocAsync(&d_out, (size_t)n * sizeof(double), st));
    PBN_CUDA_TRY(cudaMallocAsync(&d_part, (size_t)blocks * sizeof(double), st));
    P.out = d_out;
    P.partial = d_part;
    if (tbl->dtype == PBN_F64) lg_logl_kernel<double><<<blocks, 256, 0, st>>>(P);
    else lg_logl_kernel<float><<<blocks, 256, 0, st>>>(P);
    ctx->launches++;
    PBN_CUDA_TRY(cudaGetLastError());
    std::vector<double> part(blocks);
    PBN_CUDA_TRY(cudaMemcpyAsync(part.data(), d_part, (size_t)blocks * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (out_logl) PBN_CUDA_TRY(cudaMemcpyAsync(out_logl, d_out, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
    PBN_CUDA_TRY(cudaStreamSynchronize(st));
    ctx->d2h += (int64_t)blocks * 8 + (out_logl ? n * 8 : 0);
    if (d_out) PBN_CUDA_TRY(cudaFreeAsync(d_out, st));
    PBN_CUDA_TRY(cudaFreeAsync(d_part, st));
    if (out_slogl) {
        double s = 0;
        for (int b = 0; b < blocks; ++b) s += part[b];
        *out_slogl = s;
    }
    return PBN_OK;
}

}  // extern "C"
