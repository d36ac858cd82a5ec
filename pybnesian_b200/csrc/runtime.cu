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
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "internal.h"

using pbn::PairJob;

static thread_local std::string g_last_error;

int pbn_set_error(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}

// ------------------------------------------------------------------------------------
// multi-device plumbing
// ------------------------------------------------------------------------------------
int pbn_run_on_devices(int n, const std::function<int(int)>& fn) {
    if (n <= 1) return fn(0);
    std::vector<int> rc(n, PBN_OK);
    std::vector<std::string> msg(n);
    std::vector<std::thread> th;
    th.reserve(n - 1);
    for (int i = 1; i < n; ++i)
        th.emplace_back([&, i] {
            rc[i] = fn(i);
            if (rc[i] != PBN_OK) msg[i] = pbn_last_error();  // thread-local in the worker
        });
    rc[0] = fn(0);
    if (rc[0] != PBN_OK) msg[0] = pbn_last_error();
    for (auto& t : th) t.join();
    for (int i = 0; i < n; ++i)
        if (rc[i] != PBN_OK) return set_error(rc[i], msg[i]);
    return PBN_OK;
}

int pbn_morton_dims(const pbn_kde* k) {
    static const bool joint = [] { const char* e = getenv("PBN_MORTON_JOINT"); return e && e[0] == '1'; }();
    return (k->ckde && !joint) ? k->d - 1 : k->d;
}

pbn_rows pbn_sub_rows(const pbn_rows& r, int64_t begin, int64_t end) {
    const int64_t n0 = r.e0 - r.b0;
    pbn_rows o = {0, 0, 0, 0};
    if (begin < n0) {
        o.b0 = r.b0 + begin;
        o.e0 = r.b0 + std::min(end, n0);
    }
    if (end > n0) {
        const int64_t b = std::max<int64_t>(begin, n0) - n0, e = end - n0;
        if (o.e0 > o.b0) {
            o.b1 = r.b1 + b;
            o.e1 = r.b1 + e;
        } else {
            o.b0 = r.b1 + b;
            o.e0 = r.b1 + e;
        }
    }
    return o;
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
    const PairJob* jobA; // tile skipping: the pass-A job (nearest training tile of every test tile) whose sums are added, or null
    long long upbA;
    int tb;              // test rows per tile
    int ckde;
    double lognorm_joint, lognorm_marg;
    double u2ln;         // kernel exponent unit -> natural log
    double thresh;       // sums below this are re-evaluated with a shift
    double* out;         // [m]
    unsigned char* mask; // [m], zeroed by the caller: 1 = the row needs the shifted path
    int* n_flagged;
};

__global__ void finalize_kernel(FinalizeParams P) {
    const PairJob jb = *P.job;
    long long row = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (row >= jb.m) return;
    long long tt = row / P.tb;
    double sj = 0, sm = 0;
    {
        const long long ufirst = jb.unit_begin + (jb.unit_list ? jb.tile_first[tt] : tt * jb.n_train_tiles);
        const long long ucount = jb.unit_list ? jb.tile_first[tt + 1] - jb.tile_first[tt] : jb.n_train_tiles;
        if (ucount > 0) {
            int first = (int)(ufirst / P.upb);
            int last = (int)((ufirst + ucount - 1) / P.upb);
            for (int s = 0; s <= last - first; ++s) {
                sj += jb.part[(long long)s * jb.m_pad + row];
                if (P.ckde) sm += jb.part[((long long)jb.slots + s) * jb.m_pad + row];
            }
        }
    }
    if (P.jobA && jb.tile_first[tt + 1] == jb.tile_first[tt]) {
        // tile skipping: pass B starts from the pass-A sums (PairJob::init_sums); a test tile without any pass-B unit
        // has only its pass-A slots
        const PairJob ja = *P.jobA;
        const long long ufirst = ja.unit_begin + ja.tile_first[tt], ucount = ja.tile_first[tt + 1] - ja.tile_first[tt];
        if (ucount > 0) {
            const int first = (int)(ufirst / P.upbA), last = (int)((ufirst + ucount - 1) / P.upbA);
            for (int s = 0; s <= last - first; ++s) {
                sj += ja.part[(long long)s * ja.m_pad + row];
                if (P.ckde) sm += ja.part[((long long)ja.slots + s) * ja.m_pad + row];
            }
        }
    }
    bool bad = !(sj >= P.thresh) || (P.ckde && !(sm >= P.thresh));
    // NaN sums (NaN inputs) are not "underflow": propagate them
    if (sj != sj || (P.ckde && sm != sm)) bad = false;
    if (bad) {
        atomicAdd(P.n_flagged, 1);
        P.mask[row] = 1;
        return;
    }
    double v = P.lognorm_joint + log(sj);
    if (P.ckde) v -= P.lognorm_marg + log(sm);
    P.out[row] = v;
}

// ------------------------------------------------------------------------------------
// shifted second pass over the rows finalize_kernel flagged (their unshifted sums underflowed: test rows tens of
// bandwidths away from every training row).  Everything is sized and scheduled ON THE DEVICE from the flagged count, so
// the common case (no flagged row) costs five empty launches and no host synchronisation:
//   compact_flagged_kernel  mask -> ascending list of row ids (deterministic order), shifts reset to +inf
//   shift_prep_kernel       job / unit schedule of the second pair-kernel launch from the count
//   rowmin_kernel           per flagged row the smallest squared distance to a training row (FP32 arithmetic: the shift
//                           only has to bring the largest term near 1, not to be exact)
//   pair_kernel<SHIFT>      the same tiled stream-K kernel with the per-row shift on the exponent (pair_kernel.cuh)
//   finalize_shift_kernel   logl = lognorm + log(sum) - shift; rows that still have no usable sum (farther than 2^31
//                           kernel units, non-finite coordinates) go to row_kernel, one CTA per row
// Replaces the reference's max-shifted logsumexp_cols_offset (opencl/opencl_config.hpp:517-536), which it runs for every row.
// ------------------------------------------------------------------------------------
__global__ void compact_flagged_kernel(const unsigned char* __restrict__ mask, long long m, const int* __restrict__ n_flagged,
                                       int* __restrict__ flagged) {
    __shared__ int warp_tot[32];
    __shared__ int base_s;
    const int cnt = *n_flagged;
    if (cnt == 0) return;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, nw = blockDim.x >> 5;
    if (tid == 0) base_s = 0;
    __syncthreads();
    for (long long r0 = 0; r0 < m; r0 += blockDim.x) {
        const long long row = r0 + tid;
        const int f = (row < m && mask[row]) ? 1 : 0;
        const unsigned bal = __ballot_sync(0xffffffffu, f);
        if (lane == 0) warp_tot[w] = __popc(bal);
        __syncthreads();
        int off = base_s;
        for (int q = 0; q < w; ++q) off += warp_tot[q];
        if (f) flagged[off + __popc(bal & ((1u << lane) - 1u))] = (int)row;
        __syncthreads();
        if (tid == 0) {
            int t = 0;
            for (int q = 0; q < nw; ++q) t += warp_tot[q];
            base_s += t;
        }
        __syncthreads();
        if (base_s == cnt) break;  // every flagged row has been listed
    }
}

__global__ void shift_init_kernel(const int* __restrict__ n_flagged, float* __restrict__ shift_j, float* __restrict__ shift_m) {
    const int cnt = *n_flagged;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x) {
        shift_j[i] = INFINITY;
        shift_m[i] = INFINITY;
    }
}

struct ShiftPrepParams {
    PairJob job;          // train / norms / part / shifts of the second launch; m and the schedule are filled here
    const int* n_flagged;
    PairJob* d_job;
    long long* dyn;       // {total_units, upb}
    int tb, grid;
    int* n_flagged2;
};

__global__ void shift_prep_kernel(ShiftPrepParams P) {
    const long long cnt = *P.n_flagged;
    PairJob j = P.job;
    j.m = cnt;
    j.m_pad = (cnt + 31) / 32 * 32;
    j.n_test_tiles = (int)((cnt + P.tb - 1) / P.tb);
    const long long U = (long long)j.n_test_tiles * j.n_train_tiles;
    const long long upb = U > 0 ? (U + P.grid - 1) / P.grid : 1;
    long long slots = (j.n_train_tiles + upb - 1) / upb + 1;
    j.slots = (int)(slots < P.grid ? slots : P.grid);
    j.unit_begin = 0;
    *P.d_job = j;
    P.dyn[0] = U;
    P.dyn[1] = upb;
    *P.n_flagged2 = 0;
}

// grid = (row tiles, training splits); the minima of the splits are combined with an integer atomicMin (non-negative
// floats order like their bit patterns)
template <typename T, int D, bool CKDE>
__global__ void rowmin_kernel(const T* __restrict__ train, long long n, const T* __restrict__ test, const int* __restrict__ rows,
                              const int* __restrict__ n_rows, float* __restrict__ smin_j, float* __restrict__ smin_m) {
    constexpr int TP = 256;
    __shared__ float tile[TP * D];
    const int cnt = *n_rows;
    if (cnt == 0) return;
    const long long chunk = ((n + gridDim.y - 1) / gridDim.y + TP - 1) / TP * TP;
    const long long i0 = blockIdx.y * chunk;
    const long long i1 = i0 + chunk < n ? i0 + chunk : n;
    for (long long rt = blockIdx.x; rt * blockDim.x < cnt; rt += gridDim.x) {
        const long long f = rt * blockDim.x + threadIdx.x;
        const bool ok = f < cnt;
        const long long row = ok ? rows[f] : rows[0];
        float yt[D];
#pragma unroll
        for (int c = 0; c < D; ++c) yt[c] = static_cast<float>(test[row * D + c]);
        float mj = INFINITY, mm = INFINITY;
        for (long long base = i0; base < i1; base += TP) {
            const int np = (int)(i1 - base < TP ? i1 - base : TP);
            __syncthreads();
            for (int q = threadIdx.x; q < np * D; q += blockDim.x) tile[q] = static_cast<float>(train[base * D + q]);
            __syncthreads();
#pragma unroll 4
            for (int i = 0; i < np; ++i) {
                float sq = 0.f, sm = 0.f;
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    const float dl = yt[c] - tile[i * D + c];
                    sq = fmaf(dl, dl, sq);
                    if (CKDE && c == D - 2) sm = sq;
                }
                mj = fminf(mj, sq);
                if (CKDE) mm = fminf(mm, sm);
            }
        }
        if (ok && i0 < i1) {
            if (mj == mj) atomicMin(reinterpret_cast<int*>(smin_j + f), __float_as_int(mj));
            if (CKDE && mm == mm) atomicMin(reinterpret_cast<int*>(smin_m + f), __float_as_int(mm));
        }
    }
}

template <typename T>
static cudaError_t launch_rowmin(int d, bool ckde, dim3 grid, cudaStream_t st, const void* train, long long n, const void* test,
                                 const int* rows, const int* n_rows, float* smin_j, float* smin_m) {
#define PBN_RM_CASE(DD)                                                                                                        \
    case DD:                                                                                                                   \
        if (ckde) rowmin_kernel<T, (DD < 2 ? 2 : DD), true><<<grid, 256, 0, st>>>(static_cast<const T*>(train), n,              \
                                                                                 static_cast<const T*>(test), rows, n_rows,   \
                                                                                 smin_j, smin_m);                              \
        else rowmin_kernel<T, DD, false><<<grid, 256, 0, st>>>(static_cast<const T*>(train), n, static_cast<const T*>(test),    \
                                                               rows, n_rows, smin_j, smin_m);                                  \
        break;
    switch (d) {
        PBN_RM_CASE(1) PBN_RM_CASE(2) PBN_RM_CASE(3) PBN_RM_CASE(4) PBN_RM_CASE(5)
        PBN_RM_CASE(6) PBN_RM_CASE(7) PBN_RM_CASE(8) PBN_RM_CASE(9) PBN_RM_CASE(10)
        default: return cudaErrorInvalidValue;
    }
#undef PBN_RM_CASE
    return cudaGetLastError();
}

struct FinalizeShiftParams {
    const PairJob* job;      // the second launch's job (device)
    const long long* dyn;    // {total_units, upb}
    int tb, ckde, f64;
    double lognorm_joint, lognorm_marg, u2ln, thresh;
    double* out;             // [m] of the ORIGINAL row numbering
    const int* flagged;      // flagged position -> original row
    int* flagged2;           // rows still without a usable sum
    int* n_flagged2;
};

__global__ void finalize_shift_kernel(FinalizeShiftParams P) {
    const PairJob jb = *P.job;
    const long long upb = P.dyn[1];
    for (long long f = blockIdx.x * (long long)blockDim.x + threadIdx.x; f < jb.m; f += (long long)gridDim.x * blockDim.x) {
        const long long tt = f / P.tb;
        const long long ustart = tt * jb.n_train_tiles;
        const int first = (int)(ustart / upb);
        const int last = (int)((ustart + jb.n_train_tiles - 1) / upb);
        double sj = 0, sm = 0;
        for (int s = 0; s <= last - first; ++s) {
            sj += jb.part[(long long)s * jb.m_pad + f];
            if (P.ckde) sm += jb.part[((long long)jb.slots + s) * jb.m_pad + f];
        }
        // the shift the kernel applied (pair_kernel: integer part for f64, the float itself for f32)
        float fj = jb.shift_j[f], fm = P.ckde ? jb.shift_m[f] : 0.f;
        double aj = P.f64 ? (fj < 2.0e9f ? (double)rintf(fj) : 0.0) : (double)fj;
        double am = P.f64 ? (fm < 2.0e9f ? (double)rintf(fm) : 0.0) : (double)fm;
        const bool bad = !(sj >= P.thresh) || !(sj < INFINITY) || (P.ckde && (!(sm >= P.thresh) || !(sm < INFINITY)));
        const int row = P.flagged[f];
        if (bad) {
            P.flagged2[atomicAdd(P.n_flagged2, 1)] = row;
            continue;
        }
        double v = P.lognorm_joint + log(sj) - aj * P.u2ln;
        if (P.ckde) v -= P.lognorm_marg + log(sm) - am * P.u2ln;
        P.out[row] = v;
    }
}

// Robust per-row evaluation (max-shifted two-pass log-sum-exp, one CTA per test row).
// Used for rows flagged by finalize_kernel and as the generic path for d > kMaxFastD (10).
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

static int fit_one(pbn_ctx* ctx, const pbn_table* tbl, const int* cols, int d, pbn_rows rows, const double* H, bool ckde,
                   pbn_kde** out);
// tile skipping pays from a few dozen tiles on each side (tools/skip_model.py: a 90k x 10k cross-validation fold has too few)
// (PBN_SKIP_MIN_TRAIN / PBN_SKIP_MIN_TEST override them: the sanitizer runs and the small-size tests force the path)
static int64_t env_rows(const char* name, int64_t dflt) {
    const char* e = getenv(name);
    return (e && *e) ? (int64_t)atoll(e) : dflt;
}
static const int64_t kSkipMinTrainEnv = env_rows("PBN_SKIP_MIN_TRAIN", 0);
static const int64_t kSkipMinTest = env_rows("PBN_SKIP_MIN_TEST", 1 << 14);
// Training rows from which the Morton-sorted copy is kept, by the number of kernel coordinates the sums are local in (the
// marginal ones of a CKDE).  B200, N = m, float64, all pairs -> sorted path (profiles/r2_tuning.md section 10): one coordinate
// 50k rows +34% (KDE d=1) / +19% (CKDE d=2), 100k rows +65% / +33%; two coordinates 50k -8%, 100k +22%; three and more lose
// below ~130k (100k: 0% / -7%).
static int64_t skip_min_train(int d, bool ckde) {
    if (kSkipMinTrainEnv > 0) return kSkipMinTrainEnv;
    const int dn = d - (ckde ? 1 : 0);
    return dn <= 1 ? 40000 : (dn == 2 && !ckde) ? 90000 : (1 << 17);  // (CKDE d=3 at 100k: 0 .. -8%, stays at 2^17)
}

// multi-device context: the fitted model is replicated (every device whitens its own copy of the training rows)
static int fit_impl(pbn_ctx* ctx, const pbn_table* tbl, const int* cols, int d, pbn_rows rows, const double* H,
                    bool ckde, pbn_kde** out) {
    if (!ctx || !out) return set_error(PBN_ERR_ARG, "null argument");
    if (!pbn_replicated(ctx, tbl)) return fit_one(ctx, tbl, cols, d, rows, H, ckde, out);
    const int nd = pbn_num_devices(ctx);
    std::vector<pbn_kde*> k(nd, nullptr);
    int rc = pbn_run_on_devices(nd, [&](int i) {
        return fit_one(pbn_device_ctx(ctx, i), pbn_replica(const_cast<pbn_table*>(tbl), i), cols, d, rows, H, ckde, &k[i]);
    });
    if (rc != PBN_OK) {
        for (pbn_kde* q : k)
            if (q) pbn_kde_free(q);
        return rc;
    }
    k[0]->rep.assign(k.begin() + 1, k.end());
    *out = k[0];
    return PBN_OK;
}

static int fit_one(pbn_ctx* ctx, const pbn_table* tbl, const int* cols, int d, pbn_rows rows, const double* H,
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
    size_t nbytes = (k->dtype == PBN_F64 && d <= pbn::kMaxFastD) ? ((size_t)n_pad * sizeof(double) + 255) / 256 * 256 : 0;
    cudaError_t e = cudaMallocAsync(&k->y, ybytes + 256 + nbytes, ctx->stream);
    if (e != cudaSuccess) { delete k; PBN_CUDA_TRY(e); }
    e = cudaMemsetAsync(k->y, 0, ybytes + 256 + nbytes, ctx->stream);
    if (e != cudaSuccess) { cudaFreeAsync(k->y, ctx->stream); delete k; PBN_CUDA_TRY(e); }
    k->d_bound = reinterpret_cast<float*>(static_cast<char*>(k->y) + ybytes);
    k->nrm = nbytes ? reinterpret_cast<double*>(static_cast<char*>(k->y) + ybytes + 256) : nullptr;
    rc = pbn_whiten_kde(ctx, k, tbl, cols, rows, k->y, k->d_bound, k->nrm);
    if (rc != PBN_OK) { cudaFreeAsync(k->y, ctx->stream); delete k; return rc; }
    // tile skipping: a second copy of the whitened rows in Morton order with one bounding box per training tile
    if (d <= pbn::kMaxFastD && n >= skip_min_train(d, ckde)) {
        const int n_tiles = (int)((n + tile - 1) / tile);
        const size_t bbytes = ((size_t)n_tiles * 2 * d * sizeof(float) + 255) / 256 * 256;
        int* perm = nullptr;
        e = cudaMallocAsync(&k->ys, ybytes + nbytes + bbytes, ctx->stream);
        if (e == cudaSuccess) e = cudaMemsetAsync(k->ys, 0, ybytes + nbytes + bbytes, ctx->stream);
        if (e == cudaSuccess) e = cudaMallocAsync(&perm, (size_t)n * sizeof(int), ctx->stream);
        if (e == cudaSuccess) {
            k->nrm_s = nbytes ? reinterpret_cast<double*>(static_cast<char*>(k->ys) + ybytes) : nullptr;
            k->box = reinterpret_cast<float*>(static_cast<char*>(k->ys) + ybytes + nbytes);
            k->n_box_tiles = n_tiles;
            rc = pbn_spatial_sort(ctx, k->dtype, d, pbn_morton_dims(k), k->y, k->nrm, n, k->d_bound, k->ys, k->nrm_s, perm);
            if (rc == PBN_OK) rc = pbn_spatial_boxes(ctx, k->dtype, d, k->ys, n, tile, k->box);
        }
        if (perm) cudaFreeAsync(perm, ctx->stream);
        if (e != cudaSuccess || rc != PBN_OK) {  // skipping is an optimisation: the model works without it
            if (k->ys) cudaFreeAsync(k->ys, ctx->stream);
            k->ys = nullptr;
            k->nrm_s = nullptr;
            k->box = nullptr;
            k->n_box_tiles = 0;
            cudaGetLastError();
        }
    }
    *out = k;
    return PBN_OK;
}

// The shifted second pass as a unit (used by pbn_logl_impl below and by the batched scores of cv.cu): the rows
// `flagged[0 .. *n_flagged)` (ascending ids into `test`) are evaluated against all `n_train` training rows with a
// per-row exponent shift; out[id] receives lognorm + log(sum) - shift (minus the marginal for a CKDE).  Sized and
// scheduled on the device from the count; nothing is read back.
int pbn_shift_pass(pbn_ctx* ctx, Scratch& sc, const ShiftPass& A) {
    cudaStream_t st = ctx->stream;
    const int d = A.d;
    const bool f64 = A.f64;
    const int TILE = f64 ? pbn::pair_tile_f64(d) : pbn::pair_tile_f32(d);
    const int TB = f64 ? pbn::pair_tb_for_f64(d, A.ckde) : pbn::pair_tb_for_f32(d, A.ckde);
    const int max_grid = ctx->sm_count * (f64 ? pbn::pair_ctas_per_sm_f64() : pbn::pair_ctas_per_sm_f32());
    const int n_acc = A.ckde ? 2 : 1;
    const int64_t m = A.m_cap;
    const double u2ln = 1.0 / unit_scale(f64 ? PBN_F64 : PBN_F32);
    // [PairJob][dyn: 2 x i64][counter][flagged2: m int][shift_j, shift_m: m float]
    const size_t jsz = (sizeof(PairJob) + 15) / 16 * 16;
    const size_t o_dyn = jsz, o_cnt = o_dyn + 16, o_fl2 = o_cnt + 16, o_sj = o_fl2 + (size_t)m * 4, o_sm = o_sj + (size_t)m * 4;
    char* book = nullptr;
    double* part2 = nullptr;
    PBN_CUDA_TRY(sc.alloc(&book, o_sm + (size_t)m * 4));
    // slots' x m_pad' <= grid x TB + 2 (m + TB) whatever the flagged count is (see shift_prep_kernel)
    PBN_CUDA_TRY(sc.alloc(&part2, (size_t)n_acc * ((size_t)max_grid * TB + 2 * ((size_t)m + TB) + 64) * sizeof(double)));
    PairJob* d_job2 = reinterpret_cast<PairJob*>(book);
    long long* dyn = reinterpret_cast<long long*>(book + o_dyn);
    int* n_flagged2 = A.n_row_kernel ? A.n_row_kernel : reinterpret_cast<int*>(book + o_cnt);
    int* flagged2 = reinterpret_cast<int*>(book + o_fl2);
    float* shift_j = reinterpret_cast<float*>(book + o_sj);
    float* shift_m = reinterpret_cast<float*>(book + o_sm);
    shift_init_kernel<<<(int)std::min<int64_t>((m + 255) / 256, 1024), 256, 0, st>>>(A.n_flagged, shift_j, shift_m);
    ShiftPrepParams SP;
    memset(&SP.job, 0, sizeof(SP.job));
    SP.job.train = A.train;
    SP.job.test = A.test;
    SP.job.part = part2;
    SP.job.n_train = A.n_train;
    SP.job.n_train_tiles = (int)((A.n_train + TILE - 1) / TILE);
    SP.job.shift_j = shift_j;   // (no bounds, norms, unit list or initial sums: the shifted tile takes the wide-range
    SP.job.shift_m = shift_m;   //  difference form against EVERY training tile)
    SP.job.test_rows = A.flagged;
    SP.n_flagged = A.n_flagged;
    SP.d_job = d_job2;
    SP.dyn = dyn;
    SP.tb = TB;
    SP.grid = max_grid;
    SP.n_flagged2 = n_flagged2;
    shift_prep_kernel<<<1, 1, 0, st>>>(SP);
    // few flagged rows: split the training rows so that the scan still spreads over the GPU
    dim3 rmgrid((unsigned)std::max<int64_t>(1, std::min<int64_t>((m + 255) / 256, (int64_t)ctx->sm_count * 2)), 8);
    cudaError_t e2 = f64 ? launch_rowmin<double>(d, A.ckde, rmgrid, st, A.train, A.n_train, A.test, A.flagged, A.n_flagged, shift_j, shift_m)
                         : launch_rowmin<float>(d, A.ckde, rmgrid, st, A.train, A.n_train, A.test, A.flagged, A.n_flagged, shift_j, shift_m);
    PBN_CUDA_TRY(e2);
    e2 = f64 ? pbn::launch_pair_shift_f64(d, A.ckde, d_job2, dyn, max_grid, ctx->d_exp_tab, st)
             : pbn::launch_pair_shift_f32(d, A.ckde, d_job2, dyn, max_grid, ctx->d_exp_tab, st);
    PBN_CUDA_TRY(e2);
    FinalizeShiftParams FS;
    FS.job = d_job2;
    FS.dyn = dyn;
    FS.tb = TB;
    FS.ckde = A.ckde ? 1 : 0;
    FS.f64 = f64 ? 1 : 0;
    FS.lognorm_joint = A.lognorm_joint;
    FS.lognorm_marg = A.lognorm_marg;
    FS.u2ln = u2ln;
    FS.thresh = ldexp(1.0, -40);  // the largest term of a shifted row is ~1
    FS.out = A.out;
    FS.flagged = A.flagged;
    FS.flagged2 = flagged2;
    FS.n_flagged2 = n_flagged2;
    finalize_shift_kernel<<<(int)std::max<int64_t>(1, std::min<int64_t>((m + 255) / 256, (int64_t)ctx->sm_count * 4)), 256, 0, st>>>(FS);
    ctx->launches += 5;
    PBN_CUDA_TRY(cudaGetLastError());
    // rows the shifted pass could not evaluate either: exact per-row evaluation (count read on device)
    RowParams R;
    R.train = A.train;
    R.test = A.test;
    R.n = A.n_train;
    R.d = d;
    R.ckde = A.ckde ? 1 : 0;
    R.lognorm_joint = A.lognorm_joint;
    R.lognorm_marg = A.lognorm_marg;
    R.u2ln = u2ln;
    R.rows = flagged2;
    R.count_ptr = n_flagged2;
    R.count = 0;
    R.out = A.out;
    const int rgrid = (int)std::max<int64_t>(1, std::min<int64_t>(m, (int64_t)ctx->sm_count * 8));
    if (f64) row_kernel<double><<<rgrid, 256, 0, st>>>(R);
    else row_kernel<float><<<rgrid, 256, 0, st>>>(R);
    ctx->launches++;
    PBN_CUDA_TRY(cudaGetLastError());
    return PBN_OK;
}

static int logl_one(pbn_ctx* ctx, const pbn_kde* k, const pbn_table* test, const int* cols, pbn_rows rows,
                    double* d_out_logl, double* d_out_slogl, double* h_out_logl, double* h_out_slogl);

// rows each device must at least get before a call is sharded (below that one device finishes sooner than the threads start)
static bool worth_sharding(const pbn_ctx* ctx, int64_t n_train, int64_t m) {
    const int nd = pbn_num_devices(ctx);
    return m >= (int64_t)2048 * nd && (double)n_train * (double)m >= 2.0e9 * nd;
}

// multi-device context, host outputs: contiguous shards of the test rows against the replicated model, one host thread
// per device; slogl = the per-device sums added in device order (train replicated, test rows sharded)
// Pass B of tile skipping uses the group-skipping kernel (pair_kernel<..., GSKIP>, tile_f64_dot_gskip): float64 families of up
// to 8 variables with two or more kernel coordinates (one DFMA per exponent is too little to save: KDE d=1 lost 3%);
// PBN_GROUP_SKIP=0 keeps the plain kernel (A/B measurements).
// float32 (tile_f32_packed_gskip): CKDE of 3..6 variables (1M x 1M with skipping: d=4 6.96e12 -> 8.47e12, d=5 4.65e12 ->
// 5.61e12, d=6 3.73e12 -> 4.49e12, d=3 +4%; the KDE shapes and CKDE d=2 lose 2-17% - their plain kernel has the MUFU offload
// and little else to save).  PBN_GROUP_SKIP_F32 overrides the shape mask (bit d: KDE of d variables, bit 8 + d: CKDE).
static bool pair_group_skip(bool f64, int d, bool ckde) {
    static const bool enabled = !(getenv("PBN_GROUP_SKIP") && atoi(getenv("PBN_GROUP_SKIP")) == 0);
    static const unsigned f32_mask = getenv("PBN_GROUP_SKIP_F32") ? (unsigned)strtoul(getenv("PBN_GROUP_SKIP_F32"), nullptr, 0) : 0x7800u;
    if (!enabled || d > 8) return false;
    if (!f64) return (f32_mask >> ((ckde ? 8 : 0) + d)) & 1u;
    return d - (ckde ? 1 : 0) >= 2;
}

int pbn_logl_impl(pbn_ctx* ctx, const pbn_kde* k, const pbn_table* test, const int* cols, pbn_rows rows,
                     double* d_out_logl, double* d_out_slogl, double* h_out_logl, double* h_out_slogl) {
    if (!ctx || !k) return set_error(PBN_ERR_ARG, "null argument");
    const int64_t m_all = (test && rows.e0 >= rows.b0 && rows.e1 >= rows.b1) ? seg_count(rows) : 0;
    for (pbn_ctx* p : ctx->peers) p->last_units_total = p->last_units_done = 0;  // "last call" counters are per call
    if (d_out_logl || d_out_slogl || !pbn_replicated(ctx, k) || !pbn_replicated(ctx, test) || !worth_sharding(ctx, k->n, m_all))
        return logl_one(ctx, k, test, cols, rows, d_out_logl, d_out_slogl, h_out_logl, h_out_slogl);
    PBN_TRY(check_cols(test, cols, k->d));
    PBN_TRY(check_rows(test, rows));
    const int nd = pbn_num_devices(ctx);
    std::vector<double> part(nd, 0.0);
    int rc = pbn_run_on_devices(nd, [&](int i) {
        const int64_t b = m_all * i / nd, e = m_all * (i + 1) / nd;
        return logl_one(pbn_device_ctx(ctx, i), pbn_replica(const_cast<pbn_kde*>(k), i), pbn_replica(const_cast<pbn_table*>(test), i),
                        cols, pbn_sub_rows(rows, b, e), nullptr, nullptr, h_out_logl ? h_out_logl + b : nullptr,
                        h_out_slogl ? &part[i] : nullptr);
    });
    if (rc != PBN_OK) return rc;
    int64_t fb = 0, rk = 0;
    for (int i = 0; i < nd; ++i) {
        fb += pbn_device_ctx(ctx, i)->last_fallback_rows;
        rk += pbn_device_ctx(ctx, i)->last_row_kernel_rows;
    }
    ctx->last_fallback_rows = fb;
    ctx->last_row_kernel_rows = rk;
    if (h_out_slogl) {
        double s = 0;
        for (int i = 0; i < nd; ++i) s += part[i];
        *h_out_slogl = s;
    }
    return PBN_OK;
}

static int logl_one(pbn_ctx* ctx, const pbn_kde* k, const pbn_table* test, const int* cols, pbn_rows rows,
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
    const bool fast = d <= pbn::kMaxFastD;
    Scratch sc(st);

    char* ytest = nullptr;
    size_t ytbytes = ((size_t)m * d * es + 255) / 256 * 256;
    const size_t tnbytes = k->nrm ? ((size_t)m * sizeof(double) + 255) / 256 * 256 : 0;
    PBN_CUDA_TRY(sc.alloc(&ytest, ytbytes + 256 + tnbytes));
    float* bound_test = reinterpret_cast<float*>(ytest + ytbytes);
    double* nrm_test = tnbytes ? reinterpret_cast<double*>(ytest + ytbytes + 256) : nullptr;
    PBN_CUDA_TRY(cudaMemsetAsync(bound_test, 0, 256, st));
    PBN_TRY(pbn_whiten_kde(ctx, k, test, cols, rows, ytest, bound_test, nrm_test));

    double* out = d_out_logl;
    if (!out) PBN_CUDA_TRY(sc.alloc(&out, (size_t)m * sizeof(double)));
    const double u2ln = 1.0 / unit_scale(k->dtype);  // (kernel units of s'/ -t) -> natural log

    // tile skipping (spatial.cu): test rows in Morton order against the Morton-ordered copy of the training rows; every
    // kernel below then works in that order and the results are scattered back at the end
    const bool use_skip = ctx->skipping && fast && k->ys && m >= kSkipMinTest;
    const void* train_y = k->y;
    const double* train_nrm = k->nrm;
    int* perm = nullptr;
    double* out_final = out;
    char* ytest_raw = ytest;          // the caller's row order: what the call falls back to when the boxes prove too little
    double* nrm_raw = nrm_test;
    bool use_sorted = use_skip;
    if (use_skip) {
        char* ys_test = nullptr;
        PBN_CUDA_TRY(sc.alloc(&ys_test, ytbytes + tnbytes));
        PBN_CUDA_TRY(sc.alloc(&perm, (size_t)m * sizeof(int)));
        PBN_CUDA_TRY(sc.alloc(&out, (size_t)m * sizeof(double)));
        double* nrm_sorted = tnbytes ? reinterpret_cast<double*>(ys_test + ytbytes) : nullptr;
        PBN_TRY(pbn_spatial_sort(ctx, k->dtype, d, pbn_morton_dims(k), ytest, nrm_test, m, bound_test, ys_test, nrm_sorted, perm));
        ytest = ys_test;
        nrm_test = nrm_sorted;
        train_y = k->ys;
        train_nrm = k->nrm_s;
    }

    RowParams R;
    R.train = train_y;
    R.test = ytest;
    R.n = k->n;
    R.d = d;
    R.ckde = k->ckde ? 1 : 0;
    R.lognorm_joint = k->lognorm_joint;
    R.lognorm_marg = k->lognorm_marg;
    R.u2ln = u2ln;
    R.out = out;
    const int rgrid = (int)std::min<int64_t>(m, (int64_t)ctx->sm_count * 8);

    if (fast) {
        int n_test_tiles = (int)((m + TB - 1) / TB);
        int n_train_tiles = (int)((k->n + TILE - 1) / TILE);
        long long U = (long long)n_test_tiles * n_train_tiles;
        const int max_grid = ctx->sm_count * (f64 ? pbn::pair_ctas_per_sm_f64() : pbn::pair_ctas_per_sm_f32());
        int grid = (int)std::min<long long>(U, max_grid);
        long long upb = (U + grid - 1) / grid;
        grid = (int)((U + upb - 1) / upb);
        int slots = (int)std::min<long long>((n_train_tiles + upb - 1) / upb + 1, grid);
        int n_acc = k->ckde ? 2 : 1;
        long long m_pad = (m + 31) / 32 * 32;
        // one carve for the bookkeeping: [PairJob x 2][counters: 2 x int][flagged: m int][mask: m bytes]
        const size_t jsz = (sizeof(PairJob) + 15) / 16 * 16;
        const size_t o_cnt = 2 * jsz, o_fl = o_cnt + 16, o_mask = o_fl + (size_t)m * 4;
        char* book = nullptr;
        double* part = nullptr;
        PBN_CUDA_TRY(sc.alloc(&book, o_mask + (size_t)m));
        PairJob* d_job = reinterpret_cast<PairJob*>(book);
        PairJob* d_jobA = reinterpret_cast<PairJob*>(book + jsz);
        int* n_flagged = reinterpret_cast<int*>(book + o_cnt);
        int* n_flagged2 = n_flagged + 1;
        int* flagged = reinterpret_cast<int*>(book + o_fl);
        unsigned char* mask = reinterpret_cast<unsigned char*>(book + o_mask);
        PBN_CUDA_TRY(cudaMemsetAsync(mask, 0, (size_t)m, st));
        PairJob job;
        memset(&job, 0, sizeof(job));
        job.train = train_y;
        job.test = ytest;
        job.part = part;
        job.bound_train = k->d_bound;
        job.bound_test = bound_test;
        job.train_nrm = train_nrm;
        job.test_nrm = nrm_test;
        job.n_train = k->n;
        job.m = m;
        job.m_pad = m_pad;
        job.unit_begin = 0;
        job.n_train_tiles = n_train_tiles;
        job.n_test_tiles = n_test_tiles;
        job.slots = slots;
        cudaEvent_t ev0 = nullptr, ev1 = nullptr;
        if (ctx->timing) {
            PBN_CUDA_TRY(cudaEventCreate(&ev0));
            sc.events.push_back(ev0);
            PBN_CUDA_TRY(cudaEventCreate(&ev1));
            sc.events.push_back(ev1);
            PBN_CUDA_TRY(cudaEventRecord(ev0, st));
        }
        long long upbA = 1;
        bool have_A = false;
        ctx->last_units_total = U;
        ctx->last_units_done = U;
        if (use_skip) {
            // pass A: every test tile against its nearest training tiles -> lower bounds of the sums
            const int ntt = n_test_tiles, ntr = n_train_tiles;
            const int KA = std::min(pbn::kNearTiles, ntr);
            const size_t b_box = ((size_t)ntt * 2 * d * sizeof(float) + 255) / 256 * 256, b_ll = ((size_t)(ntt + 1) * 8 + 255) / 256 * 256;
            char* sk = nullptr;
            PBN_CUDA_TRY(sc.alloc(&sk, b_box + 3 * b_ll + 2 * (((size_t)ntt * 4 + 255) / 256 * 256) + ((size_t)ntt * KA * 4 + 255) / 256 * 256));
            float* box_test = reinterpret_cast<float*>(sk);
            long long* iota = reinterpret_cast<long long*>(sk + b_box);
            long long* count = reinterpret_cast<long long*>(sk + b_box + b_ll);
            long long* tile_first = reinterpret_cast<long long*>(sk + b_box + 2 * b_ll);
            float* thr = reinterpret_cast<float*>(sk + b_box + 3 * b_ll);
            int* nearest = reinterpret_cast<int*>(sk + b_box + 3 * b_ll + 2 * (((size_t)ntt * 4 + 255) / 256 * 256));
            double* partA = nullptr;
            PBN_CUDA_TRY(sc.alloc(&partA, (size_t)n_acc * (KA + 2) * m_pad * sizeof(double)));
            double* sumsA = partA + (size_t)n_acc * (KA + 1) * m_pad;  // [n_acc][m_pad]: pass-A sums per row
            PBN_TRY(pbn_spatial_boxes(ctx, k->dtype, d, ytest, m, TB, box_test));
            PBN_TRY(pbn_skip_nearest(ctx, box_test, ntt, k->box, ntr, d, KA, nearest, iota));
            PairJob jobA = job;
            jobA.part = partA;
            jobA.slots = KA + 1;  // KA consecutive units are shared by at most KA + 1 CTAs
            jobA.unit_list = nearest;
            jobA.tile_first = iota;   // tile_first[tt] = tt * KA
            const long long UA = (long long)ntt * KA;
            int gridA = (int)std::min<long long>(UA, max_grid);
            upbA = (UA + gridA - 1) / gridA;
            gridA = (int)((UA + upbA - 1) / upbA);
            write_job_kernel<<<1, 1, 0, st>>>(jobA, d_jobA, nullptr);
            cudaError_t ea = f64 ? pbn::launch_pair_f64(d, k->ckde, d_jobA, 1, UA, upbA, gridA, ctx->d_exp_tab, st)
                                 : pbn::launch_pair_f32(d, k->ckde, d_jobA, 1, UA, upbA, gridA, ctx->d_exp_tab, st);
            ctx->launches += 2;
            PBN_CUDA_TRY(ea);
            // list B: the units the boxes cannot prove negligible (one small D2H of their number)
            long long total = 0;
            PBN_TRY(pbn_skip_count(ctx, d_jobA, upbA, TB, k->ckde ? 1 : 0, k->dtype, k->n, box_test, ntt, k->box, ntr, d, nearest, KA, thr,
                                   sumsA, count, tile_first, &total));
            // Share of the units above which the call gives up the sorted path (PBN_SKIP_KEEP_FRAC overrides, tuning).  The
            // float64 kernels with group skipping keep it even when the boxes prove nothing: inside the units a warp's 96
            // neighbouring rows still skip groups of training points (1M x 1M: KDE d=6 +12%, d=7 / 8 +8%, CKDE d=6 +30%).
            static const double keep_env = getenv("PBN_SKIP_KEEP_FRAC") ? atof(getenv("PBN_SKIP_KEEP_FRAC")) : 0.0;
            const double keep_frac = keep_env > 0.0 ? keep_env : (pair_group_skip(f64, d, k->ckde) ? 2.0 : 0.92);
            if ((double)(total + UA) > keep_frac * (double)ctx->last_units_total) {
                // (almost) nothing can be dropped - families of 6+ variables at these sizes: every pair is evaluated, in
                // the caller's row order, where the exponent floor of pair_floor works best
                use_sorted = false;
                ytest = ytest_raw;
                nrm_test = nrm_raw;
                out = out_final;
                job.train = k->y;
                job.train_nrm = k->nrm;
                job.test = ytest;
                job.test_nrm = nrm_test;
                R.train = k->y;
                R.test = ytest;
                R.out = out;
                goto plain_launch;
            }
            int* unit_list = nullptr;
            PBN_CUDA_TRY(sc.alloc(&unit_list, (size_t)std::max<long long>(total, 1) * sizeof(int)));
            PBN_TRY(pbn_skip_fill(ctx, k->ckde ? 1 : 0, box_test, ntt, k->box, ntr, d, nearest, KA, thr, tile_first, unit_list));
            job.unit_list = unit_list;
            job.tile_first = tile_first;
            job.init_sums = sumsA;  // pass B continues the pass-A sums
            ctx->last_units_done = total + UA;
            have_A = true;
            U = total;
            grid = (int)std::max<long long>(1, std::min<long long>(U, max_grid));
            upb = std::max<long long>(1, (U + grid - 1) / grid);
            grid = (int)std::max<long long>(1, (U + upb - 1) / upb);
            job.slots = slots = (int)std::min<long long>((n_train_tiles + upb - 1) / upb + 1, grid);
        }
    plain_launch:
        PBN_CUDA_TRY(sc.alloc(&part, (size_t)n_acc * slots * m_pad * sizeof(double)));
        job.part = part;
        write_job_kernel<<<1, 1, 0, st>>>(job, d_job, n_flagged);
        ctx->launches++;
        PBN_CUDA_TRY(cudaGetLastError());
        if (U > 0) {
            // pass B (unit list, Morton order) in float64: the kernel that also skips groups of training points inside a unit
            const bool gs = have_A && pair_group_skip(f64, d, k->ckde);
            cudaError_t e = gs    ? (f64 ? pbn::launch_pair_gskip_f64(d, k->ckde, d_job, 1, U, upb, grid, ctx->d_exp_tab, st)
                                         : pbn::launch_pair_gskip_f32(d, k->ckde, d_job, 1, U, upb, grid, ctx->d_exp_tab, st))
                            : f64 ? pbn::launch_pair_f64(d, k->ckde, d_job, 1, U, upb, grid, ctx->d_exp_tab, st)
                                  : pbn::launch_pair_f32(d, k->ckde, d_job, 1, U, upb, grid, ctx->d_exp_tab, st);
            ctx->launches++;
            PBN_CUDA_TRY(e);
        }
        if (ctx->timing) {
            PBN_CUDA_TRY(cudaEventRecord(ev1, st));
            ctx->timed.emplace_back(ev0, ev1);
            sc.events.clear();  // owned by ctx->timed from here on
            ctx->pair_units += (int64_t)k->n * m * (k->ckde ? 2 : 1);
            ctx->units_total += ctx->last_units_total;
            ctx->units_done += ctx->last_units_done;
        }
        FinalizeParams F;
        F.job = d_job;
        F.upb = upb;
        F.jobA = have_A ? d_jobA : nullptr;
        F.upbA = upbA;
        F.tb = TB;
        F.ckde = k->ckde ? 1 : 0;
        F.lognorm_joint = k->lognorm_joint;
        F.lognorm_marg = k->lognorm_marg;
        F.u2ln = u2ln;
        F.thresh = f64 ? ldexp(1.0, -900) : ldexp(1.0, -64);
        F.out = out;
        F.mask = mask;
        F.n_flagged = n_flagged;
        finalize_kernel<<<(int)((m + 255) / 256), 256, 0, st>>>(F);
        ctx->launches++;
        PBN_CUDA_TRY(cudaGetLastError());

        // ---- second pass over the flagged rows, scheduled on the device from their count ----
        compact_flagged_kernel<<<1, 1024, 0, st>>>(mask, m, n_flagged, flagged);
        ctx->launches++;
        ShiftPass SPA;
        SPA.train = train_y;
        SPA.n_train = k->n;
        SPA.test = ytest;
        SPA.m_cap = m;
        SPA.d = d;
        SPA.ckde = k->ckde;
        SPA.f64 = f64;
        SPA.lognorm_joint = k->lognorm_joint;
        SPA.lognorm_marg = k->lognorm_marg;
        SPA.flagged = flagged;
        SPA.n_flagged = n_flagged;
        SPA.out = out;
        SPA.n_row_kernel = n_flagged2;
        PBN_TRY(pbn_shift_pass(ctx, sc, SPA));
        if (h_out_logl || h_out_slogl) {
            int nf[2] = {0, 0};
            PBN_CUDA_TRY(cudaMemcpyAsync(nf, n_flagged, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
            PBN_CUDA_TRY(cudaStreamSynchronize(st));
            ctx->last_fallback_rows = nf[0];
            ctx->last_row_kernel_rows = nf[1];
            ctx->d2h += 8;
        }
    } else {
        R.rows = nullptr;
        R.count_ptr = nullptr;
        R.count = m;
        if (f64) row_kernel<double><<<rgrid, 256, 0, st>>>(R);
        else row_kernel<float><<<rgrid, 256, 0, st>>>(R);
        ctx->launches++;
        PBN_CUDA_TRY(cudaGetLastError());
    }

    if (use_sorted) {  // back to the caller's row order
        PBN_TRY(pbn_scatter_out(ctx, out, perm, m, out_final));
        out = out_final;
    }
    double* d_sum = d_out_slogl;
    if (h_out_slogl || d_out_slogl) {
        int sb = (int)std::min<int64_t>((m + 255) / 256, (int64_t)ctx->sm_count * 4);
        double* partial = nullptr;
        PBN_CUDA_TRY(sc.alloc(&partial, (size_t)(sb + 1) * sizeof(double)));
        if (!d_sum) d_sum = partial + sb;
        sum_partial_kernel<<<sb, 256, 0, st>>>(out, m, partial);
        sum_final_kernel<<<1, 256, 0, st>>>(partial, sb, d_sum);
        ctx->launches += 2;
        PBN_CUDA_TRY(cudaGetLastError());
        if (h_out_slogl) {
            PBN_CUDA_TRY(cudaMemcpyAsync(h_out_slogl, d_sum, sizeof(double), cudaMemcpyDeviceToHost, st));
            ctx->d2h += 8;
            PBN_CUDA_TRY(cudaStreamSynchronize(st));
        }
    }
    if (h_out_logl) {
        PBN_CUDA_TRY(cudaMemcpyAsync(h_out_logl, out, (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, st));
        PBN_CUDA_TRY(cudaStreamSynchronize(st));
        ctx->d2h += m * 8;
    }
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
        // (times c2 for the completed-square polynomial, pair_kernel.cuh: PBN_EXP_SQ)
        double v = (double)(exp2l((long double)j / pbn::kExpTab) * (pbn::kExpSq ? (long double)pbn::kExpSqC2 : 1.0L));
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

// One context over several GPUs of this process (pbn_ctx_create(const int* devices, int n, ...)): the
// reference drives ONE device from ONE process (opencl/opencl_config.hpp:120-121, opencl_config.cpp:217-220), and a
// caller of KDE.logl / GreedyHillClimbing.estimate keeps doing exactly that while every visible B200 takes its share.
int pbn_ctx_create_multi(const int* devices, int n, pbn_ctx** out) {
    if (!devices || !out || n < 1) return set_error(PBN_ERR_ARG, "invalid device list");
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < i; ++j)
            if (devices[i] == devices[j]) return set_error(PBN_ERR_ARG, "a device is listed twice");
    pbn_ctx* c = nullptr;
    PBN_TRY(pbn_ctx_create(devices[0], &c));
    for (int i = 1; i < n; ++i) {
        pbn_ctx* p = nullptr;
        int rc = pbn_ctx_create(devices[i], &p);
        if (rc != PBN_OK) {
            std::string msg = pbn_last_error();
            pbn_ctx_destroy(c);
            return set_error(rc, msg);
        }
        c->peers.push_back(p);
    }
    *out = c;
    return PBN_OK;
}
int pbn_ctx_num_devices(pbn_ctx* ctx) { return ctx ? pbn_num_devices(ctx) : 0; }
int pbn_ctx_device(pbn_ctx* ctx, int i) {
    if (!ctx || i < 0 || i >= pbn_num_devices(ctx)) return -1;
    return pbn_device_ctx(ctx, i)->device;
}

int pbn_ctx_destroy(pbn_ctx* ctx) {
    if (!ctx) return PBN_OK;
    for (pbn_ctx* p : ctx->peers) pbn_ctx_destroy(p);
    ctx->peers.clear();
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
    for (pbn_ctx* p : ctx->peers) PBN_TRY(pbn_ctx_synchronize(p));
    DevSetter ds(ctx->device);
    PBN_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return PBN_OK;
}
int pbn_ctx_sm_count(pbn_ctx* ctx) { return ctx ? ctx->sm_count : 0; }
int pbn_ctx_counters(pbn_ctx* ctx, int64_t* launches, int64_t* h2d, int64_t* d2h) {
    if (!ctx) return set_error(PBN_ERR_ARG, "null context");
    int64_t l = ctx->launches, a = ctx->h2d, b = ctx->d2h;
    for (const pbn_ctx* p : ctx->peers) {  // a multi-device context reports the sums over its devices
        l += p->launches;
        a += p->h2d;
        b += p->d2h;
    }
    if (launches) *launches = l;
    if (h2d) *h2d = a;
    if (d2h) *d2h = b;
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
// Loads the kernel modules of every device of the context (CUDA loads a module on the first use of one of its kernels:
// ~2 s for the pair kernels' instantiations, which the first fit / logl / score call would otherwise pay).  Safe to call
// from another host thread while data are being prepared; pybnesian_b200 does so when a context is created.
int pbn_ctx_warmup(pbn_ctx* ctx) {
    if (!ctx) return set_error(PBN_ERR_ARG, "null context");
    return pbn_run_on_devices(pbn_num_devices(ctx), [&](int i) {
        pbn_ctx* c = pbn_device_ctx(ctx, i);
        DevSetter ds(c->device);
        PBN_CUDA_TRY(pbn::warm_pair_f64());
        PBN_CUDA_TRY(pbn::warm_pair_f32());
        PBN_CUDA_TRY(pbn::warm_pair_gskip_f64());
        PBN_CUDA_TRY(pbn::warm_pair_gskip_f32());
        PBN_CUDA_TRY(pbn::warm_pair_shift_f64());
        PBN_CUDA_TRY(pbn::warm_pair_shift_f32());
        cudaFuncAttributes a;
        PBN_CUDA_TRY(cudaFuncGetAttributes(&a, finalize_kernel));
        return PBN_OK;
    });
}
int pbn_ctx_set_skipping(pbn_ctx* ctx, int on) {
    if (!ctx) return set_error(PBN_ERR_ARG, "null context");
    ctx->skipping = on != 0;
    for (pbn_ctx* p : ctx->peers) p->skipping = on != 0;
    return PBN_OK;
}
int pbn_ctx_skip_stats(pbn_ctx* ctx, int64_t* last_total, int64_t* last_done, int64_t* timed_total, int64_t* timed_done, int reset) {
    if (!ctx) return set_error(PBN_ERR_ARG, "null context");
    int64_t lt = ctx->last_units_total, ld = ctx->last_units_done, tt = ctx->units_total, td = ctx->units_done;
    for (pbn_ctx* p : ctx->peers) {  // a multi-device context reports the sums over its devices
        lt += p->last_units_total;
        ld += p->last_units_done;
        tt += p->units_total;
        td += p->units_done;
        if (reset) { p->units_total = 0; p->units_done = 0; }
    }
    if (last_total) *last_total = lt;
    if (last_done) *last_done = ld;
    if (timed_total) *timed_total = tt;
    if (timed_done) *timed_done = td;
    if (reset) { ctx->units_total = 0; ctx->units_done = 0; }
    return PBN_OK;
}
int pbn_ctx_last_row_kernel_rows(pbn_ctx* ctx, int64_t* out) {
    if (!ctx || !out) return set_error(PBN_ERR_ARG, "null argument");
    *out = ctx->last_row_kernel_rows;
    return PBN_OK;
}

static int table_upload_one(pbn_ctx* ctx, const void* const* col_ptrs, int ncols, int64_t nrows, int dtype, pbn_table** out);

int pbn_table_upload(pbn_ctx* ctx, const void* const* col_ptrs, int ncols, int64_t nrows, int dtype, pbn_table** out) {
    if (!ctx || !col_ptrs || !out) return set_error(PBN_ERR_ARG, "null argument");
    if (ctx->peers.empty()) return table_upload_one(ctx, col_ptrs, ncols, nrows, dtype, out);
    // multi-device context: one replica per device, copied over each device's own PCIe link in parallel
    const int nd = pbn_num_devices(ctx);
    std::vector<pbn_table*> t(nd, nullptr);
    int rc = pbn_run_on_devices(nd, [&](int i) { return table_upload_one(pbn_device_ctx(ctx, i), col_ptrs, ncols, nrows, dtype, &t[i]); });
    if (rc != PBN_OK) {
        for (pbn_table* q : t)
            if (q) pbn_table_free(q);
        return rc;
    }
    t[0]->rep.assign(t.begin() + 1, t.end());
    *out = t[0];
    return PBN_OK;
}

static int table_upload_one(pbn_ctx* ctx, const void* const* col_ptrs, int ncols, int64_t nrows, int dtype, pbn_table** out) {
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
    for (pbn_table* r : t->rep) pbn_table_free(r);
    t->rep.clear();
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
    for (pbn_kde* r : k->rep) pbn_kde_free(r);
    k->rep.clear();
    DevSetter ds(k->ctx->device);
    if (k->y) cudaFreeAsync(k->y, k->ctx->stream);
    if (k->ys) cudaFreeAsync(k->ys, k->ctx->stream);
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
