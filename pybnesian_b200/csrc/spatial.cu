// spatial.cu — spatial order of whitened rows and the unit lists of tile skipping.
//
// A Gaussian kernel sum with a rule-of-thumb bandwidth is local: for CKDE d = 4 at 1M rows 84% of the (test, train)
// pairs contribute less than 2^-80 of the sum they join.  The reference evaluates every pair
// (kde/KDE.hpp:592-640); so does pair_kernel by default.  With skipping enabled (pbn_ctx_set_skipping) whitened rows are
// put in Morton order, every tile of rows gets a bounding box, and a (test tile, train tile) unit is dropped when the
// boxes alone prove that ALL its terms together are below 2^-kSkipBits of every row's sum:
//
//   pass A   each test tile against its kNearTiles nearest train tiles  ->  a lower bound S_lb(row) <= S(row) of every sum
//   list B   units (tt, nt), nt not among those, with  n_train * 2^(-Dmin^2(tt, nt))  >  2^-kSkipBits * min_rows S_lb
//            (Dmin = distance between the two boxes, in kernel units; joint AND marginal coordinates for a CKDE)
//   pass B   pair_kernel over list B (PairJob::unit_list); finalize adds the partial sums of A and B
//
// The terms of all dropped units of a row sum to less than 2^-kSkipBits S(row): a relative change of the sum, and an
// absolute change of logl, below 9.1e-13 for kSkipBits = 40 - two orders inside the 1e-10 bar even for a CKDE value that
// happens to be ~1e-2.  Results are written back in the caller's row order.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <cub/cub.cuh>

#include "internal.h"

namespace {

constexpr int kSkipBits = 40;

// ---- Morton keys ------------------------------------------------------------------------------------------------
// coordinate c of a row, quantised to `bits` bits over [-B, B] (B = largest |whitened coordinate| of the set)
// DK <= D: only the first DK coordinates enter the key (a CKDE sorts by its evidence coordinates: the marginal sum is the
// one that decides whether a unit can be dropped, and its boxes should be the compact ones)
template <typename T, int D, int DK>
__global__ void morton_key_kernel(const T* __restrict__ y, long long n, const float* __restrict__ bound, unsigned long long* __restrict__ keys,
                                  int* __restrict__ idx) {
    constexpr int bits = 60 / DK > 16 ? 16 : 60 / DK;
    const float B = fmaxf(*bound, 1e-30f);
    const float scale = (float)((1u << bits) - 1) / (2.f * B);
    for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
        unsigned q[DK];
#pragma unroll
        for (int c = 0; c < DK; ++c) {
            float v = (static_cast<float>(y[r * D + c]) + B) * scale;
            v = v != v ? 0.f : fminf(fmaxf(v, 0.f), (float)((1u << bits) - 1));
            q[c] = (unsigned)v;
        }
        unsigned long long k = 0;
#pragma unroll
        for (int b = bits - 1; b >= 0; --b)
#pragma unroll
            for (int c = 0; c < DK; ++c) k = (k << 1) | ((q[c] >> b) & 1u);
        keys[r] = k;
        idx[r] = (int)r;
    }
}

template <typename T, int D>
__global__ void gather_sorted_kernel(const T* __restrict__ y, const double* __restrict__ nrm, const int* __restrict__ perm, long long n,
                                     T* __restrict__ ys, double* __restrict__ nrm_s) {
    for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
        const long long src = perm[r];
#pragma unroll
        for (int c = 0; c < D; ++c) ys[r * D + c] = y[src * D + c];
        if (nrm) nrm_s[r] = nrm[src];
    }
}

// box[t][c] = min, box[t][D + c] = max over the rows of tile t, widened by one float ulp each way
template <typename T, int D>
__global__ void tile_box_kernel(const T* __restrict__ ys, long long n, int tile_rows, float* __restrict__ box) {
    __shared__ float slo[32][D], shi[32][D];
    const long long t = blockIdx.x;
    const long long r0 = t * tile_rows;
    const long long r1 = r0 + tile_rows < n ? r0 + tile_rows : n;
    float lo[D], hi[D];
#pragma unroll
    for (int c = 0; c < D; ++c) { lo[c] = INFINITY; hi[c] = -INFINITY; }
    bool bad = false;
    for (long long r = r0 + threadIdx.x; r < r1; r += blockDim.x)
#pragma unroll
        for (int c = 0; c < D; ++c) {
            const double v = static_cast<double>(ys[r * D + c]);
            if (!(fabs(v) < 1e30)) bad = true;  // NaN / inf / absurd: the tile gets an unbounded box and is never skipped
            lo[c] = fminf(lo[c], __double2float_rd(v));
            hi[c] = fmaxf(hi[c], __double2float_ru(v));
        }
#pragma unroll
    for (int c = 0; c < D; ++c) {
        for (int o = 16; o > 0; o >>= 1) {
            lo[c] = fminf(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
            hi[c] = fmaxf(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
        }
    }
    bad = __syncthreads_or(bad);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0)
#pragma unroll
        for (int c = 0; c < D; ++c) { slo[w][c] = lo[c]; shi[w][c] = hi[c]; }
    __syncthreads();
    if (threadIdx.x < D) {
        const int c = threadIdx.x;
        float a = INFINITY, b = -INFINITY;
        for (int q = 0; q < (int)(blockDim.x >> 5); ++q) { a = fminf(a, slo[q][c]); b = fmaxf(b, shi[q][c]); }
        if (bad) { a = -INFINITY; b = INFINITY; }
        box[t * 2 * D + c] = a;
        box[t * 2 * D + D + c] = b;
    }
}

// squared distance between two boxes over coordinates [0, dn), rounded down (a lower bound of every pair's distance)
__device__ __forceinline__ float box_gap2(const float* __restrict__ a, const float* __restrict__ b, int D, int dn) {
    float s = 0.f;
    for (int c = 0; c < dn; ++c) {
        const float g = fmaxf(fmaxf(__fsub_rd(a[c], b[D + c]), __fsub_rd(b[c], a[D + c])), 0.f);
        s = __fmaf_rd(g, g, s);
    }
    return s;
}

// nearest[tt * K + j], j = 0 .. K-1: the K train tiles closest to test tile tt (box distance, ties: closest box centres, then
// lowest index), in ascending tile order; K = min(kNearTiles, n_train_tiles).  One CTA per test tile, K selection rounds.
__global__ void nearest_tile_kernel(const float* __restrict__ box_test, const float* __restrict__ box_train, int n_train_tiles, int D,
                                    int K, int* __restrict__ nearest) {
    __shared__ float bt[2 * PBN_MAX_DIM];
    __shared__ float sbest[32][2];
    __shared__ int sidx[32];
    __shared__ int chosen[pbn::kNearTiles];
    const int tt = blockIdx.x;
    if (threadIdx.x < 2 * D) bt[threadIdx.x] = box_test[(long long)tt * 2 * D + threadIdx.x];
    __syncthreads();
    for (int round = 0; round < K; ++round) {
        float best = INFINITY, bestc = INFINITY;
        int bi = 0x7fffffff;
        for (int nt = threadIdx.x; nt < n_train_tiles; nt += blockDim.x) {
            bool taken = false;
            for (int q = 0; q < round; ++q) taken |= chosen[q] == nt;
            if (taken) continue;
            const float* b = box_train + (long long)nt * 2 * D;
            float g = box_gap2(bt, b, D, D);
            float cd = 0.f;
            for (int c = 0; c < D; ++c) {
                const float dc = 0.5f * ((bt[c] + bt[D + c]) - (b[c] + b[D + c]));
                cd = fmaf(dc, dc, cd);
            }
            if (!(g == g)) g = 0.f;        // unbounded boxes count as touching
            if (!(cd == cd)) cd = 0.f;
            if (g < best || (g == best && (cd < bestc || (cd == bestc && nt < bi)))) { best = g; bestc = cd; bi = nt; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const float g = __shfl_xor_sync(0xffffffffu, best, o), cd = __shfl_xor_sync(0xffffffffu, bestc, o);
            const int i = __shfl_xor_sync(0xffffffffu, bi, o);
            if (g < best || (g == best && (cd < bestc || (cd == bestc && i < bi)))) { best = g; bestc = cd; bi = i; }
        }
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        if (lane == 0) { sbest[w][0] = best; sbest[w][1] = bestc; sidx[w] = bi; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int q = 1; q < (int)(blockDim.x >> 5); ++q) {
                const float g = sbest[q][0], cd = sbest[q][1];
                const int i = sidx[q];
                if (g < best || (g == best && (cd < bestc || (cd == bestc && i < bi)))) { best = g; bestc = cd; bi = i; }
            }
            chosen[round] = bi;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {  // ascending tile order (insertion sort of <= 8 entries)
        for (int a = 1; a < K; ++a) {
            const int v = chosen[a];
            int b = a - 1;
            while (b >= 0 && chosen[b] > v) { chosen[b + 1] = chosen[b]; --b; }
            chosen[b + 1] = v;
        }
        for (int a = 0; a < K; ++a) nearest[(long long)tt * K + a] = chosen[a];
    }
}

__global__ void iota_kernel(long long* __restrict__ x, long long n, long long step) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) x[i] = i * step;
}

// thr[tt] (joint), thr[n_test_tiles + tt] (marginal): a unit whose box distance^2 reaches it is negligible for every row
// of the tile.  S_lb = the smallest pass-A sum of the tile's rows; no bound (thr = +inf: never skip) when some row of
// the tile has no usable pass-A sum yet.
struct LbParams {
    const pbn::PairJob* jobA;
    long long upbA;
    int tb, ckde;
    double unit_per_binade;  // kernel exponent units per factor 2 (K for f64, 1 for f32)
    double log2_ntrain;
    float* thr;
    double* sums;  // [n_acc][m_pad]: the pass-A sums of every row (PairJob::init_sums of pass B)
};

__global__ void skip_threshold_kernel(LbParams P) {
    __shared__ double sh[2][32];
    const pbn::PairJob jb = *P.jobA;
    const long long tt = blockIdx.x;
    // pass A: tile_first[tt] .. tile_first[tt + 1] units, spread over the slots of the CTAs that shared them
    const long long ufirst = jb.unit_begin + jb.tile_first[tt], ucount = jb.tile_first[tt + 1] - jb.tile_first[tt];
    const int nslots = ucount > 0 ? (int)((ufirst + ucount - 1) / P.upbA - ufirst / P.upbA) + 1 : 0;
    double mj = INFINITY, mm = INFINITY;
    for (long long row = tt * P.tb + threadIdx.x; row < (tt + 1) * P.tb && row < jb.m; row += blockDim.x) {
        double sj = 0.0, sm = P.ckde ? 0.0 : 1.0;
        for (int q = 0; q < nslots; ++q) {
            sj += jb.part[(long long)q * jb.m_pad + row];
            if (P.ckde) sm += jb.part[((long long)jb.slots + q) * jb.m_pad + row];
        }
        P.sums[row] = sj;
        if (P.ckde) P.sums[jb.m_pad + row] = sm;
        if (!(sj > 0.0) || !(sj < INFINITY)) sj = 0.0;
        if (!(sm > 0.0) || !(sm < INFINITY)) sm = 0.0;
        mj = fmin(mj, sj);
        mm = fmin(mm, sm);
    }
    for (int o = 16; o > 0; o >>= 1) {
        mj = fmin(mj, __shfl_xor_sync(0xffffffffu, mj, o));
        mm = fmin(mm, __shfl_xor_sync(0xffffffffu, mm, o));
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) { sh[0][w] = mj; sh[1][w] = mm; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int q = 1; q < (int)(blockDim.x >> 5); ++q) { mj = fmin(mj, sh[0][q]); mm = fmin(mm, sh[1][q]); }
        // 2^(-D2 / u) * n_train <= 2^-kSkipBits * S_lb   <=>   D2 >= u * (kSkipBits + log2 n_train - log2 S_lb)
        const double tj = mj > 0.0 ? P.unit_per_binade * (kSkipBits + P.log2_ntrain - log2(mj)) : INFINITY;
        const double tm = mm > 0.0 ? P.unit_per_binade * (kSkipBits + P.log2_ntrain - log2(mm)) : INFINITY;
        // floats rounded UP, never below zero: the comparison stays conservative
        P.thr[tt] = tj < 3.0e38 ? fmaxf(__double2float_ru(tj), 0.f) : INFINITY;
        P.thr[gridDim.x + tt] = tm < 3.0e38 ? fmaxf(__double2float_ru(tm), 0.f) : INFINITY;
    }
}

__device__ __forceinline__ bool unit_alive(const float* __restrict__ bt, const float* __restrict__ b, int D, int ckde, float thr_j,
                                           float thr_m) {
    const float gj = box_gap2(bt, b, D, D);
    if (!(gj >= thr_j)) return true;                     // (NaN compares false: kept)
    if (ckde) {
        const float gm = box_gap2(bt, b, D, D - 1);
        if (!(gm >= thr_m)) return true;
    }
    return false;
}

// MODE 0: count[tt] = live units of test tile tt (nearest excluded: pass A did it).  MODE 1: write them, ascending.
template <int MODE>
__global__ void skip_list_kernel(const float* __restrict__ box_test, const float* __restrict__ box_train, int n_train_tiles, int D,
                                 int ckde, const float* __restrict__ thr, const int* __restrict__ nearest, int K,
                                 long long* __restrict__ count, const long long* __restrict__ tile_first, int* __restrict__ unit_list) {
    __shared__ float bt[2 * PBN_MAX_DIM];
    __shared__ int wsum[32];
    __shared__ long long base;
    __shared__ int near_s[pbn::kNearTiles];
    const int tt = blockIdx.x;
    if (threadIdx.x < 2 * D) bt[threadIdx.x] = box_test[(long long)tt * 2 * D + threadIdx.x];
    if (threadIdx.x < K) near_s[threadIdx.x] = nearest[(long long)tt * K + threadIdx.x];
    if (threadIdx.x == 0) base = MODE == 1 ? tile_first[tt] : 0;
    __syncthreads();
    const float tj = thr[tt], tm = thr[gridDim.x + tt];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    long long total = 0;
    for (int n0 = 0; n0 < n_train_tiles; n0 += blockDim.x) {
        const int nt = n0 + threadIdx.x;
        bool in_a = false;
        for (int q = 0; q < K; ++q) in_a |= near_s[q] == nt;
        const bool alive = nt < n_train_tiles && !in_a && unit_alive(bt, box_train + (long long)nt * 2 * D, D, ckde, tj, tm);
        const unsigned bal = __ballot_sync(0xffffffffu, alive);
        if (lane == 0) wsum[w] = __popc(bal);
        __syncthreads();
        if (MODE == 1 && alive) {
            long long off = base;
            for (int q = 0; q < w; ++q) off += wsum[q];
            unit_list[off + __popc(bal & ((1u << lane) - 1u))] = nt;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int q = 0; q < nw; ++q) t += wsum[q];
            base += t;
            total += t;
        }
        __syncthreads();
    }
    if (MODE == 0 && threadIdx.x == 0) count[tt] = total;
}

// exclusive scan of count[0..n) into first[0..n], first[n] = total (single CTA; n = number of test tiles)
__global__ void scan_counts_kernel(const long long* __restrict__ count, long long n, long long* __restrict__ first) {
    __shared__ long long carry;
    __shared__ long long wtot[32];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (long long i0 = 0; i0 < n; i0 += blockDim.x) {
        const long long i = i0 + threadIdx.x;
        long long v = i < n ? count[i] : 0, x = v;
        for (int o = 1; o < 32; o <<= 1) {
            const long long y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) wtot[w] = x;
        __syncthreads();
        long long off = carry;
        for (int q = 0; q < w; ++q) off += wtot[q];
        if (i < n) first[i] = off + x - v;
        __syncthreads();
        if (threadIdx.x == 0) {
            long long t = 0;
            for (int q = 0; q < nw; ++q) t += wtot[q];
            carry += t;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) first[n] = carry;
}

template <typename T>
__global__ void scatter_out_kernel(const double* __restrict__ src, const int* __restrict__ perm, long long n, T* __restrict__ dst) {
    for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x)
        dst[perm[r]] = src[r];
}

}  // namespace

// ---- host launchers (declared in internal.h) --------------------------------------------------------------------
#define PBN_D_SWITCH(CALL)                                                         \
    switch (d) {                                                                   \
        case 1: CALL(1); break; case 2: CALL(2); break; case 3: CALL(3); break;    \
        case 4: CALL(4); break; case 5: CALL(5); break; case 6: CALL(6); break;    \
        case 7: CALL(7); break; case 8: CALL(8); break; case 9: CALL(9); break;    \
        case 10: CALL(10); break;                                                  \
        default: return set_error(PBN_ERR_UNSUPPORTED, "spatial order: unsupported dimension"); \
    }

int pbn_spatial_sort(pbn_ctx* ctx, int dtype, int d, int dk, const void* y, const double* nrm, int64_t n, const float* bound, void* ys,
                     double* nrm_s, int* perm) {
    if (dk != d && dk != d - 1) return set_error(PBN_ERR_ARG, "spatial order: key dimension must be d or d - 1");
    cudaStream_t st = ctx->stream;
    if (n == 0) return PBN_OK;
    unsigned long long *keys = nullptr, *keys2 = nullptr;
    int* idx = nullptr;
    void* tmp = nullptr;
    size_t tmp_bytes = 0;
    cudaError_t e = cudaMallocAsync(&keys, (size_t)n * 8 * 2 + (size_t)n * 4, st);
    if (e != cudaSuccess) PBN_CUDA_TRY(e);
    keys2 = keys + n;
    idx = reinterpret_cast<int*>(keys2 + n);
    const int blocks = (int)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * 16);
    const bool f64 = dtype == PBN_F64;
#define PBN_KEYS(DD)                                                                                                            \
    if (dk == DD) {                                                                                                             \
        if (f64) morton_key_kernel<double, DD, DD><<<blocks, 256, 0, st>>>(static_cast<const double*>(y), n, bound, keys, idx); \
        else morton_key_kernel<float, DD, DD><<<blocks, 256, 0, st>>>(static_cast<const float*>(y), n, bound, keys, idx);       \
    } else {                                                                                                                    \
        constexpr int DKK = DD > 1 ? DD - 1 : 1;                                                                                \
        if (f64) morton_key_kernel<double, DD, DKK><<<blocks, 256, 0, st>>>(static_cast<const double*>(y), n, bound, keys, idx);\
        else morton_key_kernel<float, DD, DKK><<<blocks, 256, 0, st>>>(static_cast<const float*>(y), n, bound, keys, idx);      \
    }
    PBN_D_SWITCH(PBN_KEYS)
#undef PBN_KEYS
    const int bits = (60 / dk > 16 ? 16 : 60 / dk) * dk;
    e = cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys2, idx, perm, (int)n, 0, bits, st);
    if (e == cudaSuccess) e = cudaMallocAsync(&tmp, tmp_bytes ? tmp_bytes : 8, st);
    if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys2, idx, perm, (int)n, 0, bits, st);
    if (e != cudaSuccess) {
        cudaFreeAsync(keys, st);
        if (tmp) cudaFreeAsync(tmp, st);
        PBN_CUDA_TRY(e);
    }
#define PBN_GATHER(DD)                                                                                                        \
    if (f64) gather_sorted_kernel<double, DD><<<blocks, 256, 0, st>>>(static_cast<const double*>(y), nrm, perm, n,            \
                                                                      static_cast<double*>(ys), nrm_s);                       \
    else gather_sorted_kernel<float, DD><<<blocks, 256, 0, st>>>(static_cast<const float*>(y), nrm, perm, n,                  \
                                                                 static_cast<float*>(ys), nrm_s)
    PBN_D_SWITCH(PBN_GATHER)
#undef PBN_GATHER
    ctx->launches += 4;
    cudaFreeAsync(keys, st);
    cudaFreeAsync(tmp, st);
    PBN_CUDA_TRY(cudaGetLastError());
    return PBN_OK;
}

int pbn_spatial_boxes(pbn_ctx* ctx, int dtype, int d, const void* ys, int64_t n, int tile_rows, float* box) {
    cudaStream_t st = ctx->stream;
    const int tiles = (int)((n + tile_rows - 1) / tile_rows);
    if (tiles == 0) return PBN_OK;
    const bool f64 = dtype == PBN_F64;
#define PBN_BOX(DD)                                                                                                       \
    if (f64) tile_box_kernel<double, DD><<<tiles, 256, 0, st>>>(static_cast<const double*>(ys), n, tile_rows, box);      \
    else tile_box_kernel<float, DD><<<tiles, 256, 0, st>>>(static_cast<const float*>(ys), n, tile_rows, box)
    PBN_D_SWITCH(PBN_BOX)
#undef PBN_BOX
    ctx->launches++;
    PBN_CUDA_TRY(cudaGetLastError());
    return PBN_OK;
}

int pbn_skip_nearest(pbn_ctx* ctx, const float* box_test, int n_test_tiles, const float* box_train, int n_train_tiles, int d, int K,
                     int* nearest, long long* first) {
    cudaStream_t st = ctx->stream;
    nearest_tile_kernel<<<n_test_tiles, 256, 0, st>>>(box_test, box_train, n_train_tiles, d, K, nearest);
    iota_kernel<<<(n_test_tiles + 256) / 256, 256, 0, st>>>(first, (long long)n_test_tiles + 1, K);
    ctx->launches += 2;
    PBN_CUDA_TRY(cudaGetLastError());
    return PBN_OK;
}

int pbn_skip_count(pbn_ctx* ctx, const pbn::PairJob* d_jobA, long long upbA, int tb, int ckde, int dtype, int64_t n_train,
                   const float* box_test, int n_test_tiles, const float* box_train, int n_train_tiles, int d, const int* nearest,
                   int K, float* thr, double* sumsA, long long* count, long long* tile_first, long long* total_out) {
    cudaStream_t st = ctx->stream;
    LbParams P;
    P.jobA = d_jobA;
    P.upbA = upbA;
    P.tb = tb;
    P.ckde = ckde;
    P.unit_per_binade = dtype == PBN_F64 ? (double)pbn::kExpTab : 1.0;
    P.log2_ntrain = log2((double)n_train);
    P.thr = thr;
    P.sums = sumsA;
    skip_threshold_kernel<<<n_test_tiles, 256, 0, st>>>(P);
    skip_list_kernel<0><<<n_test_tiles, 256, 0, st>>>(box_test, box_train, n_train_tiles, d, ckde, thr, nearest, K, count, nullptr, nullptr);
    scan_counts_kernel<<<1, 1024, 0, st>>>(count, n_test_tiles, tile_first);
    ctx->launches += 3;
    PBN_CUDA_TRY(cudaGetLastError());
    PBN_CUDA_TRY(cudaMemcpyAsync(total_out, tile_first + n_test_tiles, sizeof(long long), cudaMemcpyDeviceToHost, st));
    PBN_CUDA_TRY(cudaStreamSynchronize(st));
    ctx->d2h += 8;
    return PBN_OK;
}

int pbn_skip_fill(pbn_ctx* ctx, int ckde, const float* box_test, int n_test_tiles, const float* box_train, int n_train_tiles, int d,
                  const int* nearest, int K, const float* thr, const long long* tile_first, int* unit_list) {
    skip_list_kernel<1><<<n_test_tiles, 256, 0, ctx->stream>>>(box_test, box_train, n_train_tiles, d, ckde, thr, nearest, K, nullptr,
                                                               tile_first, unit_list);
    ctx->launches++;
    PBN_CUDA_TRY(cudaGetLastError());
    return PBN_OK;
}

int pbn_scatter_out(pbn_ctx* ctx, const double* src, const int* perm, int64_t n, double* dst) {
    if (n == 0) return PBN_OK;
    const int blocks = (int)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * 16);
    scatter_out_kernel<double><<<blocks, 256, 0, ctx->stream>>>(src, perm, n, dst);
    ctx->launches++;
    PBN_CUDA_TRY(cudaGetLastError());
    return PBN_OK;
}
