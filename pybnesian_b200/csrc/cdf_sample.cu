// cdf_sample.cu — CKDE::cdf and CKDE::sample on the device.
//
// Replaces (paths relative to /root/reference/pybnesian/):
//   factors/continuous/CKDE.hpp:506-728 (_cdf, _cdf_univariate, _cdf_multivariate) with kernels
//     univariate_normal_cdf, normal_cdf, conditional_means_*, exp_elementwise, product_elementwise,
//     division_elementwise (kde/opencl_kernels/KDE.cl.src:241-245, 366-468) and sum_cols_offset;
//   factors/continuous/CKDE.hpp:289-504 (_sample, _sample_multivariate, _sample_indices_from_weights) with
//     kernels exp_elementwise, accum_sum_mat_cols, add_accum_sum_mat_cols, normalize_accum_sum_mat_cols,
//     find_random_indices (KDE.cl.src:241-364).
// The reference materialises N x 64 matrices (weights, conditional means, cdf values, prefix sums) in global
// memory per chunk of 64 test rows; here every (train, test) pair lives in registers.
//
// Kernels: the cdf of families of up to 8 variables is the CDF mode of the fused pair kernel (pair_kernel.cuh:
// pair_kernel<T, D, CKDE, CDF = true>, same stream-K schedule and partial-sum slots as logl; its finalize is
// cdf_finalize_pair_kernel below).  weight_kernel below serves (MODE 0) the cdf of wider families with a run-time
// dimension, (MODE 1) the weight totals and (MODE 2) the index scan of the sampler.
//
// Both use the whitened layout of the fitted CKDE (runtime.cu: fit_impl): y = c L^-1 (x - mu) with the
// conditioned variable stored LAST.  Then, for a test row t and a training row i,
//   weight      w_ti  = exp(-1/2 |L_e^-1 (e_t - e_i)|^2)          = 2^(-(sum_{c<p} (yt_c - yi_c)^2) [/K])
//   cdf term    Phi((x_t - mean_ti) / sqrt(cond_var))             = Phi((yt_p - yi_p) / c)
// because the last row of L^-1 is (x - H_ve H_ee^-1 e) / sqrt(Schur complement): the conditional mean
// x_i + H_ve H_ee^-1 (e_t - e_i) and variance of CKDE.hpp:594-616 are already inside the whitening matrix.
//   cdf(t)   = sum_i w_ti Phi_ti / sum_i w_ti                       (no evidence: 1/N sum_i Phi_ti)
//   index(t) = first i with (sum_{j<=i} w_tj) > u_t sum_j w_tj,  N-1 if none   (find_random_indices); the sum is taken
//              hierarchically: split totals in split order, then the running sum inside the selected split
#include "internal.h"

#include <random>

namespace {

using namespace pbn;

constexpr int kWThreads = 256;
constexpr int kWTile = 128;  // training rows per shared-memory tile
constexpr int kWStages = 2;

__device__ __forceinline__ int64_t map_row2(int64_t r, int64_t b0, int64_t n0, int64_t b1) {
    return r < n0 ? b0 + r : b1 + (r - n0);
}

struct WParams {
    const void* train;  // whitened AoS [n (padded alloc)][d]
    const void* test;   // whitened AoS [m][dt]  (dt = d for the cdf, d - 1 for the sampling weights)
    long long n, m, m_pad;
    int d;              // joint dimension (row stride of train)
    int dt;             // row stride of test
    int n_train_tiles, tiles_per_split, n_splits;
    double inv_c;       // whitened (kernel-unit) coordinate -> standard normal unit
    double* part;       // [2][n_splits][m_pad]: sum w, sum w Phi
    const double* tab;  // exp2 table (f64)
    // scan pass
    const double* target;  // [m]: weight still to be accumulated INSIDE the selected split before the index is reached
    const int* sel_split;  // [m]: training split that holds the index of row t (-1: none, index N-1)
    int* idx;              // [m]
};

// Shared-memory pipeline of training tiles: thread 0 issues 1-D TMA bulk copies, everyone waits on the
// stage's mbarrier (same scheme as pair_kernel.cuh).
template <typename T>
struct TilePipe {
    T* buf;
    uint64_t* bar;
    const T* src;
    long long n;
    int d;
    int t_next;  // producer state (thread 0)
    uint32_t issued, consumed;
    __device__ __forceinline__ void issue() {
        int stage = issued % kWStages;
        long long start = static_cast<long long>(t_next) * kWTile;
        long long cnt = n - start;
        if (cnt > kWTile) cnt = kWTile;
        uint32_t bytes = static_cast<uint32_t>(((cnt * d * sizeof(T)) + 15) & ~15ull);
        mbar_expect_tx(&bar[stage], bytes);
        tma_bulk_g2s(buf + static_cast<size_t>(stage) * kWTile * d, src + start * d, bytes, &bar[stage]);
        ++t_next;
        ++issued;
    }
};

// MODE 0: cdf sums (sum w, sum w Phi) over the training tiles of split blockIdx.y
// MODE 1: weight sums only (sum w) - first pass of the index sampling
// MODE 2: index scan - CTA (test tile, split) walks the tiles of its split in order for the rows whose index lies in that
//         split (chosen from the split totals of MODE 1), and stops when each of them has its index
template <typename T, int DT, int MODE>
__global__ void __launch_bounds__(kWThreads) weight_kernel(const __grid_constant__ WParams P) {
    constexpr int DA = DT ? DT : PBN_MAX_DIM;
    const int D = DT ? DT : P.d;
    const int DN = D - 1;  // coordinates of the evidence (exponent)
    const int DTEST = MODE == 0 ? D : DN;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    T* tile_buf = reinterpret_cast<T*>(smem_raw);
    const size_t tile_bytes = static_cast<size_t>(kWStages) * kWTile * D * sizeof(T);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + ((tile_bytes + 15) & ~size_t(15)));
    double* tab = reinterpret_cast<double*>(smem_raw + ((tile_bytes + 15) & ~size_t(15)) + 64);

    const int tid = threadIdx.x;
    const long long row = static_cast<long long>(blockIdx.x) * kWThreads + tid;
    const bool ok = row < P.m;
    int t0 = blockIdx.y * P.tiles_per_split;
    int t1 = min(t0 + P.tiles_per_split, P.n_train_tiles);

    bool active = true;
    if (MODE == 2) {
        active = ok && P.sel_split[row] == static_cast<int>(blockIdx.y);
        if (!__syncthreads_or(active)) return;  // no row of this tile has its index in this split (nothing in flight yet)
    }
    if (sizeof(T) == 8 && (DN > 0 || MODE == 0)) exp_tab_fill(tab, P.tab, tid, kWThreads);
    if (tid == 0) {
        for (int s = 0; s < kWStages; ++s) mbar_init(&bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    TilePipe<T> pipe;
    pipe.buf = tile_buf;
    pipe.bar = bar;
    pipe.src = static_cast<const T*>(P.train);
    pipe.n = P.n;
    pipe.d = D;
    pipe.t_next = t0;
    pipe.issued = 0;
    pipe.consumed = 0;
    if (tid == 0)
        for (int s = 0; s < kWStages && pipe.t_next < t1; ++s) pipe.issue();

    T yt[DA];
    {
        const T* tp = static_cast<const T*>(P.test);
#pragma unroll
        for (int c = 0; c < DA; ++c)
            if (c < DTEST) yt[c] = ok ? tp[row * DTEST + c] : T(0);
    }
    double sw = 0.0, sp = 0.0;
    double target = 0.0;
    int found = -1;
    if (MODE == 2) {
        target = active ? P.target[row] : 0.0;
        if (!active) found = 0;
    }
    const T inv_c = static_cast<T>(P.inv_c);

    uint32_t it = 0;
    for (int t = t0; t < t1; ++t, ++it) {
        const int stage = it % kWStages;
        const uint32_t parity = (it / kWStages) & 1;
        long long cnt_ll = P.n - static_cast<long long>(t) * kWTile;
        const int cnt = cnt_ll > kWTile ? kWTile : static_cast<int>(cnt_ll);
        mbar_wait(&bar[stage], parity);
        const T* __restrict__ tp = tile_buf + static_cast<size_t>(stage) * kWTile * D;
        const int base = t * kWTile;

        if constexpr (sizeof(T) == 8) {
#pragma unroll 2
            for (int i = 0; i < cnt; ++i) {
                double acc = 0.0;
#pragma unroll
                for (int c = 0; c < DA; ++c)
                    if (c < DN) {
                        double dl = yt[c] - tp[i * D + c];
                        acc = fma(-dl, dl, acc);
                    }
                double w = 1.0;
                if (DN > 0) {
                    double st;
                    double pg = exp2_tab<true>(acc, tab, st);
                    w = __dmul_rn(st, pg);  // no contraction: MODE 1 and MODE 2 must produce the same running sums
                }
                if (MODE == 0) {
                    // w Phi(z) from the JOINT kernel value E = w exp(-z^2/2) = 2^(acc - dl^2) and the tail factor
                    // t g(u) (normal_tail_tg): one more table exp2, one MUFU reciprocal, one degree-18 Horner
                    double dl = yt[DN] - tp[i * D + DN];
                    double st2;
                    double pg2 = exp2_tab<true>(fma(-dl, dl, acc), tab, st2);
                    double q = (st2 * pg2) * normal_tail_tg(fabs(dl) * inv_c);
                    sp += (dl < 0.0) ? q : (w - q);
                }
                sw = __dadd_rn(sw, w);
                if (MODE == 2) {
                    if (found < 0 && sw > target) found = base + i;
                }
            }
        } else {
            // per-tile float partial sums folded into the double accumulators
            float fw = 0.f, fp = 0.f;
            float trel = 0.f;
            if (MODE == 2) {
                double rel = target - sw;  // remaining weight before the target is crossed
                trel = rel > 3.0e38 ? INFINITY : static_cast<float>(rel);
                // round towards the side that keeps "fw > trel" equivalent to "sw + fw > target" up to float rounding
            }
#pragma unroll 2
            for (int i = 0; i < cnt; ++i) {
                float acc = 0.f;
#pragma unroll
                for (int c = 0; c < DA; ++c)
                    if (c < DN) {
                        float dl = yt[c] - tp[i * D + c];
                        acc = fmaf(-dl, dl, acc);
                    }
                float w = 1.f;
                if (DN > 0) asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(w) : "f"(acc));
                if (MODE == 0) {
                    float dl = yt[DN] - tp[i * D + DN];
                    float e;
                    float accj = fmaf(-dl, dl, acc);
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(accj));
                    float q = e * normal_tail_tg_f(fabsf(dl) * inv_c);
                    fp += (dl < 0.f) ? q : (w - q);
                }
                fw += w;
                if (MODE == 2) {
                    if (found < 0 && fw > trel) found = base + i;
                }
            }
            sw += static_cast<double>(fw);
            sp += static_cast<double>(fp);
        }

        int done = 0;
        if (MODE == 2) done = __syncthreads_and(found >= 0);
        else __syncthreads();
        if (tid == 0) {
            pipe.consumed = it + 1;
            if (!done && pipe.t_next < t1) pipe.issue();
        }
        if (MODE == 2 && done) break;
    }
    // never leave the CTA with a bulk copy still in flight into its shared memory
    if (MODE == 2 && tid == 0) {
        for (uint32_t k = pipe.consumed; k < pipe.issued; ++k) mbar_wait(&bar[k % kWStages], (k / kWStages) & 1);
    }

    if (!ok) return;
    if (MODE == 2) {
        // the split was selected with the same predicate on the same running sums, so an active row always finds its
        // index; the last row of the split is the defensive default
        if (active) {
            long long last = static_cast<long long>(t1) * kWTile;
            if (last > P.n) last = P.n;
            P.idx[row] = found >= 0 ? found : static_cast<int>(last - 1);
        }
    } else {
        P.part[static_cast<long long>(blockIdx.y) * P.m_pad + row] = sw;
        if (MODE == 0) P.part[(static_cast<long long>(P.n_splits) + blockIdx.y) * P.m_pad + row] = sp;
    }
}

struct WFinal {
    const double* part;
    long long m, m_pad, n;
    int n_splits;
    int has_evidence;
    double thresh;
    double* out;       // cdf: [m]
    int* flagged;      // cdf: rows needing the reference-arithmetic path
    int* n_flagged;
    const void* u;     // sampling: uniform draws in the data's dtype
    int u_f64;
    double* target;    // sampling: weight left to accumulate inside the selected split
    int* sel_split;    // sampling: selected split per row (-1: none)
    int* idx;          // sampling: preset to N - 1, the reference's default
};

__global__ void cdf_finalize_kernel(WFinal F) {
    long long row = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (row >= F.m) return;
    double sw = 0, sp = 0;
    for (int s = 0; s < F.n_splits; ++s) {
        sw += F.part[(long long)s * F.m_pad + row];
        sp += F.part[((long long)F.n_splits + s) * F.m_pad + row];
    }
    if (!F.has_evidence) {
        F.out[row] = sp / (double)F.n;
        return;
    }
    if (sw == sw && !(sw >= F.thresh)) {
        int slot = atomicAdd(F.n_flagged, 1);
        F.flagged[slot] = (int)row;
        return;
    }
    F.out[row] = sp / sw;
}

// Same epilogue for the fused pair kernel's partial-sum slots (stream-K: one slot per CTA that touched the test tile,
// added in fixed order; see finalize_kernel in runtime.cu).  Slot block 0 holds sum w Phi, block 1 holds sum w.
struct CdfPairFinal {
    const PairJob* job;
    long long upb;
    int tb;
    int has_evidence;
    double thresh;
    long long n;
    double* out;
    int* flagged;
    int* n_flagged;
};

__global__ void cdf_finalize_pair_kernel(CdfPairFinal F) {
    const PairJob jb = *F.job;
    long long row = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (row >= jb.m) return;
    long long tt = row / F.tb;
    long long ustart = jb.unit_begin + tt * jb.n_train_tiles;
    int first = (int)(ustart / F.upb);
    int last = (int)((ustart + jb.n_train_tiles - 1) / F.upb);
    int ns = last - first + 1;
    double sp = 0, sw = 0;
    for (int s = 0; s < ns; ++s) {
        sp += jb.part[(long long)s * jb.m_pad + row];
        if (F.has_evidence) sw += jb.part[((long long)jb.slots + s) * jb.m_pad + row];
    }
    if (!F.has_evidence) {
        F.out[row] = sp / (double)F.n;
        return;
    }
    if (sw == sw && !(sw >= F.thresh)) {
        int slot = atomicAdd(F.n_flagged, 1);
        F.flagged[slot] = (int)row;
        return;
    }
    F.out[row] = sp / sw;
}

__global__ void write_cdf_job_kernel(PairJob j, PairJob* dst, int* zero_counter) {
    *dst = j;
    *zero_counter = 0;
}

__global__ void sample_target_kernel(WFinal F) {
    long long row = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (row >= F.m) return;
    double sw = 0;
    for (int s = 0; s < F.n_splits; ++s) sw += F.part[(long long)s * F.m_pad + row];
    double u = F.u_f64 ? static_cast<const double*>(F.u)[row] : (double)static_cast<const float*>(F.u)[row];
    // the index is the first training row whose running weight sum exceeds u * S: find the split it lies in from the
    // split totals (added in split order), keep what is left to accumulate inside that split.  Weights that underflow
    // (S below the threshold), NaN evidence and u * S >= S by rounding leave the reference's default, row N - 1.
    int sel = -1;
    double rest = 0.0;
    if (sw >= F.thresh) {
        const double target = u * sw;
        double prefix = 0.0;
        for (int s = 0; s < F.n_splits; ++s) {
            const double ps = F.part[(long long)s * F.m_pad + row];
            if (ps > target - prefix) {
                sel = s;
                rest = target - prefix;
                break;
            }
            prefix += ps;
        }
    }
    F.target[row] = rest;
    F.sel_split[row] = sel;
    F.idx[row] = static_cast<int>(F.n - 1);
}

// Rows whose unshifted weight sum underflowed: the reference's arithmetic, term by term (weights
// exp(-s/2 + lognorm_marg + log N) rounded to the data type, NOT max-shifted: 0/0 = NaN when they all vanish).
struct CdfRowParams {
    const void* train;
    const void* test;
    long long n;
    int d;
    double u2ln, c0, inv_c;
    const int* rows;
    const int* count_ptr;
    double* out;
};

__device__ __forceinline__ double block_sum2(double v, double* sh) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    double r = 0;
    if (threadIdx.x == 0)
        for (int i = 0; i < (int)((blockDim.x + 31) >> 5); ++i) r += sh[i];
    return r;
}

template <typename T>
__global__ void cdf_row_kernel(CdfRowParams P) {
    __shared__ double sh[32];
    __shared__ double yt[PBN_MAX_DIM];
    const int cnt = *P.count_ptr;
    const T* tr = static_cast<const T*>(P.train);
    const T* te = static_cast<const T*>(P.test);
    const int d = P.d;
    for (int f = blockIdx.x; f < cnt; f += gridDim.x) {
        long long row = P.rows[f];
        __syncthreads();
        if (threadIdx.x < d) yt[threadIdx.x] = static_cast<double>(te[row * d + threadIdx.x]);
        __syncthreads();
        double den = 0, num = 0;
        for (long long i = threadIdx.x; i < P.n; i += blockDim.x) {
            double s = 0;
            for (int c = 0; c < d - 1; ++c) {
                double dl = yt[c] - static_cast<double>(tr[i * d + c]);
                s = fma(dl, dl, s);
            }
            double e = -s * P.u2ln + P.c0;
            double w = sizeof(T) == 8 ? exp(e) : static_cast<double>(expf(static_cast<float>(e)));
            double z = (yt[d - 1] - static_cast<double>(tr[i * d + d - 1])) * P.inv_c;
            den += w;
            num += w * normcdf(z);
        }
        double tden = block_sum2(den, sh);
        double tnum = block_sum2(num, sh);
        if (threadIdx.x == 0) P.out[row] = tnum / tden;
    }
}

__global__ void zero_int_kernel(int* p) { *p = 0; }

// raw training values of the sampled rows: out[j * n + i] = column j at training row idx[i]
template <typename T>
__global__ void gather_train_kernel(ColPtrs cols, int d, int64_t b0, int64_t n0, int64_t b1, const int* __restrict__ idx,
                                    long long n, T* __restrict__ out) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t rr = map_row2(idx[i], b0, n0, b1);
    for (int j = 0; j < d; ++j) out[(long long)j * n + i] = static_cast<const T*>(cols.p[j])[rr];
}

size_t weight_smem(int d, size_t es, bool f64) {
    size_t tile_bytes = (size_t)kWStages * kWTile * d * es;
    return ((tile_bytes + 15) & ~size_t(15)) + 64 + (f64 ? exp_tab_smem_bytes<double>() : 0);
}

template <typename T, int DT, int MODE>
cudaError_t launch_weight_one(const WParams& P, dim3 grid, cudaStream_t st) {
    size_t smem = weight_smem(P.d, sizeof(T), sizeof(T) == 8);
    auto kern = weight_kernel<T, DT, MODE>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    kern<<<grid, kWThreads, smem, st>>>(P);
    return cudaGetLastError();
}

template <typename T, int MODE>
cudaError_t launch_weight(const WParams& P, dim3 grid, cudaStream_t st) {
    if constexpr (MODE == 0) {
        // families of up to kMaxFastD (10) variables take the CDF mode of the fused pair kernel (pair_kernel.cuh)
        return launch_weight_one<T, 0, MODE>(P, grid, st);
    } else {
        switch (P.d) {
            case 2: return launch_weight_one<T, 2, MODE>(P, grid, st);
            case 3: return launch_weight_one<T, 3, MODE>(P, grid, st);
            case 4: return launch_weight_one<T, 4, MODE>(P, grid, st);
            case 5: return launch_weight_one<T, 5, MODE>(P, grid, st);
            case 6: return launch_weight_one<T, 6, MODE>(P, grid, st);
            case 7: return launch_weight_one<T, 7, MODE>(P, grid, st);
            case 8: return launch_weight_one<T, 8, MODE>(P, grid, st);
            default: return launch_weight_one<T, 0, MODE>(P, grid, st);  // runtime dimension (9..32)
        }
    }
}

// split of the training tiles so that (test tiles x splits) fills the GPU
void plan_splits(pbn_ctx* ctx, int64_t n, int64_t m, WParams& P) {
    P.n_train_tiles = (int)((n + kWTile - 1) / kWTile);
    int n_test_tiles = (int)((m + kWThreads - 1) / kWThreads);
    int want = (2 * ctx->sm_count + n_test_tiles - 1) / n_test_tiles;
    int splits = std::max(1, std::min(std::min(want, 64), P.n_train_tiles));
    P.tiles_per_split = (P.n_train_tiles + splits - 1) / splits;
    P.n_splits = (P.n_train_tiles + P.tiles_per_split - 1) / P.tiles_per_split;
}

// whitened test rows under the CKDE's whitening matrix: all d columns (cdf) or the evidence block (sampling)
int whiten_test(pbn_ctx* ctx, const pbn_kde* k, const pbn_table* tbl, const int* cols_internal, int dt, pbn_rows rows,
                void* out) {
    const int d = k->d;
    std::vector<double> W((size_t)dt * dt);
    for (int i = 0; i < dt; ++i)
        for (int j = 0; j < dt; ++j) W[(size_t)i * dt + j] = k->W[i * d + j];
    return whiten_raw_launch(ctx, tbl, cols_internal, dt, rows, W.data(), k->mu, out, nullptr, nullptr, 0);
}

// host-side quantities of CKDE::_sample_multivariate (CKDE.hpp:346-360) from the joint bandwidth (column-major,
// variable first): transform = H_ve H_ee^-1 through chol(H_ee), cond_var = H_vv - |L_e^-1 H_ev|^2
bool cond_params(const double* H, int d, std::vector<double>& transform, double& cond_var) {
    const int p = d - 1;
    std::vector<double> Hm((size_t)p * p), L((size_t)p * p), inv((size_t)p * p);
    for (int i = 0; i < p; ++i)
        for (int j = 0; j < p; ++j) Hm[i + (size_t)j * p] = H[(i + 1) + (size_t)(j + 1) * d];
    if (!chol_lower(Hm.data(), p, L.data())) return false;
    tri_inverse_rowmajor(L.data(), p, inv.data());  // inv[i * p + j]
    std::vector<double> R(p, 0.0);
    for (int i = 0; i < p; ++i)
        for (int k = 0; k < p; ++k) R[i] += inv[(size_t)i * p + k] * H[(k + 1)];
    double nrm = 0;
    for (int i = 0; i < p; ++i) nrm += R[i] * R[i];
    cond_var = H[0] - nrm;
    transform.assign(p, 0.0);
    for (int j = 0; j < p; ++j)
        for (int i = 0; i < p; ++i) transform[j] += R[i] * inv[(size_t)i * p + j];
    return true;
}

int sample_indices_device(pbn_ctx* ctx, const pbn_kde* k, const pbn_table* ev, const int* ev_cols, pbn_rows rows,
                          const void* h_random_prob, int* d_idx) {
    cudaStream_t st = ctx->stream;
    const int d = k->d, p = d - 1;
    const int64_t m = seg_count(rows);
    const bool f64 = k->dtype == PBN_F64;
    const size_t es = elem_size(k->dtype);
    void* ytest = nullptr;
    PBN_CUDA_TRY(cudaMallocAsync(&ytest, std::max<size_t>((size_t)m * p * es, 16), st));
    PBN_TRY(whiten_test(ctx, k, ev, ev_cols, p, rows, ytest));
    WParams P;
    memset(&P, 0, sizeof(P));
    P.train = k->y;
    P.test = ytest;
    P.n = k->n;
    P.m = m;
    P.m_pad = (m + 31) / 32 * 32;
    P.d = d;
    P.dt = p;
    P.inv_c = 1.0 / sqrt(0.5 * unit_scale(k->dtype));
    P.tab = ctx->d_exp_tab;
    plan_splits(ctx, k->n, m, P);
    double* part = nullptr;
    void* d_u = nullptr;
    double* target = nullptr;
    int* sel_split = nullptr;
    PBN_CUDA_TRY(cudaMallocAsync(&part, (size_t)P.n_splits * P.m_pad * sizeof(double), st));
    PBN_CUDA_TRY(cudaMallocAsync(&d_u, (size_t)m * es, st));
    PBN_CUDA_TRY(cudaMallocAsync(&target, (size_t)m * sizeof(double), st));
    PBN_CUDA_TRY(cudaMallocAsync(&sel_split, (size_t)m * sizeof(int), st));
    PBN_CUDA_TRY(cudaMemcpyAsync(d_u, h_random_prob, (size_t)m * es, cudaMemcpyHostToDevice, st));
    ctx->h2d += (int64_t)(m * es);
    P.part = part;
    dim3 grid((unsigned)((m + kWThreads - 1) / kWThreads), (unsigned)P.n_splits);
    {
        cudaError_t le = f64 ? launch_weight<double, 1>(P, grid, st) : launch_weight<float, 1>(P, grid, st);
        PBN_CUDA_TRY(le);
    }
    ctx->launches++;
    WFinal F;
    memset(&F, 0, sizeof(F));
    F.part = part;
    F.m = m;
    F.m_pad = P.m_pad;
    F.n = k->n;
    F.n_splits = P.n_splits;
    // above n_train clamped weights (exp2_tab evaluates an out-of-range term as 2^-kExpMinK): 2^25 rows
    F.thresh = f64 ? ldexp(1.0, -pbn::kExpMinK + 25) : ldexp(1.0, -120);
    F.u = d_u;
    F.u_f64 = f64 ? 1 : 0;
    F.target = target;
    F.sel_split = sel_split;
    F.idx = d_idx;
    sample_target_kernel<<<(int)((m + 255) / 256), 256, 0, st>>>(F);
    ctx->launches++;
    PBN_CUDA_TRY(cudaGetLastError());
    P.target = target;
    P.sel_split = sel_split;
    P.idx = d_idx;
    dim3 grid2((unsigned)((m + kWThreads - 1) / kWThreads), (unsigned)P.n_splits);
    {
        cudaError_t le = f64 ? launch_weight<double, 2>(P, grid2, st) : launch_weight<float, 2>(P, grid2, st);
        PBN_CUDA_TRY(le);
    }
    ctx->launches++;
    PBN_CUDA_TRY(cudaFreeAsync(part, st));
    PBN_CUDA_TRY(cudaFreeAsync(d_u, st));
    PBN_CUDA_TRY(cudaFreeAsync(target, st));
    PBN_CUDA_TRY(cudaFreeAsync(sel_split, st));
    PBN_CUDA_TRY(cudaFreeAsync(ytest, st));
    return PBN_OK;
}

int check_ckde(const pbn_kde* k) {
    if (!k) return set_error(PBN_ERR_ARG, "null argument");
    if (k->d >= 2 && !k->ckde) return set_error(PBN_ERR_ARG, "not a CKDE: fit it with pbn_ckde_fit");
    return PBN_OK;
}

template <typename T>
void sample_host_multivariate(const T* gathered /* [d][n]: variable, evidence.. at the sampled rows */,
                              const T* const* ev_host, int p, int64_t n, const std::vector<double>& transform,
                              double cond_var, std::mt19937& rng, T* out) {
    std::vector<T> tr(p);
    for (int j = 0; j < p; ++j) tr[j] = static_cast<T>(transform[j]);
    std::normal_distribution<T> normal(0, std::sqrt(cond_var));
    for (int64_t i = 0; i < n; ++i) {
        T acc = 0;
        for (int j = 0; j < p; ++j) acc += (ev_host[j][i] - gathered[(size_t)(j + 1) * n + i]) * tr[j];
        acc += gathered[i] + normal(rng);
        out[i] = acc;
    }
}

}  // namespace

extern "C" {

static int ckde_cdf_one(pbn_ctx* ctx, const pbn_kde* k, const pbn_table* test, const int* cols, pbn_rows rows, double* out);

int pbn_ckde_cdf(pbn_ctx* ctx, const pbn_kde* k, const pbn_table* test, const int* cols, pbn_rows rows, double* out) {
    if (!ctx || !out) return set_error(PBN_ERR_ARG, "null argument");
    const int nd = pbn_num_devices(ctx);
    const int64_t m = (test && rows.e0 >= rows.b0 && rows.e1 >= rows.b1) ? seg_count(rows) : 0;
    if (!k || !pbn_replicated(ctx, k) || !pbn_replicated(ctx, test) || m < (int64_t)2048 * nd || (double)k->n * (double)m < 1.0e9 * nd)
        return ckde_cdf_one(ctx, k, test, cols, rows, out);
    PBN_TRY(check_cols(test, cols, k->d));
    PBN_TRY(check_rows(test, rows));
    // multi-device context: contiguous shards of the test rows against the replicated model (as pbn_kde_logl)
    return pbn_run_on_devices(nd, [&](int i) {
        const int64_t b = m * i / nd, e = m * (i + 1) / nd;
        return ckde_cdf_one(pbn_device_ctx(ctx, i), pbn_replica(const_cast<pbn_kde*>(k), i), pbn_replica(const_cast<pbn_table*>(test), i),
                            cols, pbn_sub_rows(rows, b, e), out + b);
    });
}

static int ckde_cdf_one(pbn_ctx* ctx, const pbn_kde* k, const pbn_table* test, const int* cols, pbn_rows rows, double* out) {
    if (!ctx || !out) return set_error(PBN_ERR_ARG, "null argument");
    PBN_TRY(check_ckde(k));
    PBN_TRY(check_cols(test, cols, k->d));
    PBN_TRY(check_rows(test, rows));
    if (test->dtype != k->dtype) return set_error(PBN_ERR_ARG, "Data type of training and test datasets is different.");
    DevSetter ds(ctx->device);
    cudaStream_t st = ctx->stream;
    const int d = k->d;
    const int64_t m = seg_count(rows);
    if (m == 0) return PBN_OK;
    const bool f64 = k->dtype == PBN_F64;
    const size_t es = elem_size(k->dtype);
    const double inv_c = 1.0 / sqrt(0.5 * unit_scale(k->dtype));
    const bool fast = d <= pbn::kMaxFastD;

    void* ytest = nullptr;
    size_t ytbytes = ((size_t)m * d * es + 255) / 256 * 256;
    const size_t tnbytes = (fast && k->nrm) ? ((size_t)m * sizeof(double) + 255) / 256 * 256 : 0;
    PBN_CUDA_TRY(cudaMallocAsync(&ytest, ytbytes + 256 + tnbytes, st));
    float* bound_test = reinterpret_cast<float*>(static_cast<char*>(ytest) + ytbytes);
    double* nrm_test = tnbytes ? reinterpret_cast<double*>(static_cast<char*>(ytest) + ytbytes + 256) : nullptr;
    PBN_CUDA_TRY(cudaMemsetAsync(bound_test, 0, 256, st));
    PBN_TRY(pbn_whiten_kde(ctx, k, test, cols, rows, ytest, bound_test, nrm_test));

    double* d_out = nullptr;
    int* flagged = nullptr;
    PBN_CUDA_TRY(cudaMallocAsync(&d_out, (size_t)m * sizeof(double), st));
    PBN_CUDA_TRY(cudaMallocAsync(&flagged, ((size_t)m + 1) * sizeof(int), st));
    int* n_flagged = flagged + m;
    double* part = nullptr;
    PairJob* d_job = nullptr;

    if (fast) {
        // CDF mode of the fused pair kernel: same stream-K schedule and partial-sum slots as pbn_logl_impl
        const int TILE = f64 ? pbn::pair_tile_f64(d) : pbn::pair_tile_f32(d);
        const int TB = f64 ? pbn::pair_tb_cdf_f64(d) : pbn::pair_tb_cdf_f32(d);  // CDF mode: pair_rows_cdf
        int n_test_tiles = (int)((m + TB - 1) / TB);
        int n_train_tiles = (int)((k->n + TILE - 1) / TILE);
        long long U = (long long)n_test_tiles * n_train_tiles;
        int grid = (int)std::min<long long>(U, (long long)ctx->sm_count * (f64 ? pbn::pair_ctas_per_sm_f64() : pbn::pair_ctas_per_sm_f32()));
        long long upb = (U + grid - 1) / grid;
        grid = (int)((U + upb - 1) / upb);
        int slots = (int)std::min<long long>((n_train_tiles + upb - 1) / upb + 1, grid);
        long long m_pad = (m + 31) / 32 * 32;
        PBN_CUDA_TRY(cudaMallocAsync(&part, (size_t)2 * slots * m_pad * sizeof(double), st));
        PBN_CUDA_TRY(cudaMallocAsync(&d_job, sizeof(PairJob) + 16, st));
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
        write_cdf_job_kernel<<<1, 1, 0, st>>>(job, d_job, n_flagged);
        ctx->launches++;
        PBN_CUDA_TRY(cudaGetLastError());
        cudaEvent_t ev0 = nullptr, ev1 = nullptr;
        if (ctx->timing) {
            PBN_CUDA_TRY(cudaEventCreate(&ev0));
            PBN_CUDA_TRY(cudaEventCreate(&ev1));
            PBN_CUDA_TRY(cudaEventRecord(ev0, st));
        }
        cudaError_t le = f64 ? pbn::launch_cdf_f64(d, d_job, 1, U, upb, grid, ctx->d_exp_tab, inv_c, st)
                             : pbn::launch_cdf_f32(d, d_job, 1, U, upb, grid, ctx->d_exp_tab, inv_c, st);
        ctx->launches++;
        PBN_CUDA_TRY(le);
        if (ctx->timing) {
            PBN_CUDA_TRY(cudaEventRecord(ev1, st));
            ctx->timed.emplace_back(ev0, ev1);
            ctx->pair_units += (int64_t)k->n * m;
        }
        CdfPairFinal F;
        F.job = d_job;
        F.upb = upb;
        F.tb = TB;
        F.has_evidence = d > 1 ? 1 : 0;
        F.thresh = f64 ? ldexp(1.0, -900) : ldexp(1.0, -64);
        F.n = k->n;
        F.out = d_out;
        F.flagged = flagged;
        F.n_flagged = n_flagged;
        cdf_finalize_pair_kernel<<<(int)((m + 255) / 256), 256, 0, st>>>(F);
        ctx->launches++;
        PBN_CUDA_TRY(cudaGetLastError());
    } else {
        // wide families (9..32 variables): runtime-dimension weight kernel, test tiles x training splits
        WParams P;
        memset(&P, 0, sizeof(P));
        P.train = k->y;
        P.test = ytest;
        P.n = k->n;
        P.m = m;
        P.m_pad = (m + 31) / 32 * 32;
        P.d = d;
        P.dt = d;
        P.inv_c = inv_c;
        P.tab = ctx->d_exp_tab;
        plan_splits(ctx, k->n, m, P);
        PBN_CUDA_TRY(cudaMallocAsync(&part, (size_t)2 * P.n_splits * P.m_pad * sizeof(double), st));
        zero_int_kernel<<<1, 1, 0, st>>>(n_flagged);
        ctx->launches++;
        P.part = part;
        dim3 grid((unsigned)((m + kWThreads - 1) / kWThreads), (unsigned)P.n_splits);
        {
            cudaError_t le = f64 ? launch_weight<double, 0>(P, grid, st) : launch_weight<float, 0>(P, grid, st);
            PBN_CUDA_TRY(le);
        }
        ctx->launches++;
        WFinal F;
        memset(&F, 0, sizeof(F));
        F.part = part;
        F.m = m;
        F.m_pad = P.m_pad;
        F.n = k->n;
        F.n_splits = P.n_splits;
        F.has_evidence = 1;
        F.thresh = f64 ? ldexp(1.0, -900) : ldexp(1.0, -64);
        F.out = d_out;
        F.flagged = flagged;
        F.n_flagged = n_flagged;
        cdf_finalize_kernel<<<(int)((m + 255) / 256), 256, 0, st>>>(F);
        ctx->launches++;
        PBN_CUDA_TRY(cudaGetLastError());
    }
    if (d > 1) {
        // rows whose unshifted weight sum underflowed: the reference's arithmetic, term by term
        CdfRowParams R;
        R.train = k->y;
        R.test = ytest;
        R.n = k->n;
        R.d = d;
        R.u2ln = 1.0 / unit_scale(k->dtype);
        R.c0 = k->lognorm_marg + log((double)k->n);
        R.inv_c = inv_c;
        R.rows = flagged;
        R.count_ptr = n_flagged;
        R.out = d_out;
        int rgrid = (int)std::min<int64_t>(m, (int64_t)ctx->sm_count * 8);
        if (f64) cdf_row_kernel<double><<<rgrid, 256, 0, st>>>(R);
        else cdf_row_kernel<float><<<rgrid, 256, 0, st>>>(R);
        ctx->launches++;
        PBN_CUDA_TRY(cudaGetLastError());
    }
    int nf = 0;
    PBN_CUDA_TRY(cudaMemcpyAsync(out, d_out, (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, st));
    PBN_CUDA_TRY(cudaMemcpyAsync(&nf, n_flagged, sizeof(int), cudaMemcpyDeviceToHost, st));
    PBN_CUDA_TRY(cudaStreamSynchronize(st));
    ctx->d2h += m * 8 + 4;
    ctx->last_fallback_rows = nf;
    PBN_CUDA_TRY(cudaFreeAsync(part, st));
    if (d_job) PBN_CUDA_TRY(cudaFreeAsync(d_job, st));
    PBN_CUDA_TRY(cudaFreeAsync(d_out, st));
    PBN_CUDA_TRY(cudaFreeAsync(flagged, st));
    PBN_CUDA_TRY(cudaFreeAsync(ytest, st));
    return PBN_OK;
}

int pbn_ckde_sample_indices(pbn_ctx* ctx, const pbn_kde* k, const pbn_table* evidence, const int* ev_cols, pbn_rows rows,
                            const void* random_prob, int32_t* out_idx) {
    if (!ctx || !random_prob || !out_idx) return set_error(PBN_ERR_ARG, "null argument");
    PBN_TRY(check_ckde(k));
    if (k->d < 2) return set_error(PBN_ERR_ARG, "a CKDE without evidence has no sampling weights");
    PBN_TRY(check_cols(evidence, ev_cols, k->d - 1));
    PBN_TRY(check_rows(evidence, rows));
    if (evidence->dtype != k->dtype)
        return set_error(PBN_ERR_ARG, "Data type of evidence values is different from CKDE training data.");
    DevSetter ds(ctx->device);
    const int64_t m = seg_count(rows);
    if (m == 0) return PBN_OK;
    int* d_idx = nullptr;
    PBN_CUDA_TRY(cudaMallocAsync(&d_idx, (size_t)m * sizeof(int), ctx->stream));
    PBN_TRY(sample_indices_device(ctx, k, evidence, ev_cols, rows, random_prob, d_idx));
    PBN_CUDA_TRY(cudaMemcpyAsync(out_idx, d_idx, (size_t)m * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    PBN_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    ctx->d2h += m * 4;
    PBN_CUDA_TRY(cudaFreeAsync(d_idx, ctx->stream));
    return PBN_OK;
}

int pbn_ckde_sample(pbn_ctx* ctx, const pbn_kde* k, const double* H, const pbn_table* train, const int* train_cols,
                    pbn_rows train_rows, const pbn_table* evidence, const int* ev_cols, const void* const* ev_host,
                    int64_t n, uint32_t seed, void* out, int32_t* idx_out) {
    if (!ctx || !H || !out) return set_error(PBN_ERR_ARG, "null argument");
    PBN_TRY(check_ckde(k));
    if (n < 0) return set_error(PBN_ERR_ARG, "n should be a non-negative number");
    PBN_TRY(check_cols(train, train_cols, k->d));
    PBN_TRY(check_rows(train, train_rows));
    if (seg_count(train_rows) != k->n || train->dtype != k->dtype)
        return set_error(PBN_ERR_ARG, "training table does not match the fitted CKDE");
    if (n == 0) return PBN_OK;
    DevSetter ds(ctx->device);
    cudaStream_t st = ctx->stream;
    const int d = k->d;
    const bool f64 = k->dtype == PBN_F64;
    const size_t es = elem_size(k->dtype);
    if (d == 1) {
        // CKDE::_sample without evidence (CKDE.hpp:295-316)
        std::vector<char> col((size_t)k->n * es);
        PBN_TRY(pbn_table_download(ctx, train, train_cols[0], train_rows, col.data()));
        std::mt19937 rng{seed};
        std::uniform_int_distribution<> uniform(0, (int)(k->n - 1));
        if (f64) {
            std::normal_distribution<double> normal(0, std::sqrt(H[0]));
            const double* x = reinterpret_cast<const double*>(col.data());
            for (int64_t i = 0; i < n; ++i) {
                auto index = uniform(rng);
                if (idx_out) idx_out[i] = index;
                static_cast<double*>(out)[i] = x[index] + normal(rng);
            }
        } else {
            std::normal_distribution<float> normal(0, std::sqrt(H[0]));
            const float* x = reinterpret_cast<const float*>(col.data());
            for (int64_t i = 0; i < n; ++i) {
                auto index = uniform(rng);
                if (idx_out) idx_out[i] = index;
                static_cast<float*>(out)[i] = x[index] + normal(rng);
            }
        }
        return PBN_OK;
    }
    const int p = d - 1;
    if (!ev_host) return set_error(PBN_ERR_ARG, "null evidence columns");
    PBN_TRY(check_cols(evidence, ev_cols, p));
    if (evidence->dtype != k->dtype)
        return set_error(PBN_ERR_ARG, "Data type of evidence values is different from CKDE training data.");
    if (evidence->nrows < n) return set_error(PBN_ERR_ARG, "evidence table has fewer than n rows");
    std::vector<double> transform;
    double cond_var = 0;
    if (!cond_params(H, d, transform, cond_var))
        return set_error(PBN_ERR_SINGULAR, "marginal bandwidth matrix is not positive definite");
    // uniform draws first, from the same generator that later draws the normals (CKDE.hpp:336-341, 386-392)
    std::mt19937 rng{seed};
    std::vector<char> rp((size_t)n * es);
    if (f64) {
        std::uniform_real_distribution<double> uniform(0, 1);
        for (int64_t i = 0; i < n; ++i) reinterpret_cast<double*>(rp.data())[i] = uniform(rng);
    } else {
        std::uniform_real_distribution<float> uniform(0, 1);
        for (int64_t i = 0; i < n; ++i) reinterpret_cast<float*>(rp.data())[i] = uniform(rng);
    }
    int* d_idx = nullptr;
    void* d_gather = nullptr;
    PBN_CUDA_TRY(cudaMallocAsync(&d_idx, (size_t)n * sizeof(int), st));
    PBN_CUDA_TRY(cudaMallocAsync(&d_gather, (size_t)n * d * es, st));
    pbn_rows ev_rows{0, n, 0, 0};
    PBN_TRY(sample_indices_device(ctx, k, evidence, ev_cols, ev_rows, rp.data(), d_idx));
    ColPtrs cp;
    for (int j = 0; j < d; ++j) cp.p[j] = col_ptr(train, train_cols[j]);
    int blocks = (int)((n + 255) / 256);
    if (f64)
        gather_train_kernel<double><<<blocks, 256, 0, st>>>(cp, d, train_rows.b0, train_rows.e0 - train_rows.b0, train_rows.b1,
                                                            d_idx, n, static_cast<double*>(d_gather));
    else
        gather_train_kernel<float><<<blocks, 256, 0, st>>>(cp, d, train_rows.b0, train_rows.e0 - train_rows.b0, train_rows.b1,
                                                           d_idx, n, static_cast<float*>(d_gather));
    ctx->launches++;
    PBN_CUDA_TRY(cudaGetLastError());
    std::vector<char> gathered((size_t)n * d * es);
    std::vector<int> idx((size_t)n);
    PBN_CUDA_TRY(cudaMemcpyAsync(gathered.data(), d_gather, gathered.size(), cudaMemcpyDeviceToHost, st));
    PBN_CUDA_TRY(cudaMemcpyAsync(idx.data(), d_idx, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, st));
    PBN_CUDA_TRY(cudaStreamSynchronize(st));
    ctx->d2h += (int64_t)(gathered.size() + n * 4);
    PBN_CUDA_TRY(cudaFreeAsync(d_idx, st));
    PBN_CUDA_TRY(cudaFreeAsync(d_gather, st));
    if (idx_out) std::copy(idx.begin(), idx.end(), idx_out);
    if (f64)
        sample_host_multivariate<double>(reinterpret_cast<const double*>(gathered.data()),
                                         reinterpret_cast<const double* const*>(ev_host), p, n, transform, cond_var, rng,
                                         static_cast<double*>(out));
    else
        sample_host_multivariate<float>(reinterpret_cast<const float*>(gathered.data()),
                                        reinterpret_cast<const float* const*>(ev_host), p, n, transform, cond_var, rng,
                                        static_cast<float*>(out));
    return PBN_OK;
}

}  // extern "C"
