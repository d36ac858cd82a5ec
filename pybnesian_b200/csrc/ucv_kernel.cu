// ucv_kernel.cu — pairwise sums of the unbiased cross-validation (UCV) bandwidth objective.
//
// Replaces UCVScorer::score_unconstrained_impl / score_diagonal_impl and their kernels
//   kde/UCV.cpp:235-358, kde/opencl_kernels/KDE.cl.src:471-574
// (4 launches per 10^6-pair chunk through an N_chunk x d scratch matrix, 32-bit chunk offsets)
// with one launch over an upper-triangular tile schedule with 64-bit unit ids:
//   S2 = sum_{j<i} exp(-s_ij/4),   S1 = sum_{j<i} exp(-s_ij/2),   s_ij = |L^-1 (x_i - x_j)|^2.
// Rows are whitened so that t = -sum_c (y_i,c - y_j,c)^2 is the exponent of exp(-s/4) in kernel
// units (f64: K*log2e*(-s/4), table exp2 as in pair_kernel.cuh; f32: log2e*(-s/4), MUFU ex2);
// exp(-s/2) is its square.  Per-CTA partial sums are written out and added in a fixed order.
#include "pair_kernel.cuh"

namespace pbn {

template <typename T, int D, bool DIAG, bool SAFE, int R>
__device__ __forceinline__ void ucv_tile(const T* __restrict__ tp, int cnt, long long col0, const T (&yi)[R][D],
                                         const long long (&rowid)[R], const double* __restrict__ tab,
                                         double (&s2)[R], double (&s1)[R]) {
    if constexpr (sizeof(T) == 8) {
        // exponent floor from this row's running sum of exp(-s/4) (pair_floor in pair_kernel.cuh): a floored term is below
        // 2^-80 of a partial sum of S2, and its square below 2^-160 N^2 of the largest term of S1
        int fl[R];
#pragma unroll
        for (int r = 0; r < R; ++r) fl[r] = pair_floor(s2[r]);
#pragma unroll 2
        for (int j = 0; j < cnt; ++j) {
            double p[D];
#pragma unroll
            for (int c = 0; c < D; ++c) p[c] = tp[j * D + c];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                double acc = 0.0;
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    double dl = yi[r][c] - p[c];
                    acc = fma(-dl, dl, acc);
                }
                double st;
                double pg = exp2_tab<SAFE>(acc, tab, st, 0, fl[r]);
                double e2 = st * pg;
                if (DIAG) e2 = (col0 + j < rowid[r]) ? e2 : 0.0;
                s2[r] += e2;
                s1[r] = fma(e2, e2, s1[r]);
            }
        }
    } else if constexpr (PBN_F32_PACKED && R % 2 == 0) {
        // packed FP32 (FFMA2 / FADD2), two rows per instruction: see tile_f32_packed in pair_kernel.cuh
        constexpr int H = R / 2;
        f32x2_t nyi[H][D], f2[H], f1[H];
#pragma unroll
        for (int h = 0; h < H; ++h) {
#pragma unroll
            for (int c = 0; c < D; ++c) nyi[h][c] = pack_f32x2(-yi[2 * h][c], -yi[2 * h + 1][c]);
            f2[h] = 0ull;
            f1[h] = 0ull;
        }
#pragma unroll 4
        for (int j = 0; j < cnt; ++j) {
            float p[D];
#pragma unroll
            for (int c = 0; c < D; ++c) p[c] = tp[j * D + c];
#pragma unroll
            for (int h = 0; h < H; ++h) {
                f32x2_t sq = 0ull;
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    f32x2_t d2 = fadd2(pack_f32x2(p[c], p[c]), nyi[h][c]);
                    sq = ffma2(d2, d2, sq);
                }
                float lo, hi;
                unpack_f32x2(sq, lo, hi);
                float e_lo = ex2_neg(lo), e_hi = ex2_neg(hi);
                if (DIAG) {
                    e_lo = (col0 + j < rowid[2 * h]) ? e_lo : 0.f;
                    e_hi = (col0 + j < rowid[2 * h + 1]) ? e_hi : 0.f;
                }
                f32x2_t e2 = pack_f32x2(e_lo, e_hi);
                f2[h] = fadd2(f2[h], e2);
                f1[h] = ffma2(e2, e2, f1[h]);
            }
        }
#pragma unroll
        for (int h = 0; h < H; ++h) {
            float lo, hi;
            unpack_f32x2(f2[h], lo, hi);
            s2[2 * h] += (double)lo;
            s2[2 * h + 1] += (double)hi;
            unpack_f32x2(f1[h], lo, hi);
            s1[2 * h] += (double)lo;
            s1[2 * h + 1] += (double)hi;
        }
    } else {
        float f2[R], f1[R];
#pragma unroll
        for (int r = 0; r < R; ++r) { f2[r] = 0.f; f1[r] = 0.f; }
#pragma unroll 4
        for (int j = 0; j < cnt; ++j) {
            float p[D];
#pragma unroll
            for (int c = 0; c < D; ++c) p[c] = tp[j * D + c];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float acc = 0.f;
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    float dl = yi[r][c] - p[c];
                    acc = fmaf(-dl, dl, acc);
                }
                float e2;
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e2) : "f"(acc));
                if (DIAG) e2 = (col0 + j < rowid[r]) ? e2 : 0.f;
                f2[r] += e2;
                f1[r] = fmaf(e2, e2, f1[r]);
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) { s2[r] += (double)f2[r]; s1[r] += (double)f1[r]; }
    }
}

// Dot-product form of the f64 tile (tile_f64_dot in pair_kernel.cuh): yi2 holds 2 y_i, the training norm nb is the addend
// of the first DFMA, the integer part of the row's own norm is added to the rounded exponent (ati) and its fraction is a
// per-row factor applied by the caller (scale on S2, scale^2 on S1).  D DFMA + 7 FP64 instructions per pair instead of 2 D + 7.
template <int D, bool DIAG, int R>
__device__ __forceinline__ void ucv_tile_dot(const double* __restrict__ tp, const double* __restrict__ nb, int cnt,
                                             long long col0, const double (&yi2)[R][D], const int (&ati)[R],
                                             const long long (&rowid)[R], const double* __restrict__ tab,
                                             double (&s2)[R], double (&s1)[R]) {
    int fl[R];
#pragma unroll
    for (int r = 0; r < R; ++r) fl[r] = pair_floor(s2[r]);
#pragma unroll 2
    for (int j = 0; j < cnt; ++j) {
        double p[D];
#pragma unroll
        for (int c = 0; c < D; ++c) p[c] = tp[j * D + c];
        const double b = nb[j];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            double acc = b;
#pragma unroll
            for (int c = 0; c < D; ++c) acc = fma(yi2[r][c], p[c], acc);
            double st;
            double pg = exp2_tab<false>(acc, tab, st, ati[r], fl[r]);
            double e2 = st * pg;
            if (DIAG) e2 = (col0 + j < rowid[r]) ? e2 : 0.0;
            s2[r] += e2;
            s1[r] = fma(e2, e2, s1[r]);
        }
    }
}

template <typename T, int D>
__global__ void __launch_bounds__(kThreads, 2)
ucv_kernel(UcvJob job, long long upb, const double* __restrict__ exp_tab_g) {
    constexpr int R = PairCfg<T>::R;
    constexpr int TILE = pair_tile<T>(D);
    constexpr int TB = kThreads * R;
    constexpr uint32_t TILE_BYTES = TILE * D * sizeof(T);
    constexpr uint32_t NRM_BYTES = pair_nrm_bytes<T>(D);  // per stage; 0 for f32
    constexpr bool DOT = PBN_F64_DOT && sizeof(T) == 8;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    T* tile_buf = reinterpret_cast<T*>(smem_raw);
    double* nrm_buf = reinterpret_cast<double*>(smem_raw + kStages * TILE_BYTES);  // [kStages][TILE]
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw + kStages * (TILE_BYTES + NRM_BYTES));
    double* tab = reinterpret_cast<double*>(smem_raw + kStages * (TILE_BYTES + NRM_BYTES) + 64);
    __shared__ double red[2][kThreads / 32];

    const int tid = threadIdx.x;
    const long long u0 = job.unit_begin + static_cast<long long>(blockIdx.x) * upb;
    long long u1 = u0 + upb;
    if (u1 > job.unit_end) u1 = job.unit_end;
    double s2[R], s1[R], scale[R];  // sums of the current row tile; scale: see ucv_tile_dot (1 in the difference form)
    double tot2 = 0, tot1 = 0;      // finished row tiles of this CTA
#pragma unroll
    for (int r = 0; r < R; ++r) { s2[r] = 0; s1[r] = 0; scale[r] = 1.0; }
    auto fold = [&]() {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            tot2 = fma(s2[r], scale[r], tot2);
            tot1 = fma(s1[r], scale[r] * scale[r], tot1);
            s2[r] = 0;
            s1[r] = 0;
        }
    };

    if (u0 < u1) {
        if (sizeof(T) == 8) exp_tab_fill(tab, exp_tab_g, tid, kThreads);
        tab += tid & (kExpRep - 1);  // this lane's copy
        if (tid == 0) {
            for (int s = 0; s < kStages; ++s) mbar_init(&full_bar[s], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        bool safe = true, dot = false;
        if (sizeof(T) == 8) {
            float a = job.bound ? *job.bound : INFINITY;
            safe = !(static_cast<float>(D) * 4.f * a * a < 2.0e9f);
            if (DOT) {  // same cancellation bound as pair_kernel (tile_f64_dot)
                float crit = sqrtf(static_cast<float>(D + 1)) * 2.f * D * a * a;
                dot = job.nrm && crit < static_cast<float>(kDotTol * 9007199254740992.0 / kExpA) && !safe;
            }
        }
        // row tile of the first unit
        int lo = 0, hi = job.n_row_tiles - 1;
        while (lo < hi) {
            int mid = (lo + hi + 1) >> 1;
            if (job.prefix[mid] <= u0) lo = mid; else hi = mid - 1;
        }
        int ptt = lo;       // producer's row tile
        long long pu = u0;  // producer's next unit
        auto issue_load = [&](int stage) {
            while (job.prefix[ptt + 1] <= pu) ++ptt;
            long long nt = pu - job.prefix[ptt];
            long long start = nt * TILE;
            long long cnt = job.n - start;
            if (cnt > TILE) cnt = TILE;
            uint32_t bytes = static_cast<uint32_t>(((cnt * D * sizeof(T)) + 15) & ~15ull);
            uint32_t nbytes = 0;
            if (DOT && dot) nbytes = static_cast<uint32_t>((cnt * sizeof(double) + 15) & ~15ull);
            mbar_expect_tx(&full_bar[stage], bytes + nbytes);
            tma_bulk_g2s(tile_buf + static_cast<size_t>(stage) * TILE * D,
                         reinterpret_cast<const T*>(job.y) + start * D, bytes, &full_bar[stage]);
            if (nbytes) tma_bulk_g2s(nrm_buf + static_cast<size_t>(stage) * TILE, job.nrm + start, nbytes, &full_bar[stage]);
            ++pu;
        };
        if (tid == 0)
            for (int s = 0; s < kStages && pu < u1; ++s) issue_load(s);

        int ctt = lo, cur_tt = -1;
        T yi[R][D];
        int ati[R];
        long long rowid[R];
        for (long long u = u0; u < u1; ++u) {
            const int stage = static_cast<int>((u - u0) % kStages);
            const uint32_t parity = static_cast<uint32_t>(((u - u0) / kStages) & 1);
            while (job.prefix[ctt + 1] <= u) ++ctt;
            const long long nt = u - job.prefix[ctt];
            if (ctt != cur_tt) {
                fold();
                cur_tt = ctt;
                const T* yp = reinterpret_cast<const T*>(job.y);
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    long long row = static_cast<long long>(ctt) * TB + r * kThreads + tid;
                    bool ok = row < job.n;
                    rowid[r] = ok ? row : -1;  // invalid rows pair with nothing (j < -1 never holds)
#pragma unroll
                    for (int c = 0; c < D; ++c) yi[r][c] = ok ? yp[row * D + c] : T(0);
                    ati[r] = 0;
                    if constexpr (DOT) {
                        if (dot) {
                            const double at = ok ? job.nrm[row] : 0.0;
                            const double ai = rint(at);  // |at| < 2^31 (the `safe` test above)
                            scale[r] = exp((at - ai) * kExpA);
                            ati[r] = static_cast<int>(ai);
#pragma unroll
                            for (int c = 0; c < D; ++c) yi[r][c] *= T(2);
                        }
                    }
                }
            }
            const long long col0 = nt * TILE;
            long long cnt_ll = job.n - col0;
            const int cnt = cnt_ll > TILE ? TILE : static_cast<int>(cnt_ll);
            // the tile needs the j < i mask unless all of its columns precede all of the CTA's valid rows
            const long long row_lo = static_cast<long long>(ctt) * TB;
            const bool full = (col0 + cnt <= row_lo) && (row_lo + TB <= job.n);
            mbar_wait(&full_bar[stage], parity);
            const T* __restrict__ tp = tile_buf + static_cast<size_t>(stage) * TILE * D;
            if constexpr (DOT) {
                if (dot) {
                    const double* nb = nrm_buf + static_cast<size_t>(stage) * TILE;
                    if (full) ucv_tile_dot<D, false, R>(tp, nb, cnt, col0, yi, ati, rowid, tab, s2, s1);
                    else ucv_tile_dot<D, true, R>(tp, nb, cnt, col0, yi, ati, rowid, tab, s2, s1);
                }
            }
            if (DOT && dot) {
            } else if (full) {
                if (safe) ucv_tile<T, D, false, true, R>(tp, cnt, col0, yi, rowid, tab, s2, s1);
                else ucv_tile<T, D, false, false, R>(tp, cnt, col0, yi, rowid, tab, s2, s1);
            } else {
                if (safe) ucv_tile<T, D, true, true, R>(tp, cnt, col0, yi, rowid, tab, s2, s1);
                else ucv_tile<T, D, true, false, R>(tp, cnt, col0, yi, rowid, tab, s2, s1);
            }
            __syncthreads();
            if (tid == 0 && pu < u1) issue_load(stage);
        }
    }
    // deterministic CTA reduction
    fold();
    double a = tot2, b = tot1;
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_down_sync(0xffffffffu, a, o);
        b += __shfl_down_sync(0xffffffffu, b, o);
    }
    if ((tid & 31) == 0) { red[0][tid >> 5] = a; red[1][tid >> 5] = b; }
    __syncthreads();
    if (tid == 0) {
        double x = 0, y = 0;
        for (int w = 0; w < kThreads / 32; ++w) { x += red[0][w]; y += red[1][w]; }
        job.partial[2 * blockIdx.x] = x;
        job.partial[2 * blockIdx.x + 1] = y;
    }
}

template <typename T, int D>
static cudaError_t launch_ucv_one(const UcvJob& job, long long upb, int grid, const double* tab, cudaStream_t stream) {
    constexpr size_t smem = kStages * (pair_tile<T>(D) * D * sizeof(T) + pair_nrm_bytes<T>(D)) + 64 + exp_tab_smem_bytes<T>();
    auto kern = ucv_kernel<T, D>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    kern<<<grid, kThreads, smem, stream>>>(job, upb, tab);
    return cudaGetLastError();
}

template <typename T>
static cudaError_t launch_ucv_t(int D, const UcvJob& job, long long upb, int grid, const double* tab, cudaStream_t s) {
    switch (D) {
        case 1: return launch_ucv_one<T, 1>(job, upb, grid, tab, s);
        case 2: return launch_ucv_one<T, 2>(job, upb, grid, tab, s);
        case 3: return launch_ucv_one<T, 3>(job, upb, grid, tab, s);
        case 4: return launch_ucv_one<T, 4>(job, upb, grid, tab, s);
        case 5: return launch_ucv_one<T, 5>(job, upb, grid, tab, s);
        case 6: return launch_ucv_one<T, 6>(job, upb, grid, tab, s);
        case 7: return launch_ucv_one<T, 7>(job, upb, grid, tab, s);
        case 8: return launch_ucv_one<T, 8>(job, upb, grid, tab, s);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_ucv(int dtype_f64, int D, const UcvJob& job, long long upb, int grid, const double* tab,
                       cudaStream_t s) {
    return dtype_f64 ? launch_ucv_t<double>(D, job, upb, grid, tab, s) : launch_ucv_t<float>(D, job, upb, grid, tab, s);
}

}  // namespace pbn
