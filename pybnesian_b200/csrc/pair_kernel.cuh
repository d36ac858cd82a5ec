// pair_kernel.cuh — the train x test Gaussian-kernel-sum kernel (sm_100a).
//
// Replaces, in one fused launch, what the reference does with 4 kernel launches per
// test row plus a 5-pass column log-sum-exp per 64 test rows:
//   kde/KDE.hpp:123-212, 592-640 (MultivariateKDE::execute_logl_mat, KDE::_logl_impl),
//   kde/opencl_kernels/KDE.cl.src:123-233 (substract, solve, square, logl_values_*),
//   opencl/opencl_config.hpp:517-536 (logsumexp_cols_offset).
// No N x m matrix is ever written: every (train, test) pair lives in registers.
//
// Inputs are *whitened* rows (see whiten_kernel in runtime.cu):  y = c * L^-1 (x - mu)
// with c chosen so that the exponent is already in table / log2 units:
//   f64:  t = -sum_c (yt_c - yi_c)^2  ==  K * log2(e) * (-1/2 s)   (s = Mahalanobis^2, K = 4096)
//         exp(-s/2) = 2^(t/K) = 2^k * T[j] * P(g),  n = rint(t), k = n>>12, j = n&4095,
//         g = t - n in [-1/2, 1/2], P a degree-2 polynomial in completed-square form c2 ((g + S)^2 + Cq) (max rel. err
//         2.2e-13 per term, zero mean; the parity bar is 1e-10).  Cost: 3 DADD + 1 DFMA + 1 DFMA (accumulate) on the FP64
//         pipe, 5 integer / shared-memory instructions for 2^k T[j] (exp2_tab).
//   f32:  t = -sum_c (..)^2 == log2(e) * (-1/2 s);  exp(-s/2) = ex2.approx(t) (1 MUFU), a per-shape share of them on the
//         FP32 pipe instead (ex2_neg_soft2).
// What bounds the f64 kernel is the scheduler's register read ports: an FP64 instruction costs 2 clk (3 with three register
// sources), anything else 1 clk, and nothing overlaps (profiles/r2_tuning.md section 6) - instruction COUNT is the lever.
// Sums are accumulated unshifted (every term <= 1); rows whose sum is too small for
// that to be accurate are re-run with a per-row shift by the caller (runtime.cu).
//
// Work decomposition ("stream-K"): a unit is (test tile of TB rows) x (train tile of
// TILE points); units of all jobs of a launch are numbered consecutively, test-tile
// major, and every CTA takes the same number of consecutive units.  A test tile that is
// covered by several CTAs gets one partial-sum slot per CTA; finalize_kernel adds the
// slots in a fixed order, so results are deterministic.
//
// Train tiles are staged in shared memory with 1-D TMA bulk copies
// (cp.async.bulk.shared::cluster.global + mbarrier complete_tx), double buffered.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "normcdf_coeffs.h"

namespace pbn {

struct PairJob {
    const void* train;     // whitened train rows, AoS [n_train (padded alloc)][D]
    const void* test;      // whitened test rows, AoS [m][D]
    double* part;          // partial sums: [n_acc][slots][m_pad]
    const float* bound_train;  // max |whitened coordinate| over the training rows (device scalar)
    const float* bound_test;   // same over the test rows
    const double* train_nrm;   // f64 only, may be null: -sum_{c<DN} y_c^2 per training row (DN = D-1 for CKDE, else D)
    const double* test_nrm;    // same per test row; both present -> the dot-product form may be used (see tile_f64_dot)
    long long n_train;
    long long m;
    long long m_pad;       // row stride of one slot
    long long unit_begin;  // first unit id of this job in the launch
    int n_train_tiles;
    int n_test_tiles;
    int slots;
    int pad_;
    // shifted second pass only (pair_kernel<..., SHIFT = true>, see runtime.cu: rows whose unshifted sums underflowed):
    // per test row, the smallest squared distance to a training row in kernel units (joint / marginal coordinates);
    // every term of that row is evaluated as 2^((t + shift)/K) (f64) or 2^(t + shift) (f32), so its largest term is ~1
    const float* shift_j;
    const float* shift_m;
    const int* test_rows;  // row r of the job is row test_rows[r] of `test`
    // tile skipping (single-job launches, runtime.cu): the units of the job are an explicit list instead of the full
    // (test tile) x (train tile) grid.  unit_list[q] = train tile of the job's q-th unit, test-tile major, ascending train
    // tile inside a test tile; tile_first[tt] .. tile_first[tt + 1] delimits the units of test tile tt.  Null = full grid.
    const int* unit_list;
    const long long* tile_first;
    // sums the FIRST CTA of every test tile starts from instead of zero: [n_acc][m_pad], the pass-A sums of tile skipping
    // (finalize then reads pass B alone).  With rows in Morton order the near training tiles no longer turn up early by
    // chance, so without them the running sums - and with them the exponent floor that keeps the table gathers free of
    // bank conflicts (pair_floor) - would stay tiny for most of a tile's sweep.  Null = start from zero.
    const double* init_sums;
};

// (test tile, train tile) of the job's `local`-th unit
__device__ __forceinline__ void pair_unit(const PairJob& jb, long long local, long long& tt, int& nt) {
    if (jb.unit_list) {
        // largest tt with tile_first[tt] <= local (test tiles without units repeat the same offset)
        long long lo = 0, hi = jb.n_test_tiles - 1;
        while (lo < hi) {
            const long long mid = (lo + hi + 1) >> 1;
            if (jb.tile_first[mid] <= local) lo = mid; else hi = mid - 1;
        }
        tt = lo;
        nt = jb.unit_list[local];
    } else {
        tt = local / jb.n_train_tiles;
        nt = static_cast<int>(local - tt * jb.n_train_tiles);
    }
}
// first unit (relative to unit_begin) of test tile tt
__device__ __forceinline__ long long pair_tile_first(const PairJob& jb, long long tt) {
    return jb.unit_list ? jb.tile_first[tt] : tt * jb.n_train_tiles;
}

constexpr int kThreads = 256;
// the pair kernel (log-likelihood and CDF modes) is instantiated for 1..kMaxFastD variables: north_star's family
// sizes (d = 1..10); wider families take the generic row kernels
constexpr int kMaxFastD = 10;
// tile skipping (spatial.cu): training tiles per test tile evaluated first, to lower-bound the sums of its rows
constexpr int kNearTiles = 8;
#ifndef PBN_F64_UNROLL
#define PBN_F64_UNROLL 4
#endif
constexpr int kF64Unroll = PBN_F64_UNROLL;  // training points per unrolled step of the f64 tile loops
constexpr int kStages = 2;

template <typename T> struct PairCfg;
#ifndef PBN_F64_R
#define PBN_F64_R 3
#endif
#ifndef PBN_F64_TILE
#define PBN_F64_TILE 512
#endif
#ifndef PBN_F64_MINCTAS
#define PBN_F64_MINCTAS 2
#endif
#ifndef PBN_F32_R
#define PBN_F32_R 4
#endif
#ifndef PBN_F32_TILE
#define PBN_F32_TILE 1024
#endif
// the persistent grid is 2 CTAs per SM (runtime.cu), so a tighter register bound only costs spills in the packed f32 tiles
#ifndef PBN_F32_MINCTAS
#define PBN_F32_MINCTAS 2
#endif
#ifndef PBN_F64_DOT
#define PBN_F64_DOT 1
#endif
// the dot-product form is used from this many norm coordinates on.  Round 1f measured no gain below DN = 4 (the extra
// norm loads cost what the saved DFMAs bought while the shared-memory pipe was saturated by table-gather conflicts);
// with the exponent floor (pair_floor) taking those conflicts away the FP64 count decides again: B200, N = m = 300k,
// CKDE d=4 (DN = 3) 1.437e12 -> 1.516e12 pair-evals/s, KDE d=2 1.365e12 -> 1.489e12, KDE d=1 1.615e12 -> 1.685e12
// (profiles/r1h_tuning.md).
#ifndef PBN_F64_DOT_MIN_DN
#define PBN_F64_DOT_MIN_DN 1
#endif
#ifndef PBN_EXP_BITS
#define PBN_EXP_BITS 12
#endif
#ifndef PBN_EXP_DEG
#define PBN_EXP_DEG 2
#endif
#ifndef PBN_EXP_REP
#define PBN_EXP_REP 1
#endif
#ifndef PBN_EXP_INT
#define PBN_EXP_INT 0
#endif
#ifndef PBN_F64_UNROLL
#define PBN_F64_UNROLL 4
#endif
// test rows per thread and training points per unrolled step, by kernel shape (B200, N = m = 300k, profiles/r1h_tuning.md):
// f64 KDE (no marginal sum) d <= 5 runs 4-5% faster with 4 rows per thread (d = 6: -2%), the CKDE kernels 1-3% slower
// (registers; CKDE d=2 +1.6%, not worth a special case);
// 8 points per step pay for d <= 2 only (KDE d=2 +5%, CKDE d=4 -3%, KDE d=8 -3%).
template <typename T> __host__ __device__ constexpr int pair_rows(int D, bool ckde);
template <> struct PairCfg<double> { static constexpr int R = PBN_F64_R; static constexpr int TILE = PBN_F64_TILE; static constexpr int MIN_CTAS = PBN_F64_MINCTAS; };
template <> struct PairCfg<float>  { static constexpr int R = PBN_F32_R; static constexpr int TILE = PBN_F32_TILE; static constexpr int MIN_CTAS = PBN_F32_MINCTAS; };

template <typename T> __host__ __device__ constexpr int pair_rows(int D, bool ckde) {
    // f64, d = 9, 10: two rows per thread (3 x 10 doubles of test rows + the training point do not fit 128 registers)
    return (sizeof(T) == 8 && D >= 9) ? 2 : (sizeof(T) == 8 && PBN_F64_R == 3 && !ckde && D <= 5) ? 4 : PairCfg<T>::R;
}
// rows per thread of the CDF mode (and of the UCV kernel, which stops at d = 8)
template <typename T> __host__ __device__ constexpr int pair_rows_cdf(int D) {
    return (sizeof(T) == 8 && D >= 9) ? 2 : PairCfg<T>::R;
}
// (round 2, with the completed-square exp2: 8 points per step pay up to d = 6 - CKDE d=4 +1.6%, KDE d=4 +1.3%, CKDE d=5 +1.2%,
// KDE d=6 +3%; KDE d=8 -0.6%)
__host__ __device__ constexpr int pair_unroll_f64(int D, bool ckde, bool cdf = false) {
    return (PBN_F64_UNROLL == 4 && (cdf ? (!ckde && D <= 2) : D <= 6)) ? 8 : PBN_F64_UNROLL;
}

// exp2 table of the f64 path: T[j] = 2^(j/K), K = 2^PBN_EXP_BITS, held in shared memory in kExpRep
// interleaved copies (copy r of entry j at index j*kExpRep + r; a thread reads copy lane mod kExpRep).
// kExpRep = 16 makes the gather bank-conflict free (2 wavefronts per 8-byte LDS instead of the ~6.6
// measured for one shared copy), but on B200 it bought nothing: the single-copy kernel reports the
// shared-memory pipe 98.5% busy (profiles/r1b_ncu_pair_f64_ckde_d4.txt) yet runs at the same speed as
// the conflict-free variant (profiles/r1c_tuning.md) - the replays hide behind the FP64 pipe, and the
// extra address arithmetic costs issue slots the DFMA stream needs (re-measured in round 1h with the current kernel:
// K = 512 x 16 copies -5%, profiles/r1h_tuning.md).  Default: one copy, K = 4096 (32 KB), degree 2.
constexpr int kExpTabBits = PBN_EXP_BITS;
constexpr int kExpTab = 1 << kExpTabBits;
constexpr int kExpRep = PBN_EXP_REP;
constexpr int kExpDeg = PBN_EXP_DEG;
static_assert(kExpRep == 1 || kExpRep == 2 || kExpRep == 4 || kExpRep == 8 || kExpRep == 16, "PBN_EXP_REP");
static_assert(kExpDeg >= 2 && kExpDeg <= 4, "PBN_EXP_DEG");
// train points per shared-memory tile: halved for wide f64 rows so that two CTAs (tiles + table) still
// fit the 227 KB of an SM
template <typename T> __host__ __device__ constexpr size_t exp_tab_smem_bytes() {
    return sizeof(T) == 8 ? static_cast<size_t>(kExpTab) * kExpRep * sizeof(double) : 0;
}
#ifndef PBN_F64_TILE_LE4
#define PBN_F64_TILE_LE4 PBN_F64_TILE
#endif
template <typename T> __host__ __device__ constexpr int pair_tile(int D) {
    return (sizeof(T) == 8 && D <= 4) ? PBN_F64_TILE_LE4
         : (sizeof(T) == 8 && (D >= 7 || (D >= 5 && exp_tab_smem_bytes<T>() > 32 * 1024))) ? PairCfg<T>::TILE / 2
                                                                                            : PairCfg<T>::TILE;
}
// per-stage bytes of the training-row norm tile (dot-product form, f64 only)
template <typename T> __host__ __device__ constexpr uint32_t pair_nrm_bytes(int D) {
    return sizeof(T) == 8 ? static_cast<uint32_t>(pair_tile<T>(D) * sizeof(double)) : 0u;
}
// P(g) ~ exp(a g), a = ln2/K, |g| <= 1/2: Taylor polynomial of degree kExpDeg + 1 with its leading term
// replaced by its Chebyshev economisation on [-h, h], h = a/2 (error = next Taylor term / 2^deg):
//   K =  512, degree 3: 1.1e-15      K = 2048, degree 3: 4.3e-18      K = 256, degree 4 (plain Taylor): 3.8e-17
//   K = 2048, degree 2: 2.0e-13      K = 4096, degree 2: 2.5e-14 (default)      K = 8192, degree 2: 3.1e-15
// Degree 2 saves one DFMA per exp2: +1.5% (CKDE d=4) .. +3% (KDE d=2) over K = 2048 / degree 3 on B200.
#ifndef PBN_EXP_SQ
#define PBN_EXP_SQ 1
#endif
// PBN_EXP_SQ (round 2, default): the degree-2 polynomial in COMPLETED-SQUARE form, which takes one FP64 instruction less
// and - what decides on sm_100 - fewer 64-bit REGISTER sources.  tools/micro/fp64_mix.cu (profiles/r2_fp64_mix.txt): an
// FP64 instruction costs its scheduler 2 clk with one or two distinct register sources and 3 clk with three, and every
// register a neighbouring integer instruction reads competes for the same two 32-bit read ports per clock; the pair
// kernel is bound by those ports, not by the FP64 pipe alone.
//   e^(a g) ~ 1 + a g + c2 g^2 = c2 ((g + S)^2 + Cq),   S = a / (2 c2),   Cq = 1 / c2 - S^2.
// S is made an INTEGER (round(K / ln 2) = 5909 for K = 4096) by giving c2 the relative error S ln2 / K - 1 = 4.7e-5 instead
// of touching the linear coefficient: the polynomial is then off by <= 4.7e-5 (a^2 / 2) g^2 + a^3 |g|^3 / 6 <= 2.7e-13 for
// |g| <= 1/2 - ten times the 2.5e-14 of the economised Horner form below, forty times inside the 1e-11 per-term budget
// (2.2e-13 and zero mean with the constant term corrected, kExpSqD).
// With S an integer, g + S comes straight out of the argument split (the second DADD subtracts MAGIC + S, still exact),
// and c2 is folded into the table (runtime.cu).  Per exp2: DADD, DADD, DADD, DFMA (g+S)^2 + Cq, DFMA accumulate = 5
// (was 6), with 1 + 1 + 2 + 1 + 3 = 8 register sources (was 11).
constexpr bool kExpSq = PBN_EXP_SQ != 0;
constexpr double kExpA = 0.693147180559945309417232121458 / kExpTab;
constexpr double kExpS = static_cast<double>(static_cast<long long>(1.0 / kExpA + 0.5));
constexpr double kExpSqC2 = kExpA / (2.0 * kExpS);           // the table holds c2 2^(j/K)
// constant term 1 + kExpSqD instead of 1: takes the MEAN of the quadratic-coefficient error (c2 - a^2/2) g^2 over |g| <= 1/2
// out, so that sums of many terms carry no bias (the error of a term stays below 2.2e-13)
constexpr double kExpSqD = -(kExpSqC2 - 0.5 * kExpA * kExpA) / 12.0;
constexpr double kExpSqC = (1.0 + kExpSqD) / kExpSqC2 - kExpS * kExpS;
// smallest power of two a scaled table entry may be moved to: c2 ~ 2^-26 for K = 4096 has to stay inside the normal range
constexpr int kExpMinK = kExpSq ? 1022 - 27 : 1022;  // a clamped term is evaluated as 2^-kExpMinK
constexpr double kExpH2 = 0.25 * kExpA * kExpA;  // h^2
constexpr double kExpC0 = kExpDeg == 3 ? 1.0 - kExpH2 * kExpH2 / 192.0 : 1.0;
constexpr double kExpC1 = kExpDeg == 2 ? kExpA * (1.0 + kExpH2 / 8.0) : kExpA;
constexpr double kExpC2 = kExpDeg == 3 ? kExpA * kExpA * (0.5 + kExpH2 / 24.0) : 0.5 * kExpA * kExpA;
constexpr double kExpC3 = kExpA * kExpA * kExpA / 6.0;
constexpr double kExpC4 = kExpA * kExpA * kExpA * kExpA / 24.0;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// Upper normal tail without branches (normcdf_coeffs.h), used by the CDF mode of the pair kernel:
//   Q(a) = 1/2 erfc(a / sqrt 2) = exp(-a^2/2) * t g(u)  for a >= 0,  t = 1/(1 + a/4),  u affine in t,
// g a degree-18 (f64: 5e-15 relative) / degree-9 (f32: 2e-7) polynomial valid for every a in [0, 39] (beyond that
// exp(-a^2/2) is 0 in double and the extrapolated polynomial stays finite).  Returns t g(u); the caller multiplies
// by exp(-a^2/2), which for CKDE::cdf is the JOINT kernel value the pair kernel computes anyway.
__device__ __forceinline__ double rcp_1toinf(double x) {  // x >= 1: MUFU.RCP64H (2^-23) + one cubic step -> 2^-69
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    return fma(r, fma(e, e, e), r);
}
__device__ __forceinline__ double normal_tail_tg(double a) {
    constexpr double c[kQDeg64 + 1] = PBN_Q64_COEFFS;
    double t = rcp_1toinf(fma(a, kQInvC, 1.0));
    double u = fma(t, kQScale, kQShift);
    double g = c[kQDeg64];
#pragma unroll
    for (int k = kQDeg64 - 1; k >= 0; --k) g = fma(g, u, c[k]);
    return t * g;
}
__device__ __forceinline__ float normal_tail_tg_f(float a) {
    constexpr float c[kQDeg32 + 1] = PBN_Q32_COEFFS;
    float t;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(a, static_cast<float>(kQInvC), 1.f)));
    float u = fmaf(t, static_cast<float>(kQScale), static_cast<float>(kQShift));
    float g = c[kQDeg32];
#pragma unroll
    for (int k = kQDeg32 - 1; k >= 0; --k) g = fmaf(g, u, c[k]);
    return t * g;
}

// exp(-s/2) from the f64 kernel exponent t (= K*log2e*(-s/2), t <= 0):
//   2^(t/K) = 2^k * T[j] * P(g),  n = rint(t) = K k + j,  g = t - n.
// The table holds T'[j] = T[j] with (j << (20 - log2 K)) subtracted from its high word, so that the
// scaled entry 2^k T[j] is obtained with ONE integer multiply-add: hi' + n * 2^(20 - log2 K).
// `tab` is the calling lane's copy of the table (tab_base + lane % kExpRep, stride kExpRep).
// Returns P(g); `scaled` receives 2^k T[j]  (PBN_EXP_SQ: (g + S)^2 + Cq and c2 2^k T[j] - the product is the same).  SAFE = false requires t > -2^31 (guaranteed by
// the caller from the bounding boxes of the whitened rows); SAFE = true accepts any t.
#ifndef PBN_DOT_TOL
#define PBN_DOT_TOL 1e-11
#endif
constexpr double kDotTol = PBN_DOT_TOL;  // worst-case per-term cancellation error accepted for the dot-product form
constexpr int kNMin = -kExpMinK * kExpTab;
// hi word of the double -(kExpMinK * K), 512 <= kExpMinK < 1024: sign | (1023 + 9 + log2 K) << 20 | top mantissa bits of
// kExpMinK / 512 - 1  (1022: 0xFF000)
static_assert(kExpMinK >= 512 && kExpMinK < 1024, "kExpMinK");
constexpr unsigned kHiLim = 0x80000000u | (static_cast<unsigned>(1023 + 9 + kExpTabBits) << 20) | (static_cast<unsigned>(kExpMinK - 512) << 11);
// `nshift` evaluates 2^((t + nshift)/K) instead: the shift is added to the rounded exponent (tile_f64_dot keeps the
// integer part of -|yt|^2 there); unsigned arithmetic, so only the shifted n has to fit 32 bits.
// `nmin` (a multiple of K, >= kNMin) is the floor of the rounded exponent: terms below 2^(nmin/K) are evaluated AS
// 2^(nmin/K) and all read table entry 0 (see pair_floor).
// hi word of the double -2^31: with WIDE the SAFE clamp only rejects |t| >= 2^31 (the rounded exponent no longer fits the
// integer side), because a positive `nshift` brings far smaller t back into range (shifted second pass)
constexpr unsigned kHiLim31 = 0x80000000u | (static_cast<unsigned>(1023 + 31) << 20);
template <bool SAFE, bool WIDE = false>
__device__ __forceinline__ double exp2_tab(double t, const double* __restrict__ tab, double& scaled,
                                           const int nshift = 0, const int nmin = kNMin) {
    const double MAGIC = 6755399441055744.0;  // 1.5 * 2^52
    double tm = t + MAGIC;
    int n = static_cast<int>(static_cast<unsigned>(__double2loint(tm)) + static_cast<unsigned>(nshift));
    double nd = tm - (kExpSq ? MAGIC + kExpS : MAGIC);  // MAGIC + S is an integer below 2^53: exact
    double g = t - nd;                                  // kExpSq: g + S
    double p;
    if (kExpSq) {
        p = fma(g, g, kExpSqC);
    } else if (kExpDeg == 4) {
        p = fma(kExpC4, g, kExpC3);
        p = fma(p, g, kExpC2);
        p = fma(p, g, kExpC1);
    } else if (kExpDeg == 3) {
        p = fma(kExpC3, g, kExpC2);
        p = fma(p, g, kExpC1);
    } else {
        p = fma(kExpC2, g, kExpC1);
    }
    if (!kExpSq) p = fma(p, g, kExpC0);
    if (SAFE) {
        // hi word of a negative double grows (as unsigned) with its magnitude;
        // kHiLim is the hi word of -(1022 * K)
        unsigned hi = static_cast<unsigned>(__double2hiint(t));
        n = (hi > (WIDE ? kHiLim31 : kHiLim)) ? kNMin : n;
        n = max(n, nmin);
    } else {
        n = max(n, nmin);
    }
#if PBN_EXP_INT == 1
    // integer glue on the ALU pipe (SHF / LOP3 / IADD3) instead of the FMA pipe (IMAD.SHL / IMAD): the FP64 stream
    // shares its dispatch port with the FMA pipe (profiles/r1c_tuning.md: DFMA + IMAD 1:1 costs 30% of the DFMA rate,
    // DFMA + LOP3 10%)
    unsigned rot, sh;
    asm("shf.l.wrap.b32 %0, %1, %1, 3;" : "=r"(rot) : "r"(n));          // (n << 3) | (n >> 29); low 3 bits masked off below
    asm("shf.l.clamp.b32 %0, %1, %2, %3;" : "=r"(sh) : "r"(0), "r"(n), "r"(20 - kExpTabBits));  // n << (20 - log2 K)
    double tj = *reinterpret_cast<const double*>(reinterpret_cast<const char*>(tab) + (rot & ((kExpTab - 1) << 3)));
    scaled = __hiloint2double(__double2hiint(tj) + static_cast<int>(sh), __double2loint(tj));
#else
    double tj = tab[(n & (kExpTab - 1)) * kExpRep];
    scaled = __hiloint2double(__double2hiint(tj) + n * (1 << (20 - kExpTabBits)), __double2loint(tj));
#endif
    return p;
}

// Floor of the rounded exponent for the terms still to be added to a running (unshifted, non-negative) sum:
//   floor = K * (exponent(sum) - kFloorBits), never below kNMin, kNMin while the sum is still zero / subnormal.
// A term below 2^(floor/K) is below 2^-kFloorBits of the sum it is added to; evaluating it as 2^(floor/K) changes the
// finished sum by less than N 2^-kFloorBits relative (N = 2^24 training rows: 1.4e-17), far below the rounding of the
// sum itself.  Why bother: the floor is a multiple of K, so every such lane reads table entry 0 - ONE shared-memory
// address per warp instead of a random one.  The table gather is what saturates the shared-memory pipe (ncu r1h: LSU
// wavefronts 99% of peak, 6.1 wavefronts per 8-byte gather against 2 without bank conflicts, FP64 pipe waiting at 73%),
// and for a KDE with a rule-of-thumb bandwidth most (test, train) pairs are that far apart, so the conflicts go with them.
// Costs nothing per pair (the clamp instruction was there already, with the constant kNMin).
#ifndef PBN_F64_FLOOR_BITS
#define PBN_F64_FLOOR_BITS 80
#endif
constexpr int kFloorBits = PBN_F64_FLOOR_BITS;  // 0 disables
__device__ __forceinline__ int pair_floor(double sum) {
    if (kFloorBits == 0) return kNMin;
    const int e = __double2hiint(sum) >> 20;  // sum >= 0 (or NaN, whose terms no longer matter)
    const int f = (e - 1023 - kFloorBits) * kExpTab;
    return (e <= 0 || e >= 2047) ? kNMin : max(f, kNMin);
}

// Fills the interleaved shared-memory copies of the table from the K-entry global table.
__device__ __forceinline__ void exp_tab_fill(double* __restrict__ tab_s, const double* __restrict__ tab_g, int tid,
                                             int nthreads) {
    for (int i = tid; i < kExpTab * kExpRep; i += nthreads) tab_s[i] = tab_g[i / kExpRep];
}

// One (test tile) x (train tile) unit on the FP64 pipe.
// CDF mode (CKDE::cdf, factors/continuous/CKDE.hpp:506-728): the "joint" accumulator receives w Phi(z) instead of the
// joint kernel value, z = (last whitened coordinate difference) * inv_c being (x_t - conditional mean_ti) / sqrt(cond_var):
//   w Phi(z) = z < 0 ? E tg(|z|) : w - E tg(|z|),   E = w exp(-z^2/2) = the joint kernel value,  w = marginal value
// (w = 1 for a CKDE without evidence, CKDE = false, D = 1).
template <int D, bool CKDE, bool SAFE, int R, bool CDF>
__device__ __forceinline__ void tile_f64(const double* __restrict__ tp, int cnt, const double (&yt)[R][D],
                                         const double* __restrict__ tab, double (&sum_j)[R], double (&sum_m)[R],
                                         double inv_c) {
    constexpr int U = pair_unroll_f64(D, CKDE, CDF);
    // exponent floors of this tile from the sums so far (log-likelihood sums only: a cdf sum may be far below its weights)
    int fl_j[R], fl_m[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        fl_j[r] = CDF ? kNMin : pair_floor(sum_j[r]);
        fl_m[r] = (CDF || !CKDE) ? kNMin : pair_floor(sum_m[r]);
    }
#pragma unroll U
    for (int i = 0; i < cnt; ++i) {
        double p[D];
#pragma unroll
        for (int c = 0; c < D; ++c) p[c] = tp[i * D + c];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            double acc = 0.0;
            double w = 1.0, dl_last = 0.0;
#pragma unroll
            for (int c = 0; c < D; ++c) {
                double dl = yt[r][c] - p[c];
                acc = fma(-dl, dl, acc);
                if (CDF && c == D - 1) dl_last = dl;
                if (CKDE && c == D - 2) {
                    double st;
                    double pm = exp2_tab<SAFE>(acc, tab, st, 0, fl_m[r]);
                    if (CDF) {
                        w = st * pm;
                        sum_m[r] += w;
                    } else {
                        sum_m[r] = fma(st, pm, sum_m[r]);
                    }
                }
            }
            double st;
            double pj = exp2_tab<SAFE>(acc, tab, st, 0, fl_j[r]);
            if (CDF) {
                double q = (st * pj) * normal_tail_tg(fabs(dl_last) * inv_c);
                sum_j[r] += (dl_last < 0.0) ? q : (w - q);
            } else {
                sum_j[r] = fma(st, pj, sum_j[r]);
            }
        }
    }
}

// Shifted form of tile_f64 for the second pass over rows whose unshifted sums underflowed (test rows tens of bandwidths
// away from every training row): the per-row integers sh_j / sh_m (~ the row's smallest squared distance, see
// rowmin_kernel in runtime.cu) are added to the rounded exponent, so the row's largest term is ~1 and the sums are as
// accurate as anywhere else.  This is what the reference's max-shifted logsumexp_cols_offset does for every row
// (opencl/opencl_config.hpp:517-536), at tile speed instead of one CTA per row.
template <int D, bool CKDE, int R>
__device__ __forceinline__ void tile_f64_shift(const double* __restrict__ tp, int cnt, const double (&yt)[R][D],
                                               const int (&sh_j)[R], const int (&sh_m)[R], const double* __restrict__ tab,
                                               double (&sum_j)[R], double (&sum_m)[R]) {
    int fl_j[R], fl_m[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        fl_j[r] = pair_floor(sum_j[r]);
        fl_m[r] = CKDE ? pair_floor(sum_m[r]) : kNMin;
    }
#pragma unroll 2
    for (int i = 0; i < cnt; ++i) {
        double p[D];
#pragma unroll
        for (int c = 0; c < D; ++c) p[c] = tp[i * D + c];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            double acc = 0.0;
#pragma unroll
            for (int c = 0; c < D; ++c) {
                double dl = yt[r][c] - p[c];
                acc = fma(-dl, dl, acc);
                if (CKDE && c == D - 2) {
                    double st;
                    double pm = exp2_tab<true, true>(acc, tab, st, sh_m[r], fl_m[r]);
                    sum_m[r] = fma(st, pm, sum_m[r]);
                }
            }
            double st;
            double pj = exp2_tab<true, true>(acc, tab, st, sh_j[r], fl_j[r]);
            sum_j[r] = fma(st, pj, sum_j[r]);
        }
    }
}

template <int D, bool CKDE, int R>
__device__ __forceinline__ void tile_f32_shift(const float* __restrict__ tp, int cnt, const float (&yt)[R][D],
                                               const float (&sh_j)[R], const float (&sh_m)[R], double (&sum_j)[R],
                                               double (&sum_m)[R]) {
    float facc_j[R], facc_m[R];
#pragma unroll
    for (int r = 0; r < R; ++r) { facc_j[r] = 0.f; facc_m[r] = 0.f; }
#pragma unroll 2
    for (int i = 0; i < cnt; ++i) {
        float p[D];
#pragma unroll
        for (int c = 0; c < D; ++c) p[c] = tp[i * D + c];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float acc = 0.f;
#pragma unroll
            for (int c = 0; c < D; ++c) {
                float dl = yt[r][c] - p[c];
                acc = fmaf(-dl, dl, acc);
                if (CKDE && c == D - 2) {
                    float e;
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(acc + sh_m[r]));
                    facc_m[r] += e;
                }
            }
            float e;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(acc + sh_j[r]));
            facc_j[r] += e;
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        sum_j[r] += static_cast<double>(facc_j[r]);
        if (CKDE) sum_m[r] += static_cast<double>(facc_m[r]);
    }
}

// Dot-product form of the same unit:  -|yt - yi|^2 = (-|yt|^2) + (-|yi|^2) + sum_c (2 yt_c) yi_c  costs DN + 1
// FP64 instructions instead of 2 DN (at = -|yt|^2 and nb = -|yi|^2 come from the whitening kernel, yt holds
// 2 yt_c for c < DN).  For a CKDE the conditioned variable (last coordinate) stays in difference form on top
// of the marginal exponent.  The cancellation error is ~ sqrt(DN+1) 2^-53 (2 DN B^2) table units (B = largest
// |whitened coordinate|); the caller only takes this path when that WORST-CASE bound (both points at the edge of the
// bounding box) is below kDotTol = 1e-11 relative per term - typical pairs are orders of magnitude better and the errors
// have random sign, so sums stay far inside the 1e-10 target (tests/test_fullsize_gpu.py checks 1e-10 on 1M x 1M).
//
// PBN_F64_HOIST: the per-pair DADD `at + nb` is removed as well.  at = A + f with A = rint(at): the integer A is added to
// the ROUNDED exponent on the integer side (one IADD on the ALU pipe: n = rint(acc) + A, the fraction g is unchanged),
// and the fraction f is a per-test-row FACTOR 2^(f/K) of every term of the row, applied once to the finished sums
// (pair_kernel's flush).  The exponent chain then starts from the training norm (the first DFMA takes nb as its
// addend): DN DFMA per pair instead of 1 DADD + DN DFMA.  Measured on B200 (N = m = 300k, profiles/r1h_tuning.md): KDE
// d=4 1.085e12 -> 1.141e12 pairs/s, d=8 8.29e11 -> 8.41e11.  Carrying A on the rounding constant instead (1.5 * 2^52 + A
// in a register, no IADD) was 3% SLOWER than not hoisting at all: the two DADDs of the exp2 then read a register pair
// where they had an immediate, and this kernel is sensitive to register-operand traffic, not to the FP64 count alone.
#ifndef PBN_F64_HOIST
#define PBN_F64_HOIST 1
#endif
// WSKIP (tile skipping on, rows in Morton order): a warp whose 32 rows ALL have this training point's term at or below
// their exponent floor skips the exp2 of that term altogether - the term would have been evaluated AS the floor, i.e. as
// less than 2^-kFloorBits of the sum it joins (pair_floor), so dropping it changes the finished sum by less than the
// floor mechanism itself already does.  The exponent (dot product, rounding) is still computed for every pair; what is
// saved is 4 of the 5 FP64 instructions of the exp2 and its table gather, for the large majority of the pairs of a
// localised kernel sum.  Only worthwhile when neighbouring lanes hold neighbouring rows, hence tied to the Morton order.
// MEASURED (B200, 1M x 1M, round 2, profiles/r2_tuning.md): slower than evaluating the floor terms - KDE d=2 5.03e12 ->
// 3.58e12, d=4 1.67e12 -> 1.39e12 pair-evals/s with skipping on: the 24 warp-uniform branches per unrolled step cut the
// 12 independent exp2 chains the scheduler interleaves into short dependent runs, and four warps per scheduler cannot
// hide them.  Kept as a build option, off.
#ifndef PBN_F64_WARPSKIP
#define PBN_F64_WARPSKIP 0
#endif
template <int D, bool CKDE, int R, bool CDF, bool WSKIP = false>
__device__ __forceinline__ void tile_f64_dot(const double* __restrict__ tp, const double* __restrict__ nb, int cnt,
                                             const double (&yt)[R][D], const double (&at)[R], const int (&ati)[R],
                                             const double* __restrict__ tab, double (&sum_j)[R], double (&sum_m)[R],
                                             double inv_c) {
    constexpr int DN = CKDE ? D - 1 : D;
    constexpr int U = pair_unroll_f64(D, CKDE, CDF);
    int fl_j[R], fl_m[R];  // see tile_f64
#pragma unroll
    for (int r = 0; r < R; ++r) {
        fl_j[r] = CDF ? kNMin : pair_floor(sum_j[r]);
        fl_m[r] = (CDF || !CKDE) ? kNMin : pair_floor(sum_m[r]);
    }
#pragma unroll U
    for (int i = 0; i < cnt; ++i) {
        double p[D];
#pragma unroll
        for (int c = 0; c < D; ++c) p[c] = tp[i * D + c];
        const double b = nb[i];
#pragma unroll
        for (int r = 0; r < R; ++r) {
#if PBN_F64_HOIST
            const int ns = ati[r];  // rint(-|yt|^2)
            double acc = b;
#else
            const int ns = 0;
            double acc = at[r] + b;
#endif
            double w = 1.0, dl_last = 0.0;
#pragma unroll
            for (int c = 0; c < DN; ++c) acc = fma(yt[r][c], p[c], acc);
            if (CKDE) {
                bool live = true;
                if (WSKIP) {
                    const int n0 = static_cast<int>(static_cast<unsigned>(__double2loint(acc + 6755399441055744.0)) + static_cast<unsigned>(ns));
                    live = !__all_sync(0xffffffffu, n0 <= fl_m[r]);
                }
                if (live) {
                    double st;
                    double pm = exp2_tab<false>(acc, tab, st, ns, fl_m[r]);
                    if (CDF) {
                        w = st * pm;
                        sum_m[r] += w;
                    } else {
                        sum_m[r] = fma(st, pm, sum_m[r]);
                    }
                }
                double dl = yt[r][D - 1] - p[D - 1];
                dl_last = dl;
                acc = fma(-dl, dl, acc);
            }
            bool livej = true;
            if (WSKIP) {
                const int n0 = static_cast<int>(static_cast<unsigned>(__double2loint(acc + 6755399441055744.0)) + static_cast<unsigned>(ns));
                livej = !__all_sync(0xffffffffu, n0 <= fl_j[r]);
            }
            if (livej) {
                double st;
                double pj = exp2_tab<false>(acc, tab, st, ns, fl_j[r]);
                if (CDF && CKDE) {
                    double q = (st * pj) * normal_tail_tg(fabs(dl_last) * inv_c);
                    sum_j[r] += (dl_last < 0.0) ? q : (w - q);
                } else {
                    sum_j[r] = fma(st, pj, sum_j[r]);
                }
            }
        }
    }
}

// GROUP SKIPPING (pass B of tile skipping, rows in Morton order; `pair_kernel<..., GSKIP = true>`): the bounding boxes of
// spatial.cu decide per (768 test rows x 512 training points); inside a unit that stays, a WARP (3 x 32 neighbouring test
// rows) still meets many groups of G = 2..4 neighbouring training points whose terms are all negligible.  Per group the
// marginal exponents (the dot products, needed anyway) are compared with a per-row threshold before anything else is
// computed; if no lane of the warp has a term above it, the group's exp2 (10 of the 15 FP64 instructions of a CKDE row
// pair, all 10 integer ones) are not executed.  One warp-uniform branch per G x R row pairs, the two sides are long
// straight-line blocks - unlike the per-term vote (PBN_F64_WARPSKIP), which was slower than no skipping at all.
//   threshold = K (exponent(running sum) - kSkipBits) - A_row: a skipped term is below 2^-kSkipBits of the sum it would
//   join, all skipped terms of a row below N 2^-kSkipBits of it (2^-44 for a million rows: 6e-14 on logl).  A CKDE's joint
//   exponent never exceeds its marginal one, so the test on the marginal exponent against the LOWER of the two thresholds
//   covers both sums.  Pass B starts from the pass-A sums, so the thresholds are in place from the first group on.
#ifndef PBN_F64_GSKIP_BITS
#define PBN_F64_GSKIP_BITS 64
#endif
#ifndef PBN_F64_GSKIP_GROUP
#define PBN_F64_GSKIP_GROUP 0
#endif
constexpr int kSkipBits = PBN_F64_GSKIP_BITS;
// training points per group: 2.  The group's R x G exponents stay live across the branch: 4 points spill next to R = 3 or 4
// test rows, and finer groups are skipped more often (B200, 1M x 1M, profiles/r2_tuning.md section 7: CKDE d=4 2.65e12
// plain, 3.06e12 with G = 4, 3.13e12 with G = 3, 3.25e12 with G = 2)
__host__ __device__ constexpr int pair_gskip_group(bool ckde) {
    return PBN_F64_GSKIP_GROUP > 0 ? PBN_F64_GSKIP_GROUP : 2;
}
__device__ __forceinline__ int pair_skip_level(double sum) {
    const int e = __double2hiint(sum) >> 20;  // sum >= 0
    const int f = (e - 1023 - kSkipBits) * kExpTab;
    return (e <= 0 || e >= 2047) ? kNMin : max(f, kNMin);
}
template <int D, bool CKDE, int R>
__device__ __forceinline__ void tile_f64_dot_gskip(const double* __restrict__ tp, const double* __restrict__ nb, int cnt,
                                                   const double (&yt)[R][D], const int (&ati)[R],
                                                   const double* __restrict__ tab, double (&sum_j)[R], double (&sum_m)[R]) {
    static_assert(PBN_F64_HOIST, "the group-skipping tile takes the test-row norm on the integer side (ati) only");
    constexpr int DN = CKDE ? D - 1 : D;
    constexpr int G = pair_gskip_group(CKDE);
    int fl_j[R], fl_m[R], thr[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        fl_j[r] = pair_floor(sum_j[r]);
        fl_m[r] = CKDE ? pair_floor(sum_m[r]) : kNMin;
        int lv = pair_skip_level(sum_j[r]);
        if (CKDE) lv = min(lv, pair_skip_level(sum_m[r]));
        thr[r] = lv - ati[r];  // |ati| < 2^31 - |kNMin| (the `safe` test of the caller): no overflow
    }
    auto finish = [&](int r, double acc, double p_last) {  // acc: marginal exponent of (row r, one training point)
        const int ns = ati[r];
        if (CKDE) {
            double st;
            double pm = exp2_tab<false>(acc, tab, st, ns, fl_m[r]);
            sum_m[r] = fma(st, pm, sum_m[r]);
            double dl = yt[r][D - 1] - p_last;
            acc = fma(-dl, dl, acc);
        }
        double st;
        double pj = exp2_tab<false>(acc, tab, st, ns, fl_j[r]);
        sum_j[r] = fma(st, pj, sum_j[r]);
    };
    int i = 0;
    for (; i + G <= cnt; i += G) {
        double acc[G][R];
        int top[R];
#pragma unroll
        for (int g = 0; g < G; ++g) {
            double p[DN];
#pragma unroll
            for (int c = 0; c < DN; ++c) p[c] = tp[(i + g) * D + c];
            const double b = nb[i + g];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                double a = b;
#pragma unroll
                for (int c = 0; c < DN; ++c) a = fma(yt[r][c], p[c], a);
                acc[g][r] = a;
                const int lo = __double2loint(a + 6755399441055744.0);  // rint(a): the same DADD opens exp2_tab
                top[r] = g == 0 ? lo : max(top[r], lo);
            }
        }
        bool live = false;
#pragma unroll
        for (int r = 0; r < R; ++r) live |= top[r] > thr[r];
        if (__any_sync(0xffffffffu, live)) {
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const double p_last = CKDE ? tp[(i + g) * D + D - 1] : 0.0;
#pragma unroll
                for (int r = 0; r < R; ++r) finish(r, acc[g][r], p_last);
            }
        }
    }
    for (; i < cnt; ++i) {  // tail of the last training tile
        double p[D];
#pragma unroll
        for (int c = 0; c < D; ++c) p[c] = tp[i * D + c];
        const double b = nb[i];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            double a = b;
#pragma unroll
            for (int c = 0; c < DN; ++c) a = fma(yt[r][c], p[c], a);
            finish(r, a, p[D - 1]);
        }
    }
}

// Same unit on the FP32 FMA pipe + MUFU.EX2; per-tile float sums are folded into the
// double accumulators by the caller (summation error stays at ~sqrt(TILE) ulp).
template <int D, bool CKDE, int R, bool CDF>
__device__ __forceinline__ void tile_f32(const float* __restrict__ tp, int cnt, const float (&yt)[R][D],
                                         double (&sum_j)[R], double (&sum_m)[R], float inv_c) {
    float facc_j[R], facc_m[R];
#pragma unroll
    for (int r = 0; r < R; ++r) { facc_j[r] = 0.f; facc_m[r] = 0.f; }
#pragma unroll 4
    for (int i = 0; i < cnt; ++i) {
        float p[D];
#pragma unroll
        for (int c = 0; c < D; ++c) p[c] = tp[i * D + c];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float acc = 0.f;
            float w = 1.f, dl_last = 0.f;
#pragma unroll
            for (int c = 0; c < D; ++c) {
                float dl = yt[r][c] - p[c];
                acc = fmaf(-dl, dl, acc);
                if (CDF && c == D - 1) dl_last = dl;
                if (CKDE && c == D - 2) {
                    float e;
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(acc));
                    facc_m[r] += e;
                    w = e;
                }
            }
            float e;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(acc));
            if (CDF) {
                float q = e * normal_tail_tg_f(fabsf(dl_last) * inv_c);
                facc_j[r] += (dl_last < 0.f) ? q : (w - q);
            } else {
                facc_j[r] += e;
            }
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        sum_j[r] += static_cast<double>(facc_j[r]);
        if (CKDE) sum_m[r] += static_cast<double>(facc_m[r]);
    }
}

// Packed form of tile_f32 (sm_100 FFMA2 / FADD2: fma.rn.f32x2 / add.rn.f32x2 work on two floats per lane).  Two test rows
// of a thread share one instruction, which halves the issue slots of the subtract / square-accumulate / sum stream; the
// FP32 pipe rate itself is unchanged (tools/micro/fp32_peak.cu: 120 lane-ops/clk/SM scalar or packed), so this helps
// where the scalar kernel is issue-bound (d >= 3).  The squared distance is accumulated positive and negated at the
// MUFU.EX2 input; p - yt is used instead of yt - p (the sign is squared away), so only -yt is needed, once per tile.
#ifndef PBN_F32_PACKED
#define PBN_F32_PACKED 1
#endif
#ifndef PBN_F32_PACKED_MIN_D
#define PBN_F32_PACKED_MIN_D 1
#endif
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t pack_f32x2(float lo, float hi) {
    f32x2_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack_f32x2(f32x2_t v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2_t fadd2(f32x2_t a, f32x2_t b) {
    f32x2_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2_t ffma2(f32x2_t a, f32x2_t b, f32x2_t c) {
    f32x2_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ float ex2_neg(float s) {  // 2^(-s)
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-s));
    return e;
}

// MUFU OFFLOAD.  The f32 kernels of 1-4 variables are bound by MUFU.EX2 (16 per clock and SM: 8 clk of the scheduler's SFU
// per warp instruction) while the FP32 pipe idles, so a share of the exponentials is evaluated on the FP32 pipe instead,
// two rows per instruction: 2^(-s) = 2^n P(f), n = rint(-s) from the mantissa of -s + 1.5 2^23, f = -s - n in [-1/2, 1/2],
// P the degree-4 interpolant of 2^f at the Chebyshev nodes (3.5e-6 relative, zero mean; MUFU.EX2: 2e-7; the float32 bar
// is 1e-4), the exponent added to the bit pattern.  3 FADD2/FFMA2 to split + 4 FFMA2 + 2 FMNMX + 4 integer instructions per
// TWO exponentials.  Which of the (point, joint / marginal) slots of a 4-point step take this path is a compile-time mask
// (PBN_F32_SOFT_MASK, bit 2 u + 1 = joint / only exponential of point u, bit 2 u = marginal of point u), so there is no
// divergence and the MUFU and FP32 streams interleave.
#ifndef PBN_F32_SOFT_MASK
#define PBN_F32_SOFT_MASK -1
#endif
// B200, N = m = 400k, pair-evals/s (profiles/r2_tuning.md section 8): KDE d=1 4.42e12 -> 5.02e12 with one exponential in four
// offloaded (the MUFU roof is 4.65e12); CKDE d=2 4.48e12 -> 5.43e12, d=3 4.40e12 -> 5.01e12 with the joint exponential of every
// second point (one in four); CKDE d=4 4.28e12 -> 4.62e12 with one in eight.  KDE d >= 2 and wider CKDEs are bound by the FP32
// pipe / issue slots already and lose (KDE d=3: -13%): mask 0.
__host__ __device__ constexpr unsigned pair_f32_soft_mask(int D, bool ckde) {
    return PBN_F32_SOFT_MASK >= 0 ? static_cast<unsigned>(PBN_F32_SOFT_MASK)
         : ckde ? (D <= 3 ? 0x88u : D == 4 ? 0x80u : 0u)
                : (D == 1 ? 0x80u : 0u);
}
__device__ __forceinline__ f32x2_t ex2_neg_soft2(f32x2_t s2) {
    float lo, hi;
    unpack_f32x2(s2, lo, hi);
    s2 = pack_f32x2(fminf(lo, 125.f), fminf(hi, 125.f));  // 2^-125 ~ 2e-38: nothing next to a sum above the 2^-64 flag level
    const f32x2_t MAGIC = pack_f32x2(12582912.f, 12582912.f), NMAGIC = pack_f32x2(-12582912.f, -12582912.f);
    const f32x2_t r = ffma2(s2, pack_f32x2(-1.f, -1.f), MAGIC);   // low mantissa bits: n = rint(-s)
    const f32x2_t g = fadd2(s2, fadd2(r, NMAGIC));                // s + n = -f
    // P(f) = 1 + c1 f + c2 f^2 + c3 f^3 + c4 f^4 in g = -f: odd coefficients change sign
    f32x2_t p = ffma2(g, pack_f32x2(0.009666368515f, 0.009666368515f), pack_f32x2(-0.05592197584f, -0.05592197584f));
    p = ffma2(p, g, pack_f32x2(0.2402234904f, 0.2402234904f));
    p = ffma2(p, g, pack_f32x2(-0.6931210452f, -0.6931210452f));
    p = ffma2(p, g, pack_f32x2(1.f, 1.f));
    float plo, phi, rlo, rhi;
    unpack_f32x2(p, plo, phi);
    unpack_f32x2(r, rlo, rhi);
    plo = __int_as_float(__float_as_int(plo) + (__float_as_int(rlo) << 23));
    phi = __int_as_float(__float_as_int(phi) + (__float_as_int(rhi) << 23));
    return pack_f32x2(plo, phi);
}

template <int D, bool CKDE, int R>
__device__ __forceinline__ void tile_f32_packed(const float* __restrict__ tp, int cnt, const float (&yt)[R][D],
                                                double (&sum_j)[R], double (&sum_m)[R]) {
    static_assert(R % 2 == 0, "packed f32 tile needs an even number of rows per thread");
    constexpr int H = R / 2;
    constexpr unsigned SOFT = pair_f32_soft_mask(D, CKDE);
    f32x2_t nyt[H][D];
    f32x2_t facc_j[H], facc_m[H];
#pragma unroll
    for (int h = 0; h < H; ++h) {
#pragma unroll
        for (int c = 0; c < D; ++c) nyt[h][c] = pack_f32x2(-yt[2 * h][c], -yt[2 * h + 1][c]);
        facc_j[h] = 0ull;
        facc_m[h] = 0ull;
    }
    auto point = [&](int i, bool soft_j, bool soft_m) {
        f32x2_t p2[D];
#pragma unroll
        for (int c = 0; c < D; ++c) {
            float v = tp[i * D + c];
            p2[c] = pack_f32x2(v, v);
        }
#pragma unroll
        for (int h = 0; h < H; ++h) {
            f32x2_t s2 = 0ull;
#pragma unroll
            for (int c = 0; c < D; ++c) {
                f32x2_t d2 = fadd2(p2[c], nyt[h][c]);
                s2 = ffma2(d2, d2, s2);
                if (CKDE && c == D - 2) {
                    if (soft_m) {
                        facc_m[h] = fadd2(facc_m[h], ex2_neg_soft2(s2));
                    } else {
                        float lo, hi;
                        unpack_f32x2(s2, lo, hi);
                        facc_m[h] = fadd2(facc_m[h], pack_f32x2(ex2_neg(lo), ex2_neg(hi)));
                    }
                }
            }
            if (soft_j) {
                facc_j[h] = fadd2(facc_j[h], ex2_neg_soft2(s2));
            } else {
                float lo, hi;
                unpack_f32x2(s2, lo, hi);
                facc_j[h] = fadd2(facc_j[h], pack_f32x2(ex2_neg(lo), ex2_neg(hi)));
            }
        }
    };
    int i = 0;
    for (; i + 4 <= cnt; i += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) point(i + u, (SOFT >> (2 * u + 1)) & 1u, (SOFT >> (2 * u)) & 1u);
    }
    for (; i < cnt; ++i) point(i, false, false);
#pragma unroll
    for (int h = 0; h < H; ++h) {
        float lo, hi;
        unpack_f32x2(facc_j[h], lo, hi);
        sum_j[2 * h] += static_cast<double>(lo);
        sum_j[2 * h + 1] += static_cast<double>(hi);
        if (CKDE) {
            unpack_f32x2(facc_m[h], lo, hi);
            sum_m[2 * h] += static_cast<double>(lo);
            sum_m[2 * h + 1] += static_cast<double>(hi);
        }
    }
}

// Group skipping for the packed f32 tile (pass B of tile skipping, `pair_kernel<float, ..., GSKIP>`; see tile_f64_dot_gskip).
// The narrow f32 shapes are bound by MUFU.EX2, so what a skipped group saves is exactly the scarce resource: per group of 2
// training points the squared distances over the marginal coordinates are formed as always (FP32 pipe), each row's smallest
// one is compared with thr = kSkipBits32 - floor(log2(running sum)) and, when no lane of the warp has a term above
// 2^-kSkipBits32 of its sum, the group's 2 (KDE) or 4 (CKDE) exponentials per row, the last-coordinate step and the
// accumulates are not executed.  All skipped terms of a row are below N 2^-44 of its sum (6e-8 at a million rows; bar 1e-4).
#ifndef PBN_F32_GSKIP_BITS
#define PBN_F32_GSKIP_BITS 44
#endif
constexpr int kSkipBits32 = PBN_F32_GSKIP_BITS;
__device__ __forceinline__ float pair_skip_level_f32(double sum) {
    const int e = __double2hiint(sum) >> 20;  // sum >= 0: biased exponent
    return (e <= 0 || e >= 2047) ? INFINITY : static_cast<float>(kSkipBits32 + 1023 - e);
}
template <int D, bool CKDE, int R>
__device__ __forceinline__ void tile_f32_packed_gskip(const float* __restrict__ tp, int cnt, const float (&yt)[R][D],
                                                      double (&sum_j)[R], double (&sum_m)[R]) {
    static_assert(R % 2 == 0, "packed f32 tile needs an even number of rows per thread");
    constexpr int H = R / 2;
    constexpr int DN = CKDE ? D - 1 : D;
    constexpr int G = 2;
    f32x2_t nyt[H][D];
    f32x2_t facc_j[H], facc_m[H];
    float thr[R];
#pragma unroll
    for (int h = 0; h < H; ++h) {
#pragma unroll
        for (int c = 0; c < D; ++c) nyt[h][c] = pack_f32x2(-yt[2 * h][c], -yt[2 * h + 1][c]);
        facc_j[h] = 0ull;
        facc_m[h] = 0ull;
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        // a CKDE's joint term never exceeds its marginal one: the test on the marginal distance against the HIGHER of the
        // two levels (the smaller sum's) covers both
        thr[r] = pair_skip_level_f32(sum_j[r]);
        if (CKDE) thr[r] = fmaxf(thr[r], pair_skip_level_f32(sum_m[r]));
    }
    auto finish = [&](int h, f32x2_t s2, f32x2_t p_last2) {
        if (CKDE) {
            float lo, hi;
            unpack_f32x2(s2, lo, hi);
            facc_m[h] = fadd2(facc_m[h], pack_f32x2(ex2_neg(lo), ex2_neg(hi)));
            f32x2_t d2 = fadd2(p_last2, nyt[h][D - 1]);
            s2 = ffma2(d2, d2, s2);
        }
        float lo, hi;
        unpack_f32x2(s2, lo, hi);
        facc_j[h] = fadd2(facc_j[h], pack_f32x2(ex2_neg(lo), ex2_neg(hi)));
    };
    int i = 0;
    for (; i + G <= cnt; i += G) {
        f32x2_t s2[G][H];
        float low[R];
#pragma unroll
        for (int g = 0; g < G; ++g) {
            f32x2_t p2[DN];
#pragma unroll
            for (int c = 0; c < DN; ++c) {
                float v = tp[(i + g) * D + c];
                p2[c] = pack_f32x2(v, v);
            }
#pragma unroll
            for (int h = 0; h < H; ++h) {
                f32x2_t a = 0ull;
#pragma unroll
                for (int c = 0; c < DN; ++c) {
                    f32x2_t d2 = fadd2(p2[c], nyt[h][c]);
                    a = ffma2(d2, d2, a);
                }
                s2[g][h] = a;
                float lo, hi;
                unpack_f32x2(a, lo, hi);
                low[2 * h] = g == 0 ? lo : fminf(low[2 * h], lo);
                low[2 * h + 1] = g == 0 ? hi : fminf(low[2 * h + 1], hi);
            }
        }
        bool live = false;
#pragma unroll
        for (int r = 0; r < R; ++r) live |= low[r] < thr[r];
        if (__any_sync(0xffffffffu, live)) {
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const float v = CKDE ? tp[(i + g) * D + D - 1] : 0.f;
                const f32x2_t p_last2 = pack_f32x2(v, v);
#pragma unroll
                for (int h = 0; h < H; ++h) finish(h, s2[g][h], p_last2);
            }
        }
    }
    for (; i < cnt; ++i) {  // tail of the last training tile
        f32x2_t p2[D];
#pragma unroll
        for (int c = 0; c < D; ++c) {
            float v = tp[i * D + c];
            p2[c] = pack_f32x2(v, v);
        }
#pragma unroll
        for (int h = 0; h < H; ++h) {
            f32x2_t a = 0ull;
#pragma unroll
            for (int c = 0; c < DN; ++c) {
                f32x2_t d2 = fadd2(p2[c], nyt[h][c]);
                a = ffma2(d2, d2, a);
            }
            finish(h, a, p2[D - 1]);
        }
    }
#pragma unroll
    for (int h = 0; h < H; ++h) {
        float lo, hi;
        unpack_f32x2(facc_j[h], lo, hi);
        sum_j[2 * h] += static_cast<double>(lo);
        sum_j[2 * h + 1] += static_cast<double>(hi);
        if (CKDE) {
            unpack_f32x2(facc_m[h], lo, hi);
            sum_m[2 * h] += static_cast<double>(lo);
            sum_m[2 * h + 1] += static_cast<double>(hi);
        }
    }
}

__device__ __forceinline__ f32x2_t fsub2(f32x2_t a, f32x2_t b) {
    f32x2_t r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2_t fmul2(f32x2_t a, f32x2_t b) {
    f32x2_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

// CDF mode of the packed f32 tile (see tile_f64 for the formula): everything per row is elementwise, so the exponent,
// the tail polynomial and the products pack two rows per instruction; only the three MUFU calls (2 ex2, 1 rcp) and the
// sign select are per lane.  Here yt - p is needed with its sign, so the training coordinate is subtracted (FADD2 with a
// negated broadcast operand).
template <int D, bool CKDE, int R>
__device__ __forceinline__ void tile_f32_packed_cdf(const float* __restrict__ tp, int cnt, const float (&yt)[R][D],
                                                    double (&sum_j)[R], double (&sum_m)[R], float inv_c) {
    static_assert(R % 2 == 0, "packed f32 tile needs an even number of rows per thread");
    constexpr int H = R / 2;
    constexpr float c[kQDeg32 + 1] = PBN_Q32_COEFFS;
    f32x2_t y2[H][D], facc_j[H], facc_m[H];
#pragma unroll
    for (int h = 0; h < H; ++h) {
#pragma unroll
        for (int k = 0; k < D; ++k) y2[h][k] = pack_f32x2(yt[2 * h][k], yt[2 * h + 1][k]);
        facc_j[h] = 0ull;
        facc_m[h] = 0ull;
    }
    const f32x2_t one2 = pack_f32x2(1.f, 1.f);
    const f32x2_t invc2 = pack_f32x2(inv_c, inv_c);
    const f32x2_t qinv2 = pack_f32x2(static_cast<float>(kQInvC), static_cast<float>(kQInvC));
    const f32x2_t qs2 = pack_f32x2(static_cast<float>(kQScale), static_cast<float>(kQScale));
    const f32x2_t qh2 = pack_f32x2(static_cast<float>(kQShift), static_cast<float>(kQShift));
#pragma unroll 2
    for (int i = 0; i < cnt; ++i) {
        f32x2_t p2[D];
#pragma unroll
        for (int k = 0; k < D; ++k) {
            float v = tp[i * D + k];
            p2[k] = pack_f32x2(v, v);
        }
#pragma unroll
        for (int h = 0; h < H; ++h) {
            f32x2_t s2 = 0ull, w2 = one2, dl2 = 0ull;
#pragma unroll
            for (int k = 0; k < D; ++k) {
                f32x2_t d2 = fsub2(y2[h][k], p2[k]);
                if (k == D - 1) dl2 = d2;
                if (CKDE && k == D - 1) {  // marginal weight from the first D - 1 coordinates
                    float lo, hi;
                    unpack_f32x2(s2, lo, hi);
                    w2 = pack_f32x2(ex2_neg(lo), ex2_neg(hi));
                    facc_m[h] = fadd2(facc_m[h], w2);
                }
                s2 = ffma2(d2, d2, s2);
            }
            float slo, shi, dlo, dhi;
            unpack_f32x2(s2, slo, shi);
            unpack_f32x2(dl2, dlo, dhi);
            const f32x2_t e2 = pack_f32x2(ex2_neg(slo), ex2_neg(shi));  // joint kernel value
            // t = 1 / (1 + |z| / 4),  u = t * scale + shift,  g(u) by Horner,  q = e t g
            const f32x2_t a2 = fmul2(pack_f32x2(fabsf(dlo), fabsf(dhi)), invc2);
            float nlo, nhi;
            unpack_f32x2(ffma2(a2, qinv2, one2), nlo, nhi);
            float tlo, thi;
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(tlo) : "f"(nlo));
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(thi) : "f"(nhi));
            const f32x2_t t2 = pack_f32x2(tlo, thi);
            const f32x2_t u2 = ffma2(t2, qs2, qh2);
            f32x2_t g2 = pack_f32x2(c[kQDeg32], c[kQDeg32]);
#pragma unroll
            for (int k = kQDeg32 - 1; k >= 0; --k) g2 = ffma2(g2, u2, pack_f32x2(c[k], c[k]));
            const f32x2_t q2 = fmul2(e2, fmul2(t2, g2));
            const f32x2_t r2 = fsub2(w2, q2);
            float qlo, qhi, rlo, rhi;
            unpack_f32x2(q2, qlo, qhi);
            unpack_f32x2(r2, rlo, rhi);
            facc_j[h] = fadd2(facc_j[h], pack_f32x2(dlo < 0.f ? qlo : rlo, dhi < 0.f ? qhi : rhi));
        }
    }
#pragma unroll
    for (int h = 0; h < H; ++h) {
        float lo, hi;
        unpack_f32x2(facc_j[h], lo, hi);
        sum_j[2 * h] += static_cast<double>(lo);
        sum_j[2 * h + 1] += static_cast<double>(hi);
        if (CKDE) {
            unpack_f32x2(facc_m[h], lo, hi);
            sum_m[2 * h] += static_cast<double>(lo);
            sum_m[2 * h + 1] += static_cast<double>(hi);
        }
    }
}

// CDF = false: KDE / CKDE log-likelihood sums.  CDF = true: CKDE::cdf sums (see tile_f64); `inv_c` converts a whitened
// (kernel-unit) coordinate difference into standard-normal units and is only read in that mode.
// SHIFT = true: the shifted second pass (tile_f64_shift / tile_f32_shift) over ONE job whose size is only known on the
// device: `dyn` = {total_units, upb} written by shift_prep_kernel (runtime.cu) replaces the by-value arguments.
// Test row of (tile tt, register slot r, thread tid).  Ordinary kernels: slot r covers rows [r * 256, r * 256 + 256) of the
// tile, a warp holds R runs of 32 rows that are 256 apart.  GSKIP: a warp holds R * 32 CONSECUTIVE rows (in Morton order:
// one compact cluster), which is what makes a whole group of training points negligible for every lane more often
// (model on config 2: 38% instead of 42% of the groups contain a significant term).  Loads stay coalesced either way.
template <bool GSKIP, int R>
__device__ __forceinline__ long long pair_row(long long tt, int r, int tid) {
    constexpr int TB = kThreads * R;
    return GSKIP ? tt * TB + (tid >> 5) * (32 * R) + r * 32 + (tid & 31) : tt * TB + r * kThreads + tid;
}
// GSKIP = true (f64, log-likelihood sums, jobs with a unit list): the group-skipping tile, see tile_f64_dot_gskip.
template <typename T, int D, bool CKDE, bool CDF = false, bool SHIFT = false, bool GSKIP = false>
__global__ void __launch_bounds__(kThreads, PairCfg<T>::MIN_CTAS)
pair_kernel(const PairJob* __restrict__ jobs, int n_jobs, long long total_units, long long upb,
            const double* __restrict__ exp_tab_g, double inv_c, const long long* __restrict__ dyn = nullptr) {
    if (SHIFT) {
        total_units = dyn[0];
        upb = dyn[1];
    }
    constexpr int R = CDF ? pair_rows_cdf<T>(D) : pair_rows<T>(D, CKDE);
    constexpr int TILE = pair_tile<T>(D);  // (test rows per tile: kThreads * R, see pair_row)
    constexpr uint32_t TILE_BYTES = TILE * D * sizeof(T);
    constexpr uint32_t NRM_BYTES = pair_nrm_bytes<T>(D);  // per stage; 0 for f32
    constexpr int DN = CKDE ? D - 1 : D;
    // (the evidence-free CDF kernel, D = 1, has no dot-product form: its only coordinate is the conditioned one)
    constexpr bool DOT = !SHIFT && PBN_F64_DOT && sizeof(T) == 8 && DN >= PBN_F64_DOT_MIN_DN && !(CDF && !CKDE);

    extern __shared__ __align__(128) unsigned char smem_raw[];
    T* tile_buf = reinterpret_cast<T*>(smem_raw);  // [kStages][TILE*D]
    double* nrm_buf = reinterpret_cast<double*>(smem_raw + kStages * TILE_BYTES);  // [kStages][TILE]
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw + kStages * (TILE_BYTES + NRM_BYTES));
    double* tab = reinterpret_cast<double*>(smem_raw + kStages * (TILE_BYTES + NRM_BYTES) + 64);

    const int tid = threadIdx.x;
    const long long u0 = static_cast<long long>(blockIdx.x) * upb;
    long long u1 = u0 + upb;
    if (u1 > total_units) u1 = total_units;
    if (u0 >= u1) return;

    if (sizeof(T) == 8) exp_tab_fill(tab, exp_tab_g, tid, kThreads);
    tab += tid & (kExpRep - 1);  // this lane's copy
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(&full_bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // locate the job of the first unit (jobs are sorted by unit_begin)
    int jlo = 0, jhi = n_jobs - 1;
    while (jlo < jhi) {
        int mid = (jlo + jhi + 1) >> 1;
        if (jobs[mid].unit_begin <= u0) jlo = mid; else jhi = mid - 1;
    }

    // producer state (thread 0): next unit to load
    int pj = jlo;
    long long pu = u0;
    auto issue_load = [&](int stage) {
        while (pj + 1 < n_jobs && jobs[pj + 1].unit_begin <= pu) ++pj;
        const PairJob& jb = jobs[pj];
        long long local = pu - jb.unit_begin;
        int nt = jb.unit_list ? jb.unit_list[local] : static_cast<int>(local % jb.n_train_tiles);
        long long start = static_cast<long long>(nt) * TILE;
        long long cnt = jb.n_train - start;
        if (cnt > TILE) cnt = TILE;
        uint32_t bytes = static_cast<uint32_t>(((cnt * D * sizeof(T)) + 15) & ~15ull);
        const T* src = reinterpret_cast<const T*>(jb.train) + start * D;
        uint32_t nbytes = 0;
        if (DOT && jb.train_nrm && jb.test_nrm) nbytes = static_cast<uint32_t>((cnt * sizeof(double) + 15) & ~15ull);
        mbar_expect_tx(&full_bar[stage], bytes + nbytes);
        tma_bulk_g2s(tile_buf + static_cast<size_t>(stage) * TILE * D, src, bytes, &full_bar[stage]);
        if (nbytes) tma_bulk_g2s(nrm_buf + static_cast<size_t>(stage) * TILE, jb.train_nrm + start, nbytes, &full_bar[stage]);
        ++pu;
    };
    if (tid == 0) {
        for (int s = 0; s < kStages && pu < u1; ++s) issue_load(s);
    }

    int cj = jlo;
    T yt[R][D];
    double at[R];
    int ati[R];
    double row_scale[R];
    double sum_j[R], sum_m[R];
    typename std::conditional<sizeof(T) == 8, int, float>::type shj[SHIFT ? R : 1], shm[SHIFT ? R : 1];
    long long cur_tt = -1;
    int cur_job = -1;
    bool safe = true, dot = false;

    auto flush = [&]() {
        if (cur_job < 0) return;
        const PairJob& jb = jobs[cur_job];
        long long ustart = jb.unit_begin + pair_tile_first(jb, cur_tt);
        int slot = static_cast<int>(blockIdx.x - ustart / upb);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            long long row = pair_row<GSKIP, R>(cur_tt, r, tid);
            if (row < jb.m) {
#if PBN_F64_HOIST
                if (DOT && dot) {  // the fraction of -|yt|^2 that tile_f64_dot left out: one factor per test row
                    sum_j[r] *= row_scale[r];
                    if (CKDE) sum_m[r] *= row_scale[r];
                }
#endif
                jb.part[static_cast<long long>(slot) * jb.m_pad + row] = sum_j[r];
                if (CKDE) jb.part[(static_cast<long long>(jb.slots) + slot) * jb.m_pad + row] = sum_m[r];
            }
        }
    };

    for (long long u = u0; u < u1; ++u) {
        const int stage = static_cast<int>((u - u0) % kStages);
        const uint32_t parity = static_cast<uint32_t>(((u - u0) / kStages) & 1);
        while (cj + 1 < n_jobs && jobs[cj + 1].unit_begin <= u) ++cj;
        const PairJob& jb = jobs[cj];
        const long long local = u - jb.unit_begin;
        long long tt;
        int nt;
        if (cj == cur_job && jb.unit_list && local < jb.tile_first[cur_tt + 1]) {  // still inside the current test tile
            tt = cur_tt;
            nt = jb.unit_list[local];
        } else {
            pair_unit(jb, local, tt, nt);
        }
        if (cj != cur_job || tt != cur_tt) {
            flush();
            cur_job = cj;
            cur_tt = tt;
            const T* tp = reinterpret_cast<const T*>(jb.test);
            // the first CTA of a test tile continues the sums the job brings along (PairJob::init_sums)
            const bool inited = !SHIFT && jb.init_sums && jb.unit_begin + pair_tile_first(jb, tt) >= u0;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                long long row = pair_row<GSKIP, R>(tt, r, tid);
                bool ok = row < jb.m;
                const long long src = (SHIFT && ok) ? static_cast<long long>(jb.test_rows[row]) : row;
#pragma unroll
                for (int c = 0; c < D; ++c) yt[r][c] = ok ? tp[src * D + c] : T(0);
                sum_j[r] = 0.0;
                sum_m[r] = 0.0;
                if (!SHIFT && inited && ok) {
                    sum_j[r] = jb.init_sums[row];
                    if (CKDE) sum_m[r] = jb.init_sums[jb.m_pad + row];
                }
                if constexpr (SHIFT) {
                    // shifts >= 2^31 kernel units cannot be carried on the integer side: such a row keeps a zero sum
                    // and falls through to the per-row kernel
                    const float sj = ok ? jb.shift_j[row] : 0.f;
                    const float sm = (ok && CKDE) ? jb.shift_m[row] : 0.f;
                    if constexpr (sizeof(T) == 8) {
                        shj[r] = sj < 2.0e9f ? static_cast<int>(rintf(sj)) : 0;
                        shm[r] = sm < 2.0e9f ? static_cast<int>(rintf(sm)) : 0;
                    } else {
                        shj[r] = sj;
                        shm[r] = sm;
                    }
                }
            }
            if (sizeof(T) == 8 && !SHIFT) {
                // |t| <= D * (max|y_train| + max|y_test|)^2 must stay below 2^31 for the
                // integer-only clamp of the fast exp2 variant
                float a = jb.bound_train ? *jb.bound_train : INFINITY;
                float b = jb.bound_test ? *jb.bound_test : INFINITY;
                float lim = static_cast<float>(D) * (a + b) * (a + b);
                safe = !(lim < 2.0e9f);
                if (DOT) {
                    // cancellation error of the dot-product form, relative per term (see tile_f64_dot):
                    // sqrt(DN+1) * 2^-53 * 2 DN B^2 * ln2/K < kDotTol
                    float B = fmaxf(a, b);
                    float crit = sqrtf(static_cast<float>(DN + 1)) * 2.f * DN * B * B;
                    dot = jb.train_nrm && jb.test_nrm && crit < static_cast<float>(kDotTol * 9007199254740992.0 / kExpA) && !safe;
                    if (dot) {
                        const double* tn = jb.test_nrm;
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            long long row = pair_row<GSKIP, R>(tt, r, tid);
                            at[r] = row < jb.m ? tn[row] : 0.0;
#if PBN_F64_HOIST
                            const double ai = rint(at[r]);  // |at| < 2^31 (the `safe` test above)
                            row_scale[r] = exp((at[r] - ai) * kExpA);
                            ati[r] = static_cast<int>(ai);
                            if (inited) {  // the running sums of this form lack the factor row_scale until flush
                                sum_j[r] /= row_scale[r];
                                if (CKDE) sum_m[r] /= row_scale[r];
                            }
#endif
#pragma unroll
                            for (int c = 0; c < DN; ++c) yt[r][c] *= T(2);
                        }
                    }
                }
            }
        }
        long long cnt_ll = jb.n_train - static_cast<long long>(nt) * TILE;
        const int cnt = cnt_ll > TILE ? TILE : static_cast<int>(cnt_ll);

        mbar_wait(&full_bar[stage], parity);
        const T* __restrict__ tp = tile_buf + static_cast<size_t>(stage) * TILE * D;

        if constexpr (SHIFT && sizeof(T) == 8) {
            tile_f64_shift<D, CKDE, R>(tp, cnt, yt, shj, shm, tab, sum_j, sum_m);
        } else if constexpr (SHIFT) {
            tile_f32_shift<D, CKDE, R>(tp, cnt, yt, shj, shm, sum_j, sum_m);
        } else if constexpr (sizeof(T) == 8) {
            if (DOT && dot) {
                if constexpr (GSKIP && !CDF)
                    tile_f64_dot_gskip<D, CKDE, R>(tp, nrm_buf + static_cast<size_t>(stage) * TILE, cnt, yt, ati, tab, sum_j, sum_m);
                else if (PBN_F64_WARPSKIP && !CDF && jb.unit_list)
                    tile_f64_dot<D, CKDE, R, CDF, true>(tp, nrm_buf + static_cast<size_t>(stage) * TILE, cnt, yt, at, ati, tab, sum_j, sum_m, inv_c);
                else
                    tile_f64_dot<D, CKDE, R, CDF>(tp, nrm_buf + static_cast<size_t>(stage) * TILE, cnt, yt, at, ati, tab, sum_j, sum_m, inv_c);
            }
            else if (safe)
                tile_f64<D, CKDE, true, R, CDF>(tp, cnt, yt, tab, sum_j, sum_m, inv_c);
            else
                tile_f64<D, CKDE, false, R, CDF>(tp, cnt, yt, tab, sum_j, sum_m, inv_c);
        } else {
            if constexpr (GSKIP && !CDF && R % 2 == 0)
                tile_f32_packed_gskip<D, CKDE, R>(tp, cnt, yt, sum_j, sum_m);
            else if constexpr (PBN_F32_PACKED && !CDF && R % 2 == 0 && D >= PBN_F32_PACKED_MIN_D)
                tile_f32_packed<D, CKDE, R>(tp, cnt, yt, sum_j, sum_m);
            else if constexpr (PBN_F32_PACKED && CDF && R % 2 == 0)
                tile_f32_packed_cdf<D, CKDE, R>(tp, cnt, yt, sum_j, sum_m, static_cast<float>(inv_c));
            else
                tile_f32<D, CKDE, R, CDF>(tp, cnt, yt, sum_j, sum_m, static_cast<float>(inv_c));
        }

        __syncthreads();  // everyone is done with this stage
        if (tid == 0 && pu < u1) issue_load(stage);
    }
    flush();
}

// Job of the UCV pair-sum kernel (ucv_kernel.cu)
struct UcvJob {
    const void* y;             // whitened rows AoS [n (padded alloc)][D]
    long long n;
    const long long* prefix;   // prefix[tt] = first unit of row tile tt, prefix[n_row_tiles] = total
    int n_row_tiles;
    const float* bound;        // max |whitened coordinate|
    const double* nrm;         // f64 only, may be null: -sum_c y_c^2 per row (dot-product form, see ucv_tile_dot)
    double* partial;           // [grid][2]
    long long unit_begin, unit_end;  // slice of units handled by this launch (multi-GPU split)
};
cudaError_t launch_ucv(int dtype_f64, int D, const UcvJob& job, long long upb, int grid, const double* tab,
                       cudaStream_t s);

}  // namespace pbn
