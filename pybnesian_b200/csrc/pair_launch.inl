// pair_launch.inl — instantiates pair_kernel<PBN_T, D, CKDE> for D = 1..kMaxFastD (10) and provides
// the launcher for one element type.  Included by pair_f64.cu / pair_f32.cu with
// PBN_T and PBN_LAUNCH_NAME defined (one translation unit per type so they build in parallel).
#include "pair_kernel.cuh"

namespace pbn {

template <int D, bool CKDE, bool CDF = false, bool SHIFT = false, bool GSKIP = false>
static cudaError_t launch_one(const PairJob* jobs, int n_jobs, long long total_units, long long upb, int grid,
                              const double* tab, cudaStream_t stream, double inv_c = 0.0, const long long* dyn = nullptr) {
    constexpr size_t smem = kStages * (pair_tile<PBN_T>(D) * D * sizeof(PBN_T) + pair_nrm_bytes<PBN_T>(D)) + 64 + exp_tab_smem_bytes<PBN_T>();
    static_assert(smem <= 113 * 1024 || PairCfg<PBN_T>::MIN_CTAS < 2, "two CTAs per SM must fit in shared memory");
    auto kern = pair_kernel<PBN_T, D, CKDE, CDF, SHIFT, GSKIP>;
    // set on every launch: the attribute is per device (and per context), and the call is a host-side table update
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    kern<<<grid, kThreads, smem, stream>>>(jobs, n_jobs, total_units, upb, tab, inv_c, dyn);
    return cudaGetLastError();
}

#ifdef PBN_LAUNCH_NAME
cudaError_t PBN_LAUNCH_NAME(int D, bool ckde, const PairJob* jobs, int n_jobs, long long total_units, long long upb,
                            int grid, const double* tab, cudaStream_t stream) {
#define PBN_CASE(d)                                                                                          \
    case d:                                                                                                  \
        return ckde ? launch_one<(d < 2 ? 2 : d), true>(jobs, n_jobs, total_units, upb, grid, tab, stream)   \
                    : launch_one<d, false>(jobs, n_jobs, total_units, upb, grid, tab, stream);
    switch (D) {
        PBN_CASE(1) PBN_CASE(2) PBN_CASE(3) PBN_CASE(4) PBN_CASE(5) PBN_CASE(6) PBN_CASE(7) PBN_CASE(8)
        PBN_CASE(9) PBN_CASE(10)
        default:
            return cudaErrorInvalidValue;
    }
#undef PBN_CASE
}

// CDF mode of the same kernel (CKDE::cdf): D = joint dimension; D == 1 is the evidence-free case (w = 1).
cudaError_t PBN_CDF_LAUNCH_NAME(int D, const PairJob* jobs, int n_jobs, long long total_units, long long upb, int grid,
                                const double* tab, double inv_c, cudaStream_t stream) {
#define PBN_CASE(d) \
    case d:         \
        return launch_one<d, true, true>(jobs, n_jobs, total_units, upb, grid, tab, stream, inv_c);
    switch (D) {
        case 1:
            return launch_one<1, false, true>(jobs, n_jobs, total_units, upb, grid, tab, stream, inv_c);
        PBN_CASE(2) PBN_CASE(3) PBN_CASE(4) PBN_CASE(5) PBN_CASE(6) PBN_CASE(7) PBN_CASE(8)
        PBN_CASE(9) PBN_CASE(10)
        default:
            return cudaErrorInvalidValue;
    }
#undef PBN_CASE
}

int PBN_TILE_NAME(int D) { return pair_tile<PBN_T>(D); }
int PBN_TB_NAME() { return kThreads * PairCfg<PBN_T>::R; }
int PBN_TB_FOR_NAME(int D, bool ckde) { return kThreads * pair_rows<PBN_T>(D, ckde); }
int PBN_TB_CDF_NAME(int D) { return kThreads * pair_rows_cdf<PBN_T>(D); }
int PBN_CTAS_NAME() { return PairCfg<PBN_T>::MIN_CTAS; }
// touches one kernel of this translation unit so that its module is loaded (pbn_ctx_warmup)
cudaError_t PBN_WARM_NAME() {
    cudaFuncAttributes a;
    return cudaFuncGetAttributes(&a, pair_kernel<PBN_T, 4, true, false, false>);
}

#endif  // PBN_LAUNCH_NAME

#ifdef PBN_SHIFT_LAUNCH_NAME
// Shifted second pass (rows whose unshifted sums underflowed): ONE job in device memory, its unit count in `dyn`
// (both written on the device by shift_prep_kernel, runtime.cu); `grid` is the full persistent grid.
cudaError_t PBN_SHIFT_LAUNCH_NAME(int D, bool ckde, const PairJob* job, const long long* dyn, int grid, const double* tab,
                                  cudaStream_t stream) {
#define PBN_CASE(d)                                                                                                   \
    case d:                                                                                                           \
        return ckde ? launch_one<(d < 2 ? 2 : d), true, false, true>(job, 1, 0, 1, grid, tab, stream, 0.0, dyn)       \
                    : launch_one<d, false, false, true>(job, 1, 0, 1, grid, tab, stream, 0.0, dyn);
    switch (D) {
        PBN_CASE(1) PBN_CASE(2) PBN_CASE(3) PBN_CASE(4) PBN_CASE(5) PBN_CASE(6) PBN_CASE(7) PBN_CASE(8)
        PBN_CASE(9) PBN_CASE(10)
        default:
            return cudaErrorInvalidValue;
    }
#undef PBN_CASE
}

cudaError_t PBN_SHIFT_WARM_NAME() {
    cudaFuncAttributes a;
    return cudaFuncGetAttributes(&a, pair_kernel<PBN_T, 4, true, false, true>);
}

#endif  // PBN_SHIFT_LAUNCH_NAME

#ifdef PBN_GSKIP_LAUNCH_NAME
// Pass B of tile skipping with group skipping inside the units (f64; families of up to 8 variables: 9 and 10 run two rows
// per thread and keep the plain kernel).
cudaError_t PBN_GSKIP_LAUNCH_NAME(int D, bool ckde, const PairJob* jobs, int n_jobs, long long total_units, long long upb,
                                  int grid, const double* tab, cudaStream_t stream) {
#define PBN_CASE(d)                                                                                                              \
    case d:                                                                                                                      \
        return ckde ? launch_one<(d < 2 ? 2 : d), true, false, false, true>(jobs, n_jobs, total_units, upb, grid, tab, stream)   \
                    : launch_one<d, false, false, false, true>(jobs, n_jobs, total_units, upb, grid, tab, stream);
    switch (D) {
        PBN_CASE(1) PBN_CASE(2) PBN_CASE(3) PBN_CASE(4) PBN_CASE(5) PBN_CASE(6) PBN_CASE(7) PBN_CASE(8)
        default:
            return cudaErrorInvalidValue;
    }
#undef PBN_CASE
}
cudaError_t PBN_GSKIP_WARM_NAME() {
    cudaFuncAttributes a;
    return cudaFuncGetAttributes(&a, pair_kernel<PBN_T, 4, true, false, false, true>);
}
#endif  // PBN_GSKIP_LAUNCH_NAME

}  // namespace pbn
