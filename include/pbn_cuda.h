/* pbn_cuda.h — C ABI of libpbn_cuda.so: the B200 (sm_100a) device layer that replaces
 * PyBNesian's OpenCL layer for the KDE / CKDE log-likelihood hot path.
 *
 * Every entry point cites the reference interface it stands in for (paths relative to
 * /root/reference/pybnesian/).  Plain pointers and sizes only; no C++/torch types.
 *
 * Conventions
 *   - every function returns PBN_OK (0) or a PBN_ERR_* code; pbn_last_error() gives the
 *     message of the last failure on the calling thread.
 *   - dtype: PBN_F64 / PBN_F32 (the Arrow column type; all columns of a table agree,
 *     dataset/dataset.cpp:253-271).
 *   - host matrices (bandwidths, covariances) are column-major double, d x d.
 *   - a "row range" is two half-open segments [b0,e0) ++ [b1,e1) of table rows: a
 *     cross-validation training set is exactly that on a table stored in shuffled order
 *     (dataset/crossvalidation_adaptator.cpp:5-31); pass b1 == e1 for a single segment.
 *   - tables hold null-free columns: dropping rows with nulls is done by the caller,
 *     as DataFrame::to_eigen does for the reference (dataset/dataset.hpp:236-338).
 */
#ifndef PBN_CUDA_H
#define PBN_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PBN_OK 0
#define PBN_ERR_CUDA 1        /* a CUDA runtime call failed  -> RuntimeError (opencl_config.hpp:19-27) */
#define PBN_ERR_ARG 2         /* invalid argument            -> ValueError */
#define PBN_ERR_SINGULAR 3    /* SingularCovarianceData      (util/exceptions.hpp, NormalReferenceRule.hpp:37-60) */
#define PBN_ERR_UNSUPPORTED 4

#define PBN_F64 0
#define PBN_F32 1

#define PBN_BW_NORMAL_REFERENCE 0 /* kde/NormalReferenceRule.hpp:109-134 */
#define PBN_BW_SCOTT 1            /* kde/ScottsBandwidth.hpp:91-117 */

#define PBN_FACTOR_CKDE 0            /* factors/continuous/CKDE.hpp */
#define PBN_FACTOR_LINEAR_GAUSSIAN 1 /* factors/continuous/LinearGaussianCPD.hpp */

#define PBN_MAX_DIM 32

typedef struct pbn_ctx pbn_ctx;     /* one GPU: device id, stream, scratch memory          */
typedef struct pbn_table pbn_table; /* resident column store (uploaded once, kept in HBM)  */
typedef struct pbn_kde pbn_kde;     /* fitted KDE or CKDE: whitened training rows on device */
typedef struct pbn_ucv pbn_ucv;     /* UCVScorer: training rows + scratch resident on device */
typedef struct pbn_cv pbn_cv;       /* CrossValidation / HoldOut over a resident table: shuffled-order copy + fold statistics */
typedef struct pbn_intset pbn_intset; /* std::unordered_set<int> (parent / children sets of the DAG) */

typedef struct pbn_rows {
    int64_t b0, e0, b1, e1;
} pbn_rows;

const char* pbn_last_error(void);
const char* pbn_version(void);
int pbn_device_count(int* out);

/* Replaces opencl::OpenCLConfig::get() (opencl/opencl_config.cpp:149-220): one context
 * per GPU instead of a process-wide platform-0/device-0 singleton. */
int pbn_ctx_create(int device, pbn_ctx** out);
/* One context over n GPUs of THIS process (the form SURVEY.md 8b sketches).  The reference is driven from a single
 * process and a single device (opencl/opencl_config.hpp:120-121, opencl_config.cpp:217-220); with this context the same
 * single-process caller uses every listed B200: tables, fitted KDEs / CKDEs, pbn_cv and pbn_ucv objects created through
 * it hold one replica per device (uploaded over each device's own PCIe link in parallel), and
 *   pbn_kde_logl / pbn_ckde_cdf   shard the test rows (contiguous ranges, training set replicated),
 *   pbn_cv_scores / _score_jobs   deal the (candidate, fold) jobs, most expensive first,
 *   pbn_ucv_score / _pair_sums    cut the pair-tile schedule once more per device,
 * one host thread per device; per-device scalars are added on the host in device order, so a result depends on the
 * number of devices (last-bit rounding of the partial sums) but not on timing.  There is no device-to-device traffic.
 * Every other entry point runs on devices[0].  Small calls stay on devices[0] as well. */
int pbn_ctx_create_multi(const int* devices, int n, pbn_ctx** out);
/* Loads the kernel modules on every device of the context now instead of at the first compute call (~2 s of CUDA
 * module loading for the instantiations of the pair kernel); may be called from another host thread while the caller
 * prepares its data. The reference pays the equivalent - OpenCL program build - lazily per kernel
 * (opencl/opencl_config.cpp:23-147). */
int pbn_ctx_warmup(pbn_ctx* ctx);
int pbn_ctx_num_devices(pbn_ctx* ctx);
int pbn_ctx_device(pbn_ctx* ctx, int i); /* CUDA ordinal of the i-th device of the context, -1 if out of range */
int pbn_ctx_destroy(pbn_ctx* ctx);
/* Use an externally owned cudaStream_t (e.g. torch's current stream); NULL restores the own stream. */
int pbn_ctx_set_stream(pbn_ctx* ctx, void* cuda_stream);
void* pbn_ctx_stream(pbn_ctx* ctx);
int pbn_ctx_synchronize(pbn_ctx* ctx);
int pbn_ctx_sm_count(pbn_ctx* ctx);
/* Counters since context creation: kernel launches issued by this library, host->device
 * and device->host bytes it copied. */
int pbn_ctx_counters(pbn_ctx* ctx, int64_t* launches, int64_t* h2d_bytes, int64_t* d2h_bytes);

/* Device timing of the pair kernel alone: when enabled, CUDA events are recorded on the
 * context's stream around every pair-kernel launch; pbn_ctx_pair_kernel_time() returns the
 * accumulated kernel milliseconds, the number of timed launches and the train x test
 * pair evaluations they processed (a CKDE pair counts 2: joint + marginal). The reference has
 * no profiling hooks (its queue is created without CL_QUEUE_PROFILING_ENABLE,
 * opencl/opencl_config.cpp:175). */
int pbn_ctx_set_timing(pbn_ctx* ctx, int on);
int pbn_ctx_pair_kernel_time(pbn_ctx* ctx, double* total_ms, int64_t* n_launches, int64_t* pair_evals, int reset);

/* Replaces OpenCLConfig::copy_to_buffer (opencl/opencl_config.hpp:226-239) applied to
 * DataFrame::to_eigen output: uploads `ncols` dense host columns of `nrows` values once;
 * the table stays resident until freed. */
int pbn_table_upload(pbn_ctx* ctx, const void* const* col_ptrs, int ncols, int64_t nrows, int dtype,
                     pbn_table** out);
int pbn_table_free(pbn_table* tbl);
int64_t pbn_table_rows(const pbn_table* tbl);
int pbn_table_cols(const pbn_table* tbl);
/* Replaces OpenCLConfig::read_from_buffer (opencl_config.hpp:241-250) for KDE pickling
 * (kde/KDE.hpp:642-666): copies rows of one column back to the host. */
int pbn_table_download(pbn_ctx* ctx, const pbn_table* tbl, int col, pbn_rows rows, void* out);

/* Sample mean and unbiased covariance of the selected columns over a row range
 * (DataFrame::cov, dataset/dataset.hpp:341-396), computed on the device (two passes,
 * fp64 accumulation) and returned to the host. */
int pbn_table_moments(pbn_ctx* ctx, const pbn_table* tbl, const int* cols, int d, pbn_rows rows,
                      double* mean_out, double* cov_out);

/* NormalReferenceRule / ScottsBandwidth ::bandwidth (kde/NormalReferenceRule.hpp:37-60,109-134;
 * kde/ScottsBandwidth.hpp:33-56,91-117).  Fails with PBN_ERR_SINGULAR when rows <= d or the
 * covariance is not positive definite (util/basic_eigen_ops.hpp:136-147). */
int pbn_bandwidth(pbn_ctx* ctx, const pbn_table* tbl, const int* cols, int d, pbn_rows rows, int rule,
                  double* H_out);
/* diag_bandwidth of the same selectors (NormalReferenceRule.hpp:72-106; ScottsBandwidth.hpp:66-89). */
int pbn_diag_bandwidth(pbn_ctx* ctx, const pbn_table* tbl, const int* cols, int d, pbn_rows rows, int rule,
                       double* h_out);

/* KDE::fit<>(bandwidth, buffer, type, N) / KDE::_fit (kde/KDE.hpp:451-511): Cholesky of H,
 * lognorm constant, training rows made resident (whitened with the Cholesky factor).
 * `cols[0..d)` are table columns in the KDE's variable order. */
int pbn_kde_fit(pbn_ctx* ctx, const pbn_table* tbl, const int* cols, int d, pbn_rows rows, const double* H,
                pbn_kde** out);
/* CKDE::_fit (factors/continuous/CKDE.hpp:182-200): cols[0] = variable, cols[1..d) = evidence,
 * H = joint bandwidth; the marginal uses H[1:,1:] on the same rows.  d == 1 is a plain KDE. */
int pbn_ckde_fit(pbn_ctx* ctx, const pbn_table* tbl, const int* cols, int d, pbn_rows rows, const double* H,
                 pbn_kde** out);
/* ProductKDE::_fit (kde/ProductKDE.hpp:153-193): diagonal bandwidth h[0..d) (the *variances* returned by
 * BandwidthSelector::diag_bandwidth), per-variable scale sqrt(h_i) (ProductKDE.cpp:6-29),
 * lognorm = -d/2 log(2 pi) - 1/2 sum log h_i - log N.  The fitted object is evaluated with pbn_kde_logl
 * (replaces logl_values_1d_mat + add_logl_values_1d_mat per variable and the column log-sum-exp,
 * ProductKDE.hpp:233-296, KDE.cl.src:143-170). */
int pbn_product_kde_fit(pbn_ctx* ctx, const pbn_table* tbl, const int* cols, int d, pbn_rows rows, const double* h,
                        pbn_kde** out);
int pbn_kde_free(pbn_kde* kde);
int64_t pbn_kde_num_instances(const pbn_kde* kde);
double pbn_kde_lognorm(const pbn_kde* kde); /* joint lognorm (kde/KDE.hpp:476-477) */

/* KDE::logl / slogl (kde/KDE.hpp:513-640) and CKDE::logl / slogl
 * (factors/continuous/CKDE.hpp:202-287) for a fitted model: one fused launch over
 * train x test tiles.  `cols` index the test table in the same variable order as fit.
 * out_logl (host, rows.count doubles) and out_slogl (host, 1 double) may each be NULL. */
int pbn_kde_logl(pbn_ctx* ctx, const pbn_kde* kde, const pbn_table* test, const int* cols, pbn_rows rows,
                 double* out_logl, double* out_slogl);
/* Asynchronous form for benchmarking / pipelining: result stays on the device
 * (d_out_logl: rows.count doubles or NULL; d_out_slogl: 1 double or NULL), no host sync. */
int pbn_kde_logl_device(pbn_ctx* ctx, const pbn_kde* kde, const pbn_table* test, const int* cols, pbn_rows rows,
                        double* d_out_logl, double* d_out_slogl);
/* Number of test rows of the last logl call on this context that needed the shifted
 * (max-subtracted) re-evaluation because their unshifted kernel sum underflowed: a second,
 * device-scheduled pass of the tiled pair kernel with a per-row exponent shift - what
 * OpenCLConfig::logsumexp_cols_offset (opencl/opencl_config.hpp:517-536) does for every row. */
int pbn_ctx_last_fallback_rows(pbn_ctx* ctx, int64_t* out);
/* ... and how many of those the shifted pass could not evaluate either (farther than 2^31
 * kernel units from every training row, non-finite coordinates): one CTA per row. */
int pbn_ctx_last_row_kernel_rows(pbn_ctx* ctx, int64_t* out);
/* Tile skipping (on by default).  The reference evaluates every (train, test) pair (kde/KDE.hpp:592-640).  For large
 * single-model calls (>= 2^17 training rows, >= 2^14 test rows, d <= 10) this library keeps the whitened rows in Morton
 * order, one bounding box per tile, and drops every (test tile, train tile) unit whose boxes prove that ALL dropped
 * terms of a row together stay below 2^-48 of that row's sum (lower-bounded by a first pass against the nearest
 * training tile): logl changes by < 4e-15 absolute.  pbn_ctx_set_skipping(ctx, 0) evaluates every pair, which is what
 * bench.py's `value` and the roofline are measured on.  pbn_ctx_skip_stats reports (test tile x train tile) units in
 * total / actually evaluated, for the last call and accumulated over the timed launches. */
int pbn_ctx_set_skipping(pbn_ctx* ctx, int on);
int pbn_ctx_skip_stats(pbn_ctx* ctx, int64_t* last_total, int64_t* last_done, int64_t* timed_total, int64_t* timed_done,
                       int reset);

/* kde::UCVScorer (kde/UCV.hpp:12-45): constructed once per (data, variables); every score call
 * evaluates all N(N-1)/2 pairs in ONE launch (64-bit pair indexing; the reference's 32-bit chunk
 * offset, UCV.cpp:32,75,136, limits it to N <= 92 682). */
int pbn_ucv_create(pbn_ctx* ctx, const pbn_table* tbl, const int* cols, int d, pbn_rows rows, pbn_ucv** out);
int pbn_ucv_free(pbn_ucv* scorer);
/* UCVScorer::score_unconstrained (H: d x d) / score_diagonal (h: d values, is_diag != 0):
 * N * UCV(H) = e^{c2} + 2/N sum_{i<j} e^{-s_ij/4 + c2} - 4/(N-1) sum_{i<j} e^{-s_ij/2 + c1}
 * (kde/UCV.cpp:235-399). */
int pbn_ucv_score(pbn_ucv* scorer, const double* H_or_hdiag, int is_diag, double* out);
/* The two pair sums  S2 = sum e^{-s/4},  S1 = sum e^{-s/2}  over slice `part` of `nparts` equal
 * slices of the tile schedule: the multi-GPU split (each rank one slice, sums all-reduced). */
int pbn_ucv_pair_sums(pbn_ucv* scorer, const double* H_or_hdiag, int is_diag, int part, int nparts, double* S2,
                      double* S1);
/* The score from pair sums that were added over all slices (the scalar epilogue of kde/UCV.cpp:296-304, 357). */
int pbn_ucv_score_from_sums(pbn_ucv* scorer, const double* H_or_hdiag, int is_diag, double S2, double S1, double* out);
int64_t pbn_ucv_pairs(const pbn_ucv* scorer);
/* UCV::bandwidth / UCV::diag_bandwidth (kde/UCV.cpp:452-525): Nelder-Mead over vech(chol(H))
 * (or sqrt of the diagonal) from the normal-reference start, ftol_rel = xtol_rel = 1e-4, with the
 * reference's determinant / score guards.  diagonal != 0 writes d values, else d x d. */
int pbn_ucv_bandwidth(pbn_ctx* ctx, const pbn_table* tbl, const int* cols, int d, pbn_rows rows, int diagonal,
                      double* H_out, int* n_evals);

/* LinearGaussianCPD::fit = MLE<LinearGaussianCPD>::estimate (learning/parameters/mle_LinearGaussianCPD.hpp:11-221):
 * cols[0] = variable, cols[1..d) = parents; writes beta[d] (intercept first) and the variance
 * (+inf when rows <= d, as the reference). */
int pbn_lg_fit(pbn_ctx* ctx, const pbn_table* tbl, const int* cols, int d, pbn_rows rows, double* beta_out,
               double* variance_out);
/* LinearGaussianCPD::logl / slogl (factors/continuous/LinearGaussianCPD.cpp:92-149, 251-292); out_logl
 * (rows.count doubles) and out_slogl may each be NULL. */
int pbn_lg_logl(pbn_ctx* ctx, const pbn_table* tbl, const int* cols, int d, pbn_rows rows, const double* beta,
                double variance, double* out_logl, double* out_slogl);

/* ---- cross-validated scores -------------------------------------------------------------------------
 * CrossValidationProperties (dataset/crossvalidation_adaptator.hpp:15-67): shuffles `indices` (the valid row
 * ids, n of them) in place with std::shuffle(std::mt19937{seed}) and writes the k + 1 fold limits.  Host
 * only, bit-exact with the reference (same libstdc++ routines).  PBN_ERR_ARG ("Cannot split ...") if
 * k <= 1 or k > n. */
int pbn_cv_split(int32_t* indices, int64_t n, int k, uint32_t seed, int32_t* limits);
/* HoldOut (dataset/holdout_adaptator.hpp:17-70): shuffle, test_rows = round(n * test_ratio); the first
 * *n_train shuffled ids are the training rows, the rest the test rows. */
int pbn_holdout_split(int32_t* indices, int64_t n, double test_ratio, uint32_t seed, int32_t* n_train);

/* Device side of dataset::CrossValidation (crossvalidation_adaptator.{hpp,cpp}) for the score classes:
 * copies rows `indices[0..n)` of every column of `tbl` into a resident table in that (shuffled) order, so
 * fold f is the contiguous row range [limits[f], limits[f+1]) and its training set is all other rows, and
 * accumulates per-fold column sums and centred cross products of all columns (one pass).  A HoldOut is the
 * k = 2 case scored on fold 1 only (train = fold 0). */
int pbn_cv_create(pbn_ctx* ctx, const pbn_table* tbl, const int32_t* indices, int64_t n, const int32_t* limits,
                  int k, pbn_cv** out);
int pbn_cv_free(pbn_cv* cv);
const pbn_table* pbn_cv_table(const pbn_cv* cv); /* the shuffled-order table (owned by cv) */
int pbn_cv_folds(const pbn_cv* cv);
/* mean and unbiased covariance of the training rows of `fold` for the columns `vars` (DataFrame::cov of the
 * fold's training DataFrame, dataset/dataset.hpp:341-396), from the fold statistics. */
int pbn_cv_train_moments(const pbn_cv* cv, int fold, const int* vars, int d, double* mean_out, double* cov_out);

/* One candidate conditional distribution: vars[0] = variable, vars[1..n_vars) = evidence (column indices
 * of the cv table), factor = PBN_FACTOR_*, rule = PBN_BW_* (CKDE only). */
typedef struct pbn_cv_item {
    int factor;
    int rule;
    int n_vars;
    int vars[PBN_MAX_DIM];
} pbn_cv_item;
/* CVLikelihood::local_score (learning/scores/cv_likelihood.cpp:11-25) for every item at once:
 * scores[i] = sum over folds f in [fold_begin, fold_end) of slogl_f(fit on the other rows; eval on fold f),
 * accumulated in fold order.  All CKDE (item, fold) jobs with the same number of variables share ONE
 * whitening launch and ONE pair-kernel launch.  status[i] (may be NULL) receives PBN_OK or the error code
 * of item i (PBN_ERR_SINGULAR: SingularCovarianceData of some fold; its score is then meaningless);
 * pbn_last_error() holds the message of the first failed item. */
int pbn_cv_scores(pbn_ctx* ctx, pbn_cv* cv, const pbn_cv_item* items, int n_items, int fold_begin, int fold_end,
                  double* scores, int* status);
/* The same scores one fold at a time: job j = fold job_fold[j] of items[job_item[j]], job_scores[j] = slogl of that
 * fold alone (the summand of cv_likelihood.cpp:19-23).  This is the unit the multi-GPU paths deal: a caller that owns a
 * subset of the (candidate, fold) jobs of a hill-climbing step (one rank of a torchrun job, pybnesian_b200/scores.py)
 * scores exactly that subset in ONE batched launch per family size, and a multi-device context splits any job list
 * over its devices.  The caller adds the folds of an item in fold order. */
int pbn_cv_score_jobs(pbn_ctx* ctx, pbn_cv* cv, const pbn_cv_item* items, int n_items, const int32_t* job_item,
                      const int32_t* job_fold, int n_jobs, double* job_scores, int* status);

/* ---- hybrid factors: one base factor per configuration of the discrete parents -----------------------
 * factors::discrete::discrete_slice_indices (factors/discrete/discrete_indices.cpp:166-201) with
 * discrete_indices (discrete_indices.hpp:52-124): codes[v][r] is the dictionary index of discrete variable v
 * at row r (Arrow DictionaryArray indices widened to int32), valid[r] != 0 marks rows that take part (NULL =
 * all rows; the reference uses the combined validity bitmap of the discrete variables).  The configuration of
 * a row is sum_v codes[v][r] * strides[v]; order_out receives the participating row ids grouped by
 * configuration, ascending inside each group (the reference's per-configuration Int32 index arrays laid end
 * to end), offsets_out[c] .. offsets_out[c + 1] delimits configuration c (num_factors + 1 values).  Host
 * only, integer, bit-exact.  PBN_ERR_ARG if a configuration index falls outside [0, num_factors). */
int pbn_discrete_slices(const int32_t* const* codes, const int32_t* strides, int nvars, int64_t nrows,
                        const uint8_t* valid, int num_factors, int32_t* order_out, int64_t* offsets_out);
/* DataFrame::take (arrow::compute::Take, used once per configuration per call by DiscreteAdaptator::fit /
 * logl / slogl, factors/discrete/DiscreteAdaptator.hpp:243,264,312): gathers rows indices[0..n) of every
 * column of a resident table into a new resident table, on the device. */
int pbn_table_take(pbn_ctx* ctx, const pbn_table* tbl, const int32_t* indices, int64_t n, pbn_table** out);
/* DiscreteAdaptator<CKDE>::logl / slogl (DiscreteAdaptator.hpp:259-325) over a table stored in
 * configuration-major order: job j evaluates kdes[j] on the row range rows[j] of `test`.  All jobs share ONE
 * whitening launch and ONE multi-job pair-kernel launch.  kdes[j] == NULL (configuration without a fitted
 * factor) yields NaN rows and a zero sum.  out_logl (host, sum of the range sizes, job-major) and out_slogl
 * (host, n_jobs sums, to be added by the caller in configuration order) may each be NULL. */
int pbn_kde_logl_multi(pbn_ctx* ctx, const pbn_kde* const* kdes, int n_jobs, const pbn_table* test, const int* cols,
                       const pbn_rows* rows, double* out_logl, double* out_slogl);

/* ---- the other side of a fitted CKDE: cdf and sampling --------------------------------
 * CKDE::cdf (factors/continuous/CKDE.hpp:506-728; kernels univariate_normal_cdf, normal_cdf, conditional_means_*,
 * exp_elementwise, product_elementwise, division_elementwise, KDE.cl.src:241-245, 366-468):
 *   cdf(t) = sum_i w_ti Phi((x_t - mean_ti) / sqrt(cond_var)) / sum_i w_ti,   w_ti = marginal kernel weight,
 * (1/N sum_i Phi((x_t - x_i) / h) without evidence) in ONE fused launch over train x test tiles; `kde` must come
 * from pbn_ckde_fit, `cols` as in pbn_kde_logl.  out: rows.count doubles (host).  As in the reference the weights are
 * not max-shifted: a row whose weights all underflow in the data's dtype yields NaN. */
int pbn_ckde_cdf(pbn_ctx* ctx, const pbn_kde* kde, const pbn_table* test, const int* cols, pbn_rows rows, double* out);
/* CKDE::_sample_indices_from_weights (CKDE.hpp:402-504; kernels accum_sum_mat_cols, add_accum_sum_mat_cols,
 * normalize_accum_sum_mat_cols, find_random_indices, KDE.cl.src:254-364): for every evidence row t the training row
 * i with  cum_t[i] <= u_t * S_t < cum_t[i+1]  (cum = running sum of the marginal kernel weights, S_t their total),
 * N-1 if there is none.  ev_cols: the d-1 evidence columns in the CKDE's evidence order; random_prob: rows.count
 * host values in the data's dtype.  Two launches: weight totals, then an in-order scan that stops early. */
int pbn_ckde_sample_indices(pbn_ctx* ctx, const pbn_kde* kde, const pbn_table* evidence, const int* ev_cols, pbn_rows rows,
                            const void* random_prob, int32_t* out_idx);
/* CKDE::sample (CKDE.cpp:97-121, CKDE.hpp:289-400).  H = joint bandwidth (d x d, variable first); train/train_cols/
 * train_rows = the table the CKDE was fitted on (raw training values of the sampled rows are gathered on the device);
 * evidence/ev_cols = resident evidence table (first n rows are used), ev_host[j] = the same n evidence values on the
 * host in the data's dtype (both ignored when d == 1).  The random streams are the reference's: std::mt19937{seed}
 * with libstdc++'s uniform_int / uniform_real / normal distributions.  out: n values in the data's dtype; idx_out
 * (may be NULL): the sampled training rows. */
int pbn_ckde_sample(pbn_ctx* ctx, const pbn_kde* kde, const double* H, const pbn_table* train, const int* train_cols,
                    pbn_rows train_rows, const pbn_table* evidence, const int* ev_cols, const void* const* ev_host,
                    int64_t n, uint32_t seed, void* out, int32_t* idx_out);
/* LinearGaussianCPD::sample (factors/continuous/LinearGaussianCPD.cpp:317-372): host only, bit-exact with the
 * reference (std::mt19937{seed}, std::normal_distribution<double>(beta[0], sqrt(variance)), then beta[j+1] *
 * evidence_j added column by column).  ev[j]: n host values of dtype ev_dtype. */
int pbn_lg_sample(const double* beta, double variance, int p, const void* const* ev, int ev_dtype, int64_t n,
                  uint32_t seed, double* out);

/* n draws of std::uniform_real_distribution<T>(0, 1) from std::mt19937{seed} (T = double for PBN_F64, float for
 * PBN_F32): the stream DiscreteFactor::sample_indices (factors/discrete/DiscreteFactor.hpp:158-199) and
 * CKDE::_sample_multivariate (CKDE.hpp:336-341) consume.  Host only, bit-exact with the reference. */
int pbn_uniform_real(int64_t n, uint32_t seed, int dtype, void* out);

/* ---- host-side integer logic that must match libstdc++ bit for bit -----------------------------------
 * ArcOperatorSet::find_max_indegree (learning/operators/operators.hpp:489-497): std::sort of the persistent
 * candidate index vector by delta, descending (unstable: ties resolve as in the reference). */
int pbn_sort_desc(int32_t* idx, int64_t n, const double* delta);
/* std::unordered_set<int>, the container behind DNode::parents()/children() (graph/graph_types.hpp:12-51):
 * BayesianNetwork::parents(node) lists parents in this container's iteration order. */
int pbn_intset_new(pbn_intset** out);
int pbn_intset_clone(const pbn_intset* s, pbn_intset** out);
int pbn_intset_free(pbn_intset* s);
int pbn_intset_insert(pbn_intset* s, int v);
int pbn_intset_erase(pbn_intset* s, int v);
int pbn_intset_clear(pbn_intset* s);
int pbn_intset_contains(const pbn_intset* s, int v);
int pbn_intset_size(const pbn_intset* s);
int pbn_intset_list(const pbn_intset* s, int* out);

/* Device scratch for callers that keep results on the GPU (e.g. bench.py). */
int pbn_device_alloc(pbn_ctx* ctx, int64_t bytes, void** out);
int pbn_device_free(pbn_ctx* ctx, void* ptr);
int pbn_device_read(pbn_ctx* ctx, const void* dptr, int64_t bytes, void* host_out);

#ifdef __cplusplus
}
#endif
#endif /* PBN_CUDA_H */
