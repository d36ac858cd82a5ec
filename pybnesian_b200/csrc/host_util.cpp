// host_util.cpp — host-side integer logic of the path that must be BIT-EXACT with the
// reference (fold indices, operator order, parent order).  The reference gets these from
// libstdc++ (std::shuffle + std::mt19937, std::sort, std::unordered_set<int>), so the only
// faithful restatement is to call the same library routines; this file does exactly that
// behind the C ABI (include/pbn_cuda.h).  No CUDA here.
#include <stdint.h>

#include <algorithm>
#include <cmath>
#include <numeric>
#include <random>
#include <string>
#include <unordered_set>
#include <vector>

#include "../../include/pbn_cuda.h"

int pbn_set_error(int code, const std::string& msg);

struct pbn_intset {
    std::unordered_set<int> s;
};

extern "C" {

// CrossValidationProperties (dataset/crossvalidation_adaptator.hpp:15-67)
int pbn_cv_split(int32_t* indices, int64_t n, int k, uint32_t seed, int32_t* limits) {
    if (!indices || !limits) return pbn_set_error(PBN_ERR_ARG, "null argument");
    // the reference checks k against df->num_rows() (all rows, nulls included) before the null rows are dropped
    // (crossvalidation_adaptator.hpp:24-29), so k may exceed the number of VALID rows passed here and the trailing
    // folds are then empty; that check is the caller's (dataset.py), only k <= 1 is rejected here
    if (k <= 1 || n < 0)
        return pbn_set_error(PBN_ERR_ARG, "Cannot split " + std::to_string(n) + " instances into " + std::to_string(k) +
                                              " folds.");
    std::vector<int> v(indices, indices + n);
    std::mt19937 rng{seed};
    std::shuffle(v.begin(), v.end(), rng);
    std::copy(v.begin(), v.end(), indices);
    int fold_size = static_cast<int>(v.size() / k);
    int folds_extra = static_cast<int>(v.size() % k);
    int cur = 0, pos = 0;
    limits[pos++] = 0;
    for (int i = 0; i < folds_extra; ++i) {
        cur += fold_size + 1;
        limits[pos++] = cur;
    }
    for (int i = folds_extra; i < k; ++i) {
        cur += fold_size;
        limits[pos++] = cur;
    }
    return PBN_OK;
}

// factors::discrete::discrete_slice_indices (factors/discrete/discrete_indices.cpp:166-201): a counting sort
// of the participating rows by configuration index; rows keep ascending order inside a configuration.
int pbn_discrete_slices(const int32_t* const* codes, const int32_t* strides, int nvars, int64_t nrows,
                        const uint8_t* valid, int num_factors, int32_t* order_out, int64_t* offsets_out) {
    if ((nvars > 0 && (!codes || !strides)) || !offsets_out || (nrows > 0 && !order_out))
        return pbn_set_error(PBN_ERR_ARG, "null argument");
    if (nvars < 0 || nrows < 0 || num_factors < 1) return pbn_set_error(PBN_ERR_ARG, "invalid slice arguments");
    std::vector<int32_t> conf(nrows, -1);
    std::vector<int64_t> count(num_factors + 1, 0);
    for (int64_t r = 0; r < nrows; ++r) {
        if (valid && !valid[r]) continue;
        int32_t idx = 0;
        for (int v = 0; v < nvars; ++v) idx += codes[v][r] * strides[v];
        if (idx < 0 || idx >= num_factors)
            return pbn_set_error(PBN_ERR_ARG, "discrete configuration index out of range at row " + std::to_string(r));
        conf[r] = idx;
        ++count[idx + 1];
    }
    offsets_out[0] = 0;
    for (int c = 0; c < num_factors; ++c) offsets_out[c + 1] = offsets_out[c] + count[c + 1];
    std::vector<int64_t> pos(offsets_out, offsets_out + num_factors);
    for (int64_t r = 0; r < nrows; ++r)
        if (conf[r] >= 0) order_out[pos[conf[r]]++] = static_cast<int32_t>(r);
    return PBN_OK;
}

// HoldOut (dataset/holdout_adaptator.hpp:17-70)
int pbn_holdout_split(int32_t* indices, int64_t n, double test_ratio, uint32_t seed, int32_t* n_train) {
    if (!indices || !n_train) return pbn_set_error(PBN_ERR_ARG, "null argument");
    if (test_ratio <= 0 || test_ratio >= 1.0)
        return pbn_set_error(PBN_ERR_ARG, "test_ratio must be a number between 0 and 1.");
    std::vector<int> v(indices, indices + n);
    std::mt19937 rng{seed};
    std::shuffle(v.begin(), v.end(), rng);
    std::copy(v.begin(), v.end(), indices);
    int test_rows = static_cast<int>(std::round(v.size() * test_ratio));
    int train_rows = static_cast<int>(v.size()) - test_rows;
    *n_train = train_rows;
    if (test_rows == 0 || train_rows == 0)
        return pbn_set_error(PBN_ERR_ARG, "Wrong test_ratio (" + std::to_string(test_ratio) +
                                              "selected for HoldOut.\nGenerated train instances: " +
                                              std::to_string(train_rows) +
                                              "\nGenerated test instances: " + std::to_string(test_rows));
    return PBN_OK;
}

// ArcOperatorSet::find_max_indegree (learning/operators/operators.hpp:489-497): the persistent
// index vector is re-sorted by delta, descending, with the (unstable) std::sort.
int pbn_sort_desc(int32_t* idx, int64_t n, const double* delta) {
    if (n > 0 && (!idx || !delta)) return pbn_set_error(PBN_ERR_ARG, "null argument");
    std::vector<int> v(idx, idx + n);
    auto delta_ptr = delta;
    std::sort(v.begin(), v.end(), [&delta_ptr](auto i1, auto i2) { return delta_ptr[i1] > delta_ptr[i2]; });
    std::copy(v.begin(), v.end(), idx);
    return PBN_OK;
}

// LinearGaussianCPD::sample (factors/continuous/LinearGaussianCPD.cpp:317-372): the noise comes from
// std::normal_distribution<double>(beta0, sqrt(variance)) on std::mt19937{seed}; the evidence terms are then
// added column by column (double accumulation, float evidence widened per element).
int pbn_lg_sample(const double* beta, double variance, int p, const void* const* ev, int ev_dtype, int64_t n,
                  uint32_t seed, double* out) {
    if (n < 0) return pbn_set_error(PBN_ERR_ARG, "n should be a non-negative number");
    if (!beta || (n > 0 && !out) || (p > 0 && !ev)) return pbn_set_error(PBN_ERR_ARG, "null argument");
    std::mt19937 rng{seed};
    std::normal_distribution<> normal(beta[0], std::sqrt(variance));
    for (int64_t i = 0; i < n; ++i) out[i] = normal(rng);
    for (int j = 0; j < p; ++j) {
        if (ev_dtype == PBN_F64) {
            const double* e = static_cast<const double*>(ev[j]);
            for (int64_t i = 0; i < n; ++i) out[i] += beta[j + 1] * e[i];
        } else {
            const float* e = static_cast<const float*>(ev[j]);
            for (int64_t i = 0; i < n; ++i) out[i] += beta[j + 1] * e[i];
        }
    }
    return PBN_OK;
}

int pbn_uniform_real(int64_t n, uint32_t seed, int dtype, void* out) {
    if (n < 0 || (n > 0 && !out)) return pbn_set_error(PBN_ERR_ARG, "invalid argument");
    std::mt19937 rng{seed};
    if (dtype == PBN_F64) {
        std::uniform_real_distribution<double> u(0, 1);
        for (int64_t i = 0; i < n; ++i) static_cast<double*>(out)[i] = u(rng);
    } else {
        std::uniform_real_distribution<float> u(0, 1);
        for (int64_t i = 0; i < n; ++i) static_cast<float*>(out)[i] = u(rng);
    }
    return PBN_OK;
}

// std::unordered_set<int>: DNode::m_parents / m_children (graph/graph_types.hpp:12-51).  The order in
// which BayesianNetwork::parents() lists a node's parents is this container's iteration order.
int pbn_intset_new(pbn_intset** out) {
    if (!out) return pbn_set_error(PBN_ERR_ARG, "null argument");
    *out = new pbn_intset();
    return PBN_OK;
}
int pbn_intset_clone(const pbn_intset* s, pbn_intset** out) {
    if (!s || !out) return pbn_set_error(PBN_ERR_ARG, "null argument");
    *out = new pbn_intset(*s);
    return PBN_OK;
}
int pbn_intset_free(pbn_intset* s) {
    delete s;
    return PBN_OK;
}
int pbn_intset_insert(pbn_intset* s, int v) {
    if (!s) return pbn_set_error(PBN_ERR_ARG, "null argument");
    s->s.insert(v);
    return PBN_OK;
}
int pbn_intset_erase(pbn_intset* s, int v) {
    if (!s) return pbn_set_error(PBN_ERR_ARG, "null argument");
    s->s.erase(v);
    return PBN_OK;
}
int pbn_intset_clear(pbn_intset* s) {
    if (!s) return pbn_set_error(PBN_ERR_ARG, "null argument");
    s->s.clear();
    return PBN_OK;
}
int pbn_intset_contains(const pbn_intset* s, int v) { return s && s->s.count(v) ? 1 : 0; }
int pbn_intset_size(const pbn_intset* s) { return s ? static_cast<int>(s->s.size()) : 0; }
int pbn_intset_list(const pbn_intset* s, int* out) {
    if (!s || !out) return pbn_set_error(PBN_ERR_ARG, "null argument");
    int i = 0;
    for (auto v : s->s) out[i++] = v;
    return PBN_OK;
}

}  // extern "C"
