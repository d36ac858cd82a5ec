#!/usr/bin/env python3
"""Golden vectors for SURVEY §8 f3 (CKDE.cdf, CKDE.sample), same conventions as make_golden.py:
  * `ref_cdf_*`, `ref_idx_*`: outputs of the REFERENCE'S OWN OpenCL-C kernels run through oracle/_ref
    (univariate_normal_cdf, normal_cdf, conditional_means_*, exp_elementwise, product/division_elementwise,
    accum_sum_mat_cols, add_accum_sum_mat_cols, normalize_accum_sum_mat_cols, find_random_indices) driven by the
    restated host logic of CKDE.hpp:402-728;
  * `scipy_cdf_*`: the oracle of the reference's own test (tests/factors/continuous/CKDE_test.py:181-219), float64;
  * `u_*`, `sample_*`, `lg_sample_*`: libstdc++ random streams (std::mt19937 + uniform_real / uniform_int / normal
    distributions) as the reference consumes them (CKDE.hpp:289-400, LinearGaussianCPD.cpp:317-372).
Inputs are regenerated from seeds by tests/util_data.py.

    python tests/golden/make_golden_f3.py        (needs /root/reference for oracle/_ref)
"""
import os
import sys

import numpy as np
from scipy.stats import multivariate_normal as mvn
from scipy.stats import norm

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
import util_data  # noqa: E402

VARSETS = [["a"], ["b", "a"], ["c", "a", "b"], ["d", "a", "b", "c"]]
CASES = [(500, 50), (40, 90), (300, 70)]  # column variant, row variant (N <= chunk), 2 chunks


def scipy_cdf(X, T, H):
    if X.shape[1] == 1:
        return norm.cdf(T[:, :1], X[:, 0][None, :], np.sqrt(H[0, 0])).mean(axis=1)
    inv = np.linalg.inv(H[1:, 1:])
    cond_var = H[0, 0] - H[0, 1:].dot(inv).dot(H[1:, 0])
    out = np.empty(T.shape[0])
    for t in range(T.shape[0]):
        w = np.exp(mvn.logpdf(X[:, 1:], mean=T[t, 1:], cov=H[1:, 1:]))
        cm = X[:, 0] + H[0, 1:].dot(inv).dot((T[t, 1:] - X[:, 1:]).T)
        out[t] = np.dot(w, norm.cdf(T[t, 0], cm, np.sqrt(cond_var))) / w.sum()
    return out


def main():
    assert oracle.ref_available(), "build oracle/_ref first (make -C oracle ref)"
    out = {}
    for dt in ("float64", "float32"):
        for variables in VARSETS:
            for N, m in CASES:
                X = util_data.generate_normal_data(N, 0)[variables].to_numpy().astype(dt)
                T = util_data.generate_normal_data(m, 1)[variables].to_numpy().astype(dt)
                H = oracle.bandwidth(X)
                key = "%s_%s_%d_%d" % (dt, "".join(variables), N, m)
                out["ref_cdf_" + key] = oracle.ref_ckde_cdf(X, T, H)
                if dt == "float64":
                    out["scipy_cdf_" + key] = scipy_cdf(X, T, H)
                if len(variables) > 1:
                    u = oracle.uniform_real(m, 7, dt)
                    out["ref_idx_" + key] = oracle.ref_ckde_sample_indices(X[:, 1:], T[:, 1:], H[1:, 1:], u)
                smp, idx = oracle.ckde_sample(X, H, T[:, 1:] if len(variables) > 1 else None, m, 11)
                out["sample_" + key] = smp
                out["sample_idx_" + key] = idx
        out["u_%s" % dt] = oracle.uniform_real(16, 7, dt)
    beta = np.array([1.5, -0.7, 2.25])
    ev = util_data.generate_normal_data(64, 3)
    out["lg_sample_f64"] = oracle.lg_sample(beta, 0.81, [ev["a"].to_numpy(), ev["b"].to_numpy()], 64, 5)
    out["lg_sample_f32"] = oracle.lg_sample(beta, 0.81, [ev["a"].to_numpy().astype(np.float32),
                                                         ev["b"].to_numpy().astype(np.float32)], 64, 5)
    out["lg_sample_noev"] = oracle.lg_sample(beta[:1], 0.81, [], 64, 5)
    np.savez_compressed(os.path.join(HERE, "f3_golden.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
