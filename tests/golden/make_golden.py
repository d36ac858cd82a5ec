#!/usr/bin/env python3
"""Generates the committed golden vectors under tests/golden/.

Run in the build container (needs /root/reference for oracle/_ref):
    python tests/golden/make_golden.py

Sources of truth, in order of authority:
  * `ref_*`  : outputs of the REFERENCE'S OWN OpenCL-C kernels
               (/root/reference/pybnesian/kde/opencl_kernels/KDE.cl.src) compiled as C++ through
               oracle/ref_shim and driven by the restated host logic (oracle/_ref/libref_kernels.so);
  * `scipy_*`: scipy.stats.gaussian_kde.logpdf with the normal-reference factor, i.e. the oracle the
               reference's own tests use (tests/factors/continuous/KDE_test.py:167-203,
               CKDE_test.py:146-179) — float64 only;
  * shuffles : libstdc++ std::shuffle(std::mt19937{seed}) index vectors
               (dataset/crossvalidation_adaptator.hpp:15-67), cross-checked against the known answers
               recorded in SURVEY.md §7.
Inputs are regenerated from seeds by tests/util_data.py (same recipe as the reference's
tests/helpers/util_test.py:5-19), so only outputs are stored.
"""
import os
import sys

import numpy as np
from scipy.stats import gaussian_kde

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
import util_data  # noqa: E402

VARSETS = [["a"], ["b", "a"], ["c", "a", "b"], ["d", "a", "b", "c"]]
CASES = [(500, 50), (10, 50), (300, 70)]  # (N, m): column variant, row variant (N <= chunk), 2 chunks


def nr_factor(s):
    return np.power(4 / (s.d + 2), 1 / (s.d + 4)) * s.scotts_factor()


def main():
    assert oracle.ref_available(), "build oracle/_ref first (make -C oracle ref)"
    out = {}
    for dt in ("float64", "float32"):
        for variables in VARSETS:
            for N, m in CASES:
                if N <= len(variables):
                    continue
                X = util_data.generate_normal_data(N, 0)[variables].to_numpy().astype(dt)
                T = util_data.generate_normal_data(m, 1)[variables].to_numpy().astype(dt)
                H = oracle.bandwidth(X)
                key = "%s_%s_%d_%d" % (dt, "".join(variables), N, m)
                out["H_" + key] = H
                logl, slogl = oracle.ref_kde_logl(X, T, H)
                out["ref_kde_logl_" + key] = logl
                out["ref_kde_slogl_" + key] = np.array(slogl)
                cl, cs = oracle.ref_ckde_logl(X, T, H)
                out["ref_ckde_logl_" + key] = cl
                out["ref_ckde_slogl_" + key] = np.array(cs)
                if dt == "float64":
                    sk = gaussian_kde(X.T, bw_method=nr_factor)
                    out["scipy_kde_logl_" + key] = sk.logpdf(T.T)
                    out["scipy_H_" + key] = sk.covariance
                    if len(variables) > 1:
                        km = gaussian_kde(X[:, 1:].T, bw_method=sk.covariance_factor())
                        out["scipy_ckde_logl_" + key] = sk.logpdf(T.T) - km.logpdf(T[:, 1:].T)
    for dt in ("float64", "float32"):
        for variables in VARSETS:
            X = util_data.generate_normal_data(200, 0)[variables].to_numpy().astype(dt)
            H = oracle.bandwidth(X)
            key = "%s_%s_200" % (dt, "".join(variables))
            out["ref_ucv_" + key] = np.array(oracle.ref_ucv_score_unconstrained(X, H))
            out["ref_ucv_half_" + key] = np.array(oracle.ref_ucv_score_unconstrained(X, 0.5 * H))
    for n, k, seed in [(10, 3, 0), (23, 10, 0), (1000, 10, 0), (1000, 7, 123), (100000, 10, 0)]:
        idx, lim = oracle.cv_indices(np.arange(n), k, seed)
        out["cv_idx_%d_%d_%d" % (n, k, seed)] = idx if n <= 1000 else idx[:64]
        out["cv_lim_%d_%d_%d" % (n, k, seed)] = lim
        if n > 1000:
            out["cv_idxsum_%d_%d_%d" % (n, k, seed)] = np.array([int(np.sum(idx.astype(np.int64) * np.arange(1, n + 1) % 1000003))])
    assert list(out["cv_idx_10_3_0"]) == [0, 2, 1, 5, 9, 8, 4, 7, 6, 3]
    assert list(out["cv_idx_23_10_0"]) == [10, 4, 6, 12, 22, 1, 9, 17, 7, 20, 11, 18, 19, 15, 21, 3, 14, 13, 2, 0, 16, 8, 5]
    assert list(out["cv_idx_1000_10_0"][:10]) == [882, 396, 136, 545, 569, 298, 709, 664, 519, 504]
    np.savez_compressed(os.path.join(HERE, "kde_golden.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
