#!/usr/bin/env python3
"""Golden vectors for the SURVEY §8(f) "next" rows (ProductKDE, CKDE.cdf, ...), same conventions as
make_golden.py: outputs of the REFERENCE'S OWN OpenCL-C kernels run through oracle/_ref (`ref_*`), and the
SciPy oracles the reference's tests use (`scipy_*`).  Inputs are regenerated from seeds by tests/util_data.py.

    python tests/golden/make_golden_next.py        (needs /root/reference for oracle/_ref)
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
CASES = [(500, 50), (150, 70)]


def scipy_product_logpdf(X, T):
    """tests/factors/continuous/ProductKDE_test.py:186-225 (factor_product_kernel + gaussian_kde)."""
    cov = np.atleast_2d(np.cov(X, rowvar=False, bias=False))
    delta = np.diag(np.reciprocal(np.diag(cov))).dot(cov)
    delta_inv = np.linalg.inv(delta)
    N, d = X.shape
    k = 4 * d * np.sqrt(np.linalg.det(delta)) / (2 * np.trace(np.dot(delta_inv, delta_inv)) + np.trace(delta_inv) ** 2)
    factor = (k / N) ** (1. / (d + 4.))
    sk = gaussian_kde(X.T, bw_method=lambda g: factor * np.eye(d))
    sk.cho_cov = np.linalg.cholesky(sk.covariance)
    sk.log_det = 2 * np.log(np.diag(sk.cho_cov * np.sqrt(2 * np.pi))).sum()
    return sk.logpdf(T.T)


def main():
    assert oracle.ref_available(), "build oracle/_ref first (make -C oracle ref)"
    out = {}
    for dt in ("float64", "float32"):
        for variables in VARSETS:
            for N, m in CASES:
                X = util_data.generate_normal_data(N, 0)[variables].to_numpy().astype(dt)
                T = util_data.generate_normal_data(m, 1)[variables].to_numpy().astype(dt)
                key = "%s_%s_%d_%d" % (dt, "".join(variables), N, m)
                for rule in ("normal_reference", "scott"):
                    h = oracle.diag_bandwidth(X, rule)
                    out["h_%s_%s" % (rule, key)] = h
                    logl, slogl = oracle.ref_product_kde_logl(X, T, h)
                    out["ref_product_logl_%s_%s" % (rule, key)] = logl
                    out["ref_product_slogl_%s_%s" % (rule, key)] = np.array(slogl)
                if dt == "float64":
                    out["scipy_product_logl_" + key] = scipy_product_logpdf(X, T)
    np.savez_compressed(os.path.join(HERE, "next_golden.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
