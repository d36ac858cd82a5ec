"""CPU tests (no GPU) for the SURVEY §8(f) "next" rows of the oracle: ProductKDE (f2) against the committed golden
vectors of the reference's own kernels (tests/golden/next_golden.npz), against SciPy the way the reference's
ProductKDE_test.py does, and against oracle/_ref live when it is built."""
import os

import numpy as np
import pytest

import oracle
import util_data

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "next_golden.npz"))
VARSETS = [["a"], ["b", "a"], ["c", "a", "b"], ["d", "a", "b", "c"]]
CASES = [(500, 50), (150, 70)]


def data(variables, N, m, dt):
    X = util_data.generate_normal_data(N, 0)[variables].to_numpy().astype(dt)
    T = util_data.generate_normal_data(m, 1)[variables].to_numpy().astype(dt)
    return X, T


def py_nr_bandwidth(X):
    """tests/factors/continuous/ProductKDE_test.py:39-47."""
    cov = np.atleast_2d(np.cov(X, rowvar=False))
    delta = np.linalg.inv(np.diag(np.diag(cov))).dot(cov)
    delta_inv = np.linalg.inv(delta)
    N, d = X.shape
    k = 4 * d * np.sqrt(np.linalg.det(delta)) / (2 * (delta_inv.dot(delta_inv)).trace() + delta_inv.trace() ** 2)
    return np.power(k / N, 2 / (d + 4)) * np.diag(cov)


@pytest.mark.parametrize("dt", ["float64", "float32"])
@pytest.mark.parametrize("variables", VARSETS)
@pytest.mark.parametrize("N,m", CASES)
def test_product_kde_bit_exact_with_reference_kernels_golden(dt, variables, N, m):
    X, T = data(variables, N, m, dt)
    key = "%s_%s_%d_%d" % (dt, "".join(variables), N, m)
    for rule in ("normal_reference", "scott"):
        h = oracle.diag_bandwidth(X, rule)
        assert np.array_equal(h, GOLD["h_%s_%s" % (rule, key)])
        logl, slogl = oracle.product_kde_logl(X, T, h)
        assert np.array_equal(logl, GOLD["ref_product_logl_%s_%s" % (rule, key)])
        assert slogl == float(GOLD["ref_product_slogl_%s_%s" % (rule, key)])


@pytest.mark.parametrize("variables", VARSETS)
@pytest.mark.parametrize("N,m", CASES)
def test_product_kde_vs_scipy(variables, N, m):
    X, T = data(variables, N, m, "float64")
    key = "float64_%s_%d_%d" % ("".join(variables), N, m)
    h = oracle.diag_bandwidth(X)
    assert np.allclose(h, py_nr_bandwidth(X), rtol=1e-11, atol=0)
    assert np.allclose(oracle.diag_bandwidth(X, "scott"), N ** (-2 / (len(variables) + 4)) * np.var(X, axis=0, ddof=1),
                       rtol=1e-12, atol=0)
    logl, _ = oracle.product_kde_logl(X, T, h)
    assert np.allclose(logl, GOLD["scipy_product_logl_" + key], rtol=1e-10, atol=1e-12)


@pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref not built (needs /root/reference)")
def test_product_kde_vs_ref_kernels_live():
    for dt in ("float64", "float32"):
        X, T = data(["c", "a", "b"], 130, 66, dt)
        h = oracle.diag_bandwidth(X)
        a, sa = oracle.product_kde_logl(X, T, h)
        b, sb = oracle.ref_product_kde_logl(X, T, h)
        assert np.array_equal(a, b) and sa == sb


def test_product_kde_is_kde_with_diagonal_bandwidth():
    X, T = data(["d", "a", "b", "c"], 400, 30, "float64")
    h = oracle.diag_bandwidth(X)
    a, _ = oracle.product_kde_logl(X, T, h)
    b, _ = oracle.kde_logl(X, T, np.diag(h))
    assert np.allclose(a, b, rtol=1e-12, atol=0)
