"""CPU tests (no GPU) for SURVEY §8 f3: the oracle's CKDE.cdf / CKDE.sample restatement against the committed golden
vectors of the reference's own kernels (tests/golden/f3_golden.npz), against SciPy the way the reference's
CKDE_test.py:181-219,406-553 does, against oracle/_ref live when it is built, and the host pieces of the product
(pbn_lg_sample, LinearGaussianCPD.cdf, the Dag root set behind BayesianNetwork.sample's per-node seeds)."""
import ctypes
import os

import numpy as np
import pandas as pd
import pytest
from scipy.stats import norm

import oracle
import util_data

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "f3_golden.npz"))
VARSETS = [["a"], ["b", "a"], ["c", "a", "b"], ["d", "a", "b", "c"]]
CASES = [(500, 50), (40, 90), (300, 70)]


def data(variables, N, m, dt):
    X = util_data.generate_normal_data(N, 0)[variables].to_numpy().astype(dt)
    T = util_data.generate_normal_data(m, 1)[variables].to_numpy().astype(dt)
    return X, T


@pytest.mark.parametrize("dt", ["float64", "float32"])
@pytest.mark.parametrize("variables", VARSETS)
@pytest.mark.parametrize("N,m", CASES)
def test_cdf_vs_reference_kernels_golden(dt, variables, N, m):
    X, T = data(variables, N, m, dt)
    key = "%s_%s_%d_%d" % (dt, "".join(variables), N, m)
    got = oracle.ckde_cdf(X, T, oracle.bandwidth(X))
    want = GOLD["ref_cdf_" + key]
    # the kernels are restated op for op; the host-side transform / Cholesky (Eigen in the reference, numpy in the
    # golden script, plain loops in the oracle) agree to rounding only
    tol = 1e-13 if dt == "float64" else 2e-6
    assert np.allclose(got, want, rtol=tol, atol=tol)
    if dt == "float64":
        assert np.allclose(got, GOLD["scipy_cdf_" + key], rtol=1e-9, atol=1e-13)
    else:
        assert np.allclose(got, GOLD["scipy_cdf_float64_" + key[len("float32_"):]], atol=5e-4)  # the reference's own bound


@pytest.mark.parametrize("dt", ["float64", "float32"])
@pytest.mark.parametrize("variables", VARSETS[1:])
@pytest.mark.parametrize("N,m", CASES)
def test_sample_indices_vs_reference_kernels_golden(dt, variables, N, m):
    X, T = data(variables, N, m, dt)
    key = "%s_%s_%d_%d" % (dt, "".join(variables), N, m)
    H = oracle.bandwidth(X)
    u = oracle.uniform_real(m, 7, dt)
    got = oracle.ckde_sample_indices(X[:, 1:], T[:, 1:], H[1:, 1:], u)
    assert np.array_equal(got, GOLD["ref_idx_" + key])


@pytest.mark.parametrize("dt", ["float64", "float32"])
@pytest.mark.parametrize("variables", VARSETS)
def test_sample_stream_golden(dt, variables):
    N, m = CASES[0]
    X, T = data(variables, N, m, dt)
    key = "%s_%s_%d_%d" % (dt, "".join(variables), N, m)
    smp, idx = oracle.ckde_sample(X, oracle.bandwidth(X), T[:, 1:] if len(variables) > 1 else None, m, 11)
    assert smp.dtype == np.dtype(dt)
    assert np.array_equal(idx, GOLD["sample_idx_" + key])
    assert np.array_equal(smp, GOLD["sample_" + key])
    assert np.array_equal(oracle.uniform_real(16, 7, dt), GOLD["u_" + dt])


def test_libstdcxx_uniform_real_known_answers():
    """std::mt19937{0}: first outputs 2357136044, 2546248239 -> generate_canonical<double, 53> uses two draws."""
    u = oracle.uniform_real(2, 0, np.float64)
    r = 4294967296.0
    assert u[0] == (2357136044.0 + 2546248239.0 * r) / (r * r)
    f = oracle.uniform_real(1, 0, np.float32)
    assert f[0] == np.float32(2357136044.0 / r)


@pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref not built (needs /root/reference)")
def test_cdf_and_indices_vs_ref_kernels_live():
    for dt in ("float64", "float32"):
        X, T = data(["d", "a", "b", "c"], 270, 33, dt)
        H = oracle.bandwidth(X)
        tol = 1e-13 if dt == "float64" else 2e-6
        assert np.allclose(oracle.ckde_cdf(X, T, H), oracle.ref_ckde_cdf(X, T, H), rtol=tol, atol=tol)
        u = oracle.uniform_real(33, 1, dt)
        assert np.array_equal(oracle.ckde_sample_indices(X[:, 1:], T[:, 1:], H[1:, 1:], u),
                              oracle.ref_ckde_sample_indices(X[:, 1:], T[:, 1:], H[1:, 1:], u))


def test_cdf_underflow_is_nan_like_the_reference():
    X, T = data(["b", "a"], 200, 4, "float64")
    T[0, 1] = 1e4  # evidence far from every training point: all weights underflow, 0/0
    out = oracle.ckde_cdf(X, T, oracle.bandwidth(X))
    assert np.isnan(out[0]) and np.all(np.isfinite(out[1:]))


def test_sample_distribution_sanity():
    X, _ = data(["b", "a"], 2000, 1, "float64")
    H = oracle.bandwidth(X)
    ev = np.full((4000, 1), 3.0)
    smp, idx = oracle.ckde_sample(X, H, ev, 4000, 0)
    near = np.abs(X[idx, 1] - 3.0)
    assert np.mean(near) < 3 * np.sqrt(H[1, 1])
    # E[b | a = 3] = 2.5 + 1.65 * 3
    assert abs(np.mean(smp) - (2.5 + 1.65 * 3.0)) < 0.2


# ---- host pieces of the product (no GPU needed) ----------------------------------------------------------
def test_pbn_lg_sample_bit_exact():
    from pybnesian_b200 import _lib
    L = _lib.lib()
    beta = np.array([1.5, -0.7, 2.25])
    ev = util_data.generate_normal_data(64, 3)
    dp = ctypes.POINTER(ctypes.c_double)
    for name, cols, code in (("lg_sample_f64", [ev["a"].to_numpy(), ev["b"].to_numpy()], 0),
                             ("lg_sample_f32", [ev["a"].to_numpy().astype(np.float32), ev["b"].to_numpy().astype(np.float32)], 1),
                             ("lg_sample_noev", [], 0)):
        cols = [np.ascontiguousarray(c) for c in cols]
        ptrs = (ctypes.c_void_p * max(1, len(cols)))(*[c.ctypes.data for c in cols])
        out = np.empty(64)
        b = np.ascontiguousarray(beta[:len(cols) + 1])
        assert L.pbn_lg_sample(b.ctypes.data_as(dp), 0.81, len(cols), ptrs, code, 64, 5, out.ctypes.data_as(dp)) == 0
        assert np.array_equal(out, GOLD[name])
    assert L.pbn_lg_sample(beta.ctypes.data_as(dp), 0.81, 0, None, 0, -1, 5, None) == _lib.PBN_ERR_ARG


def test_linear_gaussian_sample_and_cdf_host():
    import pyarrow as pa
    import pybnesian_b200 as pbn
    ev = util_data.generate_normal_data(64, 3)
    cpd = pbn.LinearGaussianCPD("c", ["a", "b"], [1.5, -0.7, 2.25], 0.81)
    s = cpd.sample(64, ev, 5)
    assert s.type == pa.float64() and np.array_equal(s.to_numpy(), GOLD["lg_sample_f64"])
    assert np.array_equal(pbn.LinearGaussianCPD("c", [], [1.5], 0.81).sample(64, None, 5).to_numpy(), GOLD["lg_sample_noev"])
    with pytest.raises(ValueError):
        cpd.sample(-1, ev, 0)
    with pytest.raises(ValueError):
        cpd.sample(10, ev[["a"]], 0)
    # LinearGaussianCPD_test.py:155-187: cdf against scipy.stats.norm
    want = norm.cdf(ev["c"], 1.5 - 0.7 * ev["a"] + 2.25 * ev["b"], np.sqrt(0.81))
    assert np.allclose(cpd.cdf(ev), want, rtol=1e-12, atol=1e-15)
    assert np.allclose(cpd.cdf(ev.astype("float32")), want, atol=5e-4)
    evn = ev.copy()
    evn.loc[3, "a"] = np.nan
    got = cpd.cdf(evn)
    assert np.isnan(got[3]) and np.allclose(np.delete(got, 3), np.delete(want, 3), rtol=1e-12)


def test_dag_roots_follow_unordered_set_history():
    """DagImpl::topological_sort seeds its stack with the iteration order of ArcGraph::m_roots
    (std::unordered_set<int>): libstdc++ lists small int sets in reverse insertion order."""
    from pybnesian_b200.models import Dag
    g = Dag(["a", "b", "c", "d"])
    assert g._roots.list() == [3, 2, 1, 0]
    assert g.topological_sort() == ["a", "b", "c", "d"]  # the stack pops from the back
    g.add_arc("a", "b")
    g.add_arc("b", "c")
    assert sorted(g._roots.list()) == [0, 3]
    order = g.topological_sort()
    assert order.index("a") < order.index("b") < order.index("c")
    g.remove_arc("a", "b")  # b becomes a root again: re-inserted at the front of the bucket list
    assert g._roots.list()[0] == 1
    h = g.clone()
    assert h._roots.list() == g._roots.list() and h.topological_sort() == g.topological_sort()
    g.flip_arc("b", "c")
    assert 2 in g._roots and 1 not in g._roots
