"""GPU parity tests for SURVEY §8 f3 through the C ABI: CKDE.cdf, CKDE.sample (device index selection),
LinearGaussianCPD.sample and BayesianNetwork.sample.  Shapes follow the reference's
tests/factors/continuous/CKDE_test.py:181-219, 406-553; the checkers are the CPU oracle (oracle/), the committed
golden vectors of the reference's own kernels (tests/golden/f3_golden.npz) and SciPy.

Tolerances: cdf values live in [0, 1], so the bar is absolute + relative: 1e-10 in float64 (north_star),
1e-4 in float32 (the reference's own test uses atol 5e-4).  Sampled training-row indices are integers: they must
equal the oracle's; sampled values then agree to rounding (same libstdc++ random streams on both sides)."""
import os

import numpy as np
import pandas as pd
import pyarrow as pa
import pytest

import oracle
import util_data

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "f3_golden.npz"))
VARSETS = [["a"], ["b", "a"], ["c", "a", "b"], ["d", "a", "b", "c"]]
CASES = [(500, 50), (40, 90), (300, 70)]
TOL = {"float64": 1e-10, "float32": 1e-4}


@pytest.fixture(scope="module")
def pbn():
    import pybnesian_b200 as pbn
    return pbn


def frames(variables, N, m, dt):
    df = util_data.generate_normal_data(N, 0).astype(dt)
    test = util_data.generate_normal_data(m, 1).astype(dt)
    return df, test, df[variables].to_numpy(), test[variables].to_numpy()


@pytest.mark.parametrize("dt", ["float64", "float32"])
@pytest.mark.parametrize("variables", VARSETS)
@pytest.mark.parametrize("N,m", CASES)
def test_cdf_vs_oracle_and_reference_golden(pbn, dt, variables, N, m):
    df, test, X, T = frames(variables, N, m, dt)
    cpd = pbn.CKDE(variables[0], variables[1:])
    cpd.fit(df)
    got = cpd.cdf(test)
    tol = TOL[dt]
    want = oracle.ckde_cdf(X, T, oracle.bandwidth(X))
    assert np.allclose(got, want, rtol=tol, atol=tol)
    key = "%s_%s_%d_%d" % (dt, "".join(variables), N, m)
    assert np.allclose(got, GOLD["ref_cdf_" + key], rtol=tol, atol=tol)
    if dt == "float64":
        assert np.allclose(got, GOLD["scipy_cdf_" + key], rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("dt", ["float64", "float32"])
def test_cdf_larger_and_split_over_ctas(pbn, dt):
    """10 000 training rows, 3 000 test rows: several test tiles x several training splits."""
    variables = ["d", "a", "b", "c"]
    df, test, X, T = frames(variables, 10000, 3000, dt)
    cpd = pbn.CKDE("d", ["a", "b", "c"])
    cpd.fit(df)
    got = cpd.cdf(test)
    sub = np.arange(0, 3000, 37)
    want = oracle.ckde_cdf(X, T[sub], oracle.bandwidth(X))
    assert np.allclose(got[sub], want, rtol=TOL[dt], atol=TOL[dt])
    assert np.all((got >= 0) & (got <= 1))


def test_cdf_nulls_and_evidence_order(pbn):
    df = util_data.generate_normal_data(500, 0)
    test = util_data.generate_normal_data(60, 1)
    np.random.seed(0)
    for col in "abcd":
        test.loc[np.random.randint(0, 60, size=5), col] = np.nan
    cpd = pbn.CKDE("d", ["a", "b", "c"])
    cpd.fit(df)
    cpd2 = pbn.CKDE("d", ["c", "b", "a"])
    cpd2.fit(df)
    got, got2 = cpd.cdf(test), cpd2.cdf(test)
    nan_rows = test[["a", "b", "c", "d"]].isna().any(axis=1).to_numpy()
    assert np.array_equal(np.isnan(got), nan_rows)
    assert np.allclose(got, got2, rtol=1e-9, atol=1e-12, equal_nan=True)
    X = df[["d", "a", "b", "c"]].to_numpy()
    T = test[["d", "a", "b", "c"]].to_numpy()[~nan_rows]
    assert np.allclose(got[~nan_rows], oracle.ckde_cdf(X, T, oracle.bandwidth(X)), rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize("dt", ["float64", "float32"])
def test_cdf_underflowing_weights_follow_the_reference(pbn, dt):
    """Evidence z bandwidths beyond the largest training value.  The reference's weights exp(-z^2/2 + c) are not
    max-shifted: once they all underflow in the data's dtype the row is 0/0 = NaN; before that (but below the fused
    kernel's unshifted-sum threshold) the row goes through the reference-arithmetic row kernel."""
    df, test, X, T = frames(["b", "a"], 300, 8, dt)
    H = oracle.bandwidth(X)
    h = np.sqrt(H[1, 1])
    z_fallback, z_nan = (37.0, 45.0) if dt == "float64" else (11.0, 16.0)
    test = test.astype("float64").copy()
    test.loc[0, "a"] = 1e4
    test.loc[1, "a"] = float(df["a"].max()) + z_fallback * h
    test.loc[2, "a"] = float(df["a"].max()) + z_nan * h
    test = test.astype(dt)
    cpd = pbn.CKDE("b", ["a"])
    cpd.fit(df)
    got = cpd.cdf(test)
    want = oracle.ckde_cdf(X, test[["b", "a"]].to_numpy(), H)
    assert np.isnan(want[0]) and np.isnan(want[2]) and np.isfinite(want[1])
    assert np.array_equal(np.isnan(got), np.isnan(want))
    ok = ~np.isnan(want)
    assert np.allclose(got[ok], want[ok], rtol=TOL[dt] * 100, atol=TOL[dt])
    from pybnesian_b200 import default_context
    import ctypes
    from pybnesian_b200 import _lib
    nf = ctypes.c_int64()
    _lib.check(_lib.lib().pbn_ctx_last_fallback_rows(default_context().handle, ctypes.byref(nf)))
    assert nf.value == 3


def test_cdf_wide_family_runtime_dimension(pbn):
    """11 variables: the runtime-dimension instantiation of the weight kernel."""
    df = util_data.iid_normal(400, 11, seed=0)
    df["x0"] = df["x0"] + 0.5 * df["x1"] - 0.25 * df["x7"]
    test = util_data.iid_normal(40, 11, seed=1)
    names = list(df.columns)
    cpd = pbn.CKDE(names[0], names[1:])
    cpd.fit(df)
    X, T = df[names].to_numpy(), test[names].to_numpy()
    want = oracle.ckde_cdf(X, T, oracle.bandwidth(X))
    assert np.allclose(cpd.cdf(test), want, rtol=1e-10, atol=1e-10)


def test_cdf_errors(pbn):
    df = util_data.generate_normal_data(100, 0)
    cpd = pbn.CKDE("b", ["a"])
    with pytest.raises(ValueError, match="not fitted"):
        cpd.cdf(df)
    cpd.fit(df)
    with pytest.raises(ValueError, match="Data type of training and test datasets is different"):
        cpd.cdf(df.astype("float32"))


# ---- sampling ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dt", ["float64", "float32"])
@pytest.mark.parametrize("variables", VARSETS)
@pytest.mark.parametrize("N,m", CASES)
def test_sample_vs_oracle_and_golden(pbn, dt, variables, N, m):
    df, test, X, T = frames(variables, N, m, dt)
    cpd = pbn.CKDE(variables[0], variables[1:])
    cpd.fit(df)
    ev = test[variables[1:]] if len(variables) > 1 else None
    arr, idx = cpd.sample(m, ev, 11, _return_indices=True)
    assert arr.type == (pa.float64() if dt == "float64" else pa.float32()) and len(arr) == m
    want, want_idx = oracle.ckde_sample(X, oracle.bandwidth(X), T[:, 1:] if len(variables) > 1 else None, m, 11)
    key = "%s_%s_%d_%d" % (dt, "".join(variables), N, m)
    if (N, m) == CASES[0]:
        assert np.array_equal(want_idx, GOLD["sample_idx_" + key])
    # float32: the device accumulates the running weight sum in double, the reference in float; an index may move
    # to a neighbouring boundary when u lands within float rounding of it
    mismatch = int(np.sum(idx != want_idx))
    assert mismatch == 0 if dt == "float64" else mismatch <= max(1, m // 25)
    same = idx == want_idx
    tol = 1e-12 if dt == "float64" else 1e-5
    assert np.allclose(arr.to_numpy()[same], want[same], rtol=tol, atol=tol * 10)


@pytest.mark.parametrize("dt", ["float64", "float32"])
def test_sample_indices_match_reference_kernels_golden(pbn, dt):
    """pbn_ckde_sample_indices against the indices the reference's own scan kernels select."""
    import ctypes
    from pybnesian_b200 import _lib
    from pybnesian_b200.dataset import DataFrame
    for variables in VARSETS[1:]:
        for N, m in CASES:
            df, test, X, T = frames(variables, N, m, dt)
            cpd = pbn.CKDE(variables[0], variables[1:])
            cpd.fit(df)
            u = oracle.uniform_real(m, 7, dt)
            tbl, cols, _ = DataFrame.wrap(test).device_table(variables[1:])
            out = np.empty(m, dtype=np.int32)
            _lib.check(_lib.lib().pbn_ckde_sample_indices(tbl.ctx.handle, cpd._handle.handle, tbl.handle, _lib.int_array(cols),
                                                          tbl.rows(), u.ctypes.data_as(ctypes.c_void_p),
                                                          out.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))))
            want = GOLD["ref_idx_%s_%s_%d_%d" % (dt, "".join(variables), N, m)]
            mismatch = int(np.sum(out != want))
            assert mismatch == 0 if dt == "float64" else mismatch <= max(1, m // 25), (variables, N, m, mismatch)


def test_sample_many_rows_early_exit_and_tail(pbn):
    """More samples than one CTA, training set spanning many tiles; u close to 1 must land on late rows and
    evidence far from the data falls back to the last training row like the reference."""
    df = util_data.generate_normal_data(5000, 0)
    cpd = pbn.CKDE("c", ["a", "b"])
    cpd.fit(df)
    n = 1500
    ev = util_data.generate_normal_data(n, 2)[["a", "b"]].copy()
    ev.loc[0, "a"] = 1e5
    arr, idx = cpd.sample(n, ev, 3, _return_indices=True)
    X = df[["c", "a", "b"]].to_numpy()
    want, want_idx = oracle.ckde_sample(X, oracle.bandwidth(X), ev.to_numpy(), n, 3)
    assert want_idx[0] == 4999 and idx[0] == 4999
    assert np.array_equal(idx, want_idx)
    assert np.allclose(arr.to_numpy(), want, rtol=1e-12, atol=1e-11)


def test_sample_api_contract(pbn):
    """CKDE_test.py:406-553."""
    df = util_data.generate_normal_data(1000, 0)
    cpd = pbn.CKDE("a", [])
    cpd.fit(df)
    s = cpd.sample(1000, None, 0)
    assert s.type == pa.float64() and len(s) == 1000
    cpd = pbn.CKDE("c", ["a", "b"])
    cpd.fit(df.astype("float32"))
    ev = pd.DataFrame({"a": np.full(1000, 3.0, dtype=np.float32), "b": np.full(1000, 7.45, dtype=np.float32)})
    s = cpd.sample(1000, ev, 0)
    assert s.type == pa.float32() and len(s) == 1000
    assert len(cpd.sample(0, ev, 0)) == 0
    with pytest.raises(ValueError, match="non-negative"):
        cpd.sample(-1, ev, 0)
    with pytest.raises(ValueError, match="Evidence values not present"):
        cpd.sample(10, ev[["a"]], 0)
    with pytest.raises(ValueError, match="different from CKDE training data"):
        cpd.sample(10, ev.astype("float64"), 0)
    with pytest.raises(ValueError, match="not fitted"):
        pbn.CKDE("a", []).sample(3, None, 0)
    # the same seed gives the same draw; no seed draws a fresh one
    assert np.array_equal(cpd.sample(50, ev, 9).to_numpy(), cpd.sample(50, ev, 9).to_numpy())


def test_bayesian_network_sample(pbn):
    """BNGeneric::sample: ancestral sampling, node i of the topological order uses seed + i."""
    df = util_data.generate_normal_data(2000, 0)
    model = pbn.SemiparametricBN(["a", "b", "c", "d"], [("a", "b"), ("a", "c"), ("b", "c"), ("c", "d")],
                                 [("b", pbn.CKDEType()), ("c", pbn.CKDEType())])
    model.fit(df)
    n, seed = 700, 42
    s = model.sample(n, seed, ordered=True)
    assert isinstance(s, pa.RecordBatch) and s.schema.names == ["a", "b", "c", "d"] and s.num_rows == n
    s = s.to_pandas()
    order = model.graph().topological_sort()
    assert order == ["a", "b", "c", "d"]
    # replay the chain with the oracle
    a = oracle.lg_sample(model.cpd("a").beta, model.cpd("a").variance, [], n, seed)
    assert np.array_equal(s["a"].to_numpy(), a)
    Xb = df[["b", "a"]].to_numpy()
    b, _ = oracle.ckde_sample(Xb, oracle.bandwidth(Xb), a.reshape(-1, 1), n, seed + 1)
    assert np.allclose(s["b"].to_numpy(), b, rtol=1e-12, atol=1e-11)
    Xc = df[["c"] + model.cpd("c").evidence()].to_numpy()
    evc = s[model.cpd("c").evidence()].to_numpy()
    c, _ = oracle.ckde_sample(Xc, oracle.bandwidth(Xc), evc, n, seed + 2)
    assert np.allclose(s["c"].to_numpy(), c, rtol=1e-12, atol=1e-10)
    d = oracle.lg_sample(model.cpd("d").beta, model.cpd("d").variance, [s["c"].to_numpy()], n, seed + 3)
    assert np.allclose(s["d"].to_numpy(), d, rtol=1e-12, atol=1e-10)
    # the sample follows the generating law: E[b] = 2.5 + 1.65 * 3
    assert abs(s["b"].mean() - (2.5 + 1.65 * 3)) < 0.3
    with pytest.raises(ValueError):
        model.sample(-1, 0)


@pytest.mark.parametrize("dt", ["float64", "float32"])
@pytest.mark.parametrize("d", [5, 6, 8])
def test_cdf_and_sample_medium_families(pbn, dt, d):
    """5-8 variables: the CDF mode of the pair kernel on top of the dot-product exponent form (float64, >= 4 evidence
    columns), the packed float32 tile, and the sampler's weight kernels at the same widths."""
    df = util_data.iid_normal(3000, d, seed=0).astype(dt)
    test = util_data.iid_normal(300, d, seed=1).astype(dt)
    names = list(df.columns)
    df["x0"] = (df["x0"] + 0.6 * df["x1"] - 0.4 * df[names[-1]]).astype(dt)
    test["x0"] = (test["x0"] + 0.6 * test["x1"] - 0.4 * test[names[-1]]).astype(dt)
    cpd = pbn.CKDE("x0", names[1:])
    cpd.fit(df)
    X, T = df[names].to_numpy(), test[names].to_numpy()
    H = oracle.bandwidth(X)
    tol = TOL[dt]
    assert np.allclose(cpd.cdf(test), oracle.ckde_cdf(X, T, H), rtol=tol, atol=tol)
    want_l, _ = oracle.ckde_logl(X, T, H)
    assert np.allclose(cpd.logl(test), want_l, rtol=tol, atol=tol)
    arr, idx = cpd.sample(300, test[names[1:]], 4, _return_indices=True)
    want, want_idx = oracle.ckde_sample(X, H, T[:, 1:], 300, 4)
    mismatch = int(np.sum(idx != want_idx))
    assert mismatch == 0 if dt == "float64" else mismatch <= 12
    same = idx == want_idx
    assert np.allclose(arr.to_numpy()[same], want[same], rtol=1e-12 if dt == "float64" else 1e-5, atol=1e-10 if dt == "float64" else 1e-4)
