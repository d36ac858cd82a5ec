"""GPU parity tests for KDE / CKDE logl & slogl through the C ABI (via the Python mirror of
the reference API).  Shapes follow the reference's tests/factors/continuous/KDE_test.py and
CKDE_test.py; the checker is the CPU oracle (oracle/), with SciPy as an independent check.

Tolerances (BASELINE.json north_star): 1e-10 relative in float64, 1e-4 in float32.
"""
import numpy as np
import pandas as pd
import pytest
from scipy.stats import gaussian_kde

import oracle
import util_data

pytestmark = pytest.mark.gpu

RTOL64 = 1e-10
RTOL32 = 1e-4

VARSETS = [["a"], ["b", "a"], ["c", "a", "b"], ["d", "a", "b", "c"]]


@pytest.fixture(scope="module")
def pbn():
    import pybnesian_b200 as pbn
    return pbn


def relerr(got, want, floor=1e-300):
    return np.max(np.abs(got - want) / np.maximum(np.abs(want), floor))


def relerr32(got, want):
    """float32 bar of the north star: 1e-4 RELATIVE, no floor (round 1 measured against max(|logl|, 1))."""
    return relerr(got, want)


def ckde_err32(got, want, X, T, H):
    """float32 bar for a CKDE log-likelihood = joint - marginal: both terms are held to 1e-4 relative (the callers
    check kde_joint / kde_marg that way), so the difference - which crosses zero - can only be held to
    1e-4 (|joint| + |marginal|).  Returns max |got - want| / (|joint| + |marginal|) with the terms from the oracle."""
    joint, _ = oracle.kde_logl(X, T, H)
    scale = np.abs(joint)
    if X.shape[1] > 1:
        marg, _ = oracle.kde_logl(np.ascontiguousarray(X[:, 1:]), np.ascontiguousarray(T[:, 1:]), H[1:, 1:])
        scale = scale + np.abs(marg)
    return np.max(np.abs(got - want) / scale)


@pytest.mark.parametrize("variables", VARSETS)
@pytest.mark.parametrize("n_train", [10, 500, 10000])
def test_kde_logl_f64_vs_oracle(pbn, variables, n_train):
    df = util_data.generate_normal_data(n_train, seed=0)
    test = util_data.generate_normal_data(50, seed=1)
    if n_train <= len(variables):
        pytest.skip("not enough instances")
    k = pbn.KDE(variables)
    k.fit(df)
    X, T = df[variables].to_numpy(), test[variables].to_numpy()
    H = oracle.bandwidth(X)
    assert relerr(k.bandwidth, H) < 1e-12
    want, want_s = oracle.kde_logl(X, T, H)
    got = k.logl(test)
    assert relerr(got, want) < RTOL64
    assert abs(k.slogl(test) - want_s) <= RTOL64 * abs(want_s)
    sk = gaussian_kde(X.T, bw_method=lambda s: np.power(4 / (s.d + 2), 1 / (s.d + 4)) * s.scotts_factor())
    assert np.allclose(got, sk.logpdf(T.T), rtol=1e-9, atol=0)


@pytest.mark.parametrize("variables", VARSETS)
@pytest.mark.parametrize("n_train", [500, 10000])
def test_kde_logl_f32_vs_oracle(pbn, variables, n_train):
    df = util_data.generate_normal_data(n_train, seed=0).astype("float32")
    test = util_data.generate_normal_data(50, seed=1).astype("float32")
    k = pbn.KDE(variables)
    k.fit(df)
    X, T = df[variables].to_numpy(), test[variables].to_numpy()
    H = oracle.bandwidth(X)
    assert relerr(k.bandwidth, H) < 1e-5
    want, want_s = oracle.kde_logl(X, T, H)
    got = k.logl(test)
    assert relerr32(got, want) < RTOL32
    assert abs(k.slogl(test) - want_s) <= RTOL32 * abs(want_s)


@pytest.mark.parametrize("rule", ["normal_reference", "scott"])
def test_bandwidth_rules(pbn, rule):
    df = util_data.generate_normal_data(1000, seed=0)
    sel = pbn.NormalReferenceRule() if rule == "normal_reference" else pbn.ScottsBandwidth()
    for variables in VARSETS:
        H = sel.bandwidth(df, variables)
        assert relerr(H, oracle.bandwidth(df[variables].to_numpy(), rule)) < 1e-12


@pytest.mark.parametrize("evidence", [[], ["a"], ["a", "b"], ["a", "b", "c"]])
@pytest.mark.parametrize("n_train", [10, 10000])
@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_ckde_logl_vs_oracle(pbn, evidence, n_train, dtype):
    variables = ["d"] + evidence
    if n_train <= len(variables):
        pytest.skip("not enough instances")
    df = util_data.generate_normal_data(n_train, seed=0).astype(dtype)
    test = util_data.generate_normal_data(50, seed=1).astype(dtype)
    cpd = pbn.CKDE("d", evidence)
    cpd.fit(df)
    X, T = df[variables].to_numpy(), test[variables].to_numpy()
    H = oracle.bandwidth(X)
    want, want_s = oracle.ckde_logl(X, T, H)
    got = cpd.logl(test)
    tol = RTOL64 if dtype == "float64" else RTOL32
    err = relerr(got, want) if dtype == "float64" else ckde_err32(got, want, X, T, H)
    assert err < tol
    assert abs(cpd.slogl(test) - want_s) <= tol * abs(want_s)
    if dtype == "float32":  # the two terms of the difference, strictly relative
        wj, _ = oracle.kde_logl(X, T, H)
        assert relerr32(cpd.kde_joint().logl(test), wj) < tol
        if evidence:
            wm, _ = oracle.kde_logl(np.ascontiguousarray(X[:, 1:]), np.ascontiguousarray(T[:, 1:]), H[1:, 1:])
            assert relerr32(cpd.kde_marg().logl(test), wm) < tol
    if evidence:
        assert relerr(cpd.kde_marg().bandwidth, H[1:, 1:]) < (1e-12 if dtype == "float64" else 1e-5)


def test_nulls_give_nan_in_place(pbn):
    df = util_data.generate_normal_data(2000, seed=0)
    test = util_data.generate_normal_data(300, seed=1)
    rng = np.random.default_rng(3)
    test_null = test.copy()
    for col in ["a", "b"]:
        test_null.loc[rng.choice(300, 40, replace=False), col] = np.nan
    variables = ["c", "a", "b"]
    k = pbn.KDE(variables)
    k.fit(df)
    got = k.logl(test_null)
    isnull = test_null[variables].isna().any(axis=1).to_numpy()
    assert np.all(np.isnan(got[isnull])) and not np.any(np.isnan(got[~isnull]))
    full = k.logl(test)
    assert relerr(got[~isnull], full[~isnull]) < 1e-13
    assert abs(k.slogl(test_null) - np.nansum(got)) < 1e-9
    # training data with nulls: rows dropped (KDE::_fit contains_null branch)
    df_null = df.copy()
    df_null.loc[rng.choice(2000, 100, replace=False), "a"] = np.nan
    k2 = pbn.KDE(variables)
    k2.fit(df_null)
    keep = ~df_null[variables].isna().any(axis=1).to_numpy()
    assert k2.num_instances() == keep.sum()
    X = df_null[variables].to_numpy()[keep]
    want, _ = oracle.kde_logl(X, test[variables].to_numpy(), oracle.bandwidth(X))
    assert relerr(k2.logl(test), want) < RTOL64


def test_type_mismatch_and_errors(pbn):
    df = util_data.generate_normal_data(500, seed=0)
    dff = df.astype("float32")
    k = pbn.KDE(["a"])
    with pytest.raises(ValueError, match="KDE factor not fitted"):
        k.logl(df)
    k.fit(df)
    for fn in (k.logl, k.slogl):
        with pytest.raises(ValueError, match="Data type of training and test datasets is different."):
            fn(dff)
    with pytest.raises(ValueError, match="Cannot create a KDE model with 0 variables"):
        pbn.KDE([])
    with pytest.raises(pbn.SingularCovarianceData):
        pbn.KDE(["a", "b", "c"]).fit(df.iloc[:3])
    dup = df.copy()
    dup["b"] = 2 * dup["a"]
    with pytest.raises(pbn.SingularCovarianceData, match="not positive-definite"):
        pbn.KDE(["a", "b"]).fit(dup)


def test_variable_order_invariance_and_set_bandwidth(pbn):
    df = util_data.generate_normal_data(3000, seed=0)
    test = util_data.generate_normal_data(200, seed=1)
    k1 = pbn.KDE(["a", "b", "c"]); k1.fit(df)
    k2 = pbn.KDE(["c", "a", "b"]); k2.fit(df)
    assert relerr(k1.logl(test), k2.logl(test)) < 1e-11
    k = pbn.KDE(["a"]); k.fit(df)
    k.bandwidth = [[1.0]]
    want, _ = oracle.kde_logl(df[["a"]].to_numpy(), test[["a"]].to_numpy(), np.array([[1.0]]))
    assert relerr(k.logl(test), want) < RTOL64


def test_python_bandwidth_selector(pbn):
    class Halved(pbn.BandwidthSelector):
        def bandwidth(self, df, variables):
            return 0.5 * pbn.NormalReferenceRule().bandwidth(df, variables)

    df = util_data.generate_normal_data(2000, seed=0)
    test = util_data.generate_normal_data(100, seed=1)
    k = pbn.KDE(["a", "b"], Halved()); k.fit(df)
    X = df[["a", "b"]].to_numpy()
    want, _ = oracle.kde_logl(X, test[["a", "b"]].to_numpy(), 0.5 * oracle.bandwidth(X))
    assert relerr(k.logl(test), want) < RTOL64


def test_far_test_points_use_shifted_path(pbn):
    """Test rows tens of bandwidths away from every training point: the unshifted kernel sum
    underflows and the rows are re-evaluated with a max shift (what the reference's
    logsumexp_cols_offset does for every row)."""
    df = util_data.generate_normal_data(4000, seed=0)
    test = util_data.generate_normal_data(64, seed=1)
    test.loc[:7, "a"] += 40.0
    test.loc[8:11, "b"] -= 300.0
    for dtype, tol in (("float64", RTOL64), ("float32", RTOL32)):
        tr, te = df.astype(dtype), test.astype(dtype)
        for variables in (["a"], ["a", "b"], ["c", "a", "b"]):
            k = pbn.KDE(variables); k.fit(tr)
            X = tr[variables].to_numpy()
            want, _ = oracle.kde_logl(X, te[variables].to_numpy(), oracle.bandwidth(X))
            got = k.logl(te)
            assert np.all(np.isfinite(got))
            assert (relerr(got, want) if dtype == "float64" else relerr32(got, want)) < tol
            assert pbn.default_context().last_fallback_rows() > 0
        cpd = pbn.CKDE("b", ["a", "c"]); cpd.fit(tr)
        X = tr[["b", "a", "c"]].to_numpy()
        Tb = te[["b", "a", "c"]].to_numpy()
        want, _ = oracle.ckde_logl(X, Tb, oracle.bandwidth(X))
        got = cpd.logl(te)
        assert (relerr(got, want) if dtype == "float64" else ckde_err32(got, want, X, Tb, oracle.bandwidth(X))) < tol


@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_every_test_row_far_from_the_data(pbn, dtype):
    """A shifted test set: EVERY row is tens of bandwidths from every training row, so every unshifted sum underflows.
    Round 1 sent such rows to one CTA per row (>= 10x slower); now they take a second pass of the tiled pair kernel
    with a per-row exponent shift (runtime.cu: compact_flagged / rowmin / pair_kernel<SHIFT> / finalize_shift), and
    none of them needs the per-row kernel.  Values against the oracle on a sub-sample, KDE and CKDE."""
    n, m = 100_000, 20_000
    tr = util_data.generate_normal_data(n, seed=0).astype(dtype)
    te = util_data.generate_normal_data(m, seed=1)
    te["a"] += 6.0          # 12 standard deviations of a: > 50 bandwidths
    te["d"] -= 40.0
    te = te.astype(dtype)
    tol = RTOL64 if dtype == "float64" else RTOL32
    ctx = pbn.default_context()
    rows = np.random.default_rng(0).choice(m, 256, replace=False)

    def oracle_logl(fn, X, T, H):
        # float32: at |logl| ~ 4e4 with a strongly correlated bandwidth the reference's float arithmetic (forward
        # substitution on raw float differences) is itself only good to ~5e-4 relative, so the float32 rows are held to
        # the 1e-4 bar against the float64 evaluation of the same (float-valued) data and bandwidth
        return fn(X.astype(np.float64), T.astype(np.float64), np.asarray(H, dtype=np.float64))[0]

    for variables in (["a"], ["b", "a"], ["d", "a", "b", "c"]):
        k = pbn.KDE(variables); k.fit(tr)
        got = k.logl(te)
        assert ctx.last_fallback_rows() >= 0.99 * m and ctx.last_row_kernel_rows() == 0
        X, T = tr[variables].to_numpy(), te[variables].to_numpy()[rows]
        want = oracle_logl(oracle.kde_logl, X, T, k.bandwidth)
        assert np.all(np.isfinite(got)) and relerr(got[rows], want) < tol
        total = k.slogl(te)
        assert abs(total - got.sum()) <= 1e-12 * abs(total)
    cpd = pbn.CKDE("d", ["a", "b", "c"]); cpd.fit(tr)
    got = cpd.logl(te)
    assert ctx.last_fallback_rows() >= 0.99 * m and ctx.last_row_kernel_rows() == 0
    V = ["d", "a", "b", "c"]
    X, T = tr[V].to_numpy(), te[V].to_numpy()[rows]
    H = np.asarray(cpd.kde_joint().bandwidth)
    want = oracle_logl(oracle.ckde_logl, X, T, H)
    assert relerr(got[rows], want) < tol
    # a mixed set: near rows keep the one-pass result, far rows take the second pass, in any interleaving
    mixed = pd.concat([util_data.generate_normal_data(3000, seed=3).astype(dtype), te.iloc[:1500]], ignore_index=True)
    mixed = mixed.iloc[np.random.default_rng(1).permutation(len(mixed))].reset_index(drop=True)
    got = cpd.logl(mixed)
    assert ctx.last_fallback_rows() >= 1500 and ctx.last_row_kernel_rows() == 0
    Tm = mixed[V].to_numpy()[:400]
    want = oracle_logl(oracle.ckde_logl, X, Tm, H)
    X64, T64 = X.astype(np.float64), Tm.astype(np.float64)
    assert (relerr(got[:400], want) if dtype == "float64" else ckde_err32(got[:400], want, X64, T64, H)) < tol
    # beyond the reach of the integer shift (2^31 kernel units ~ 850 bandwidths in f64): the per-row kernel still answers
    huge = te.iloc[:64].copy()
    huge["a"] += np.asarray(1.0e4, dtype=dtype)
    k = pbn.KDE(["a"]); k.fit(tr)
    got = k.logl(huge)
    X = tr[["a"]].to_numpy()
    want = oracle_logl(oracle.kde_logl, X, huge[["a"]].to_numpy(), k.bandwidth)
    assert np.all(np.isfinite(got)) and relerr(got, want) < tol


@pytest.mark.parametrize("d", [5, 8, 9, 10, 11, 12])
def test_higher_dimensions(pbn, d):
    """d = 9, 10 run the fused pair kernel (pair_kernel.cuh: kMaxFastD = 10, two rows per thread in f64), d > 10 the
    generic row kernel; the reference runs any d through the same kernels (KDE.cl.src:123-135)."""
    tr = util_data.iid_normal(3000, d, seed=0)
    te = util_data.iid_normal(100, d, seed=1)
    variables = list(tr.columns)
    k = pbn.KDE(variables); k.fit(tr)
    X = tr.to_numpy()
    want, _ = oracle.kde_logl(X, te.to_numpy(), oracle.bandwidth(X))
    assert relerr(k.logl(te), want) < RTOL64
    cpd = pbn.CKDE(variables[0], variables[1:]); cpd.fit(tr)
    want, _ = oracle.ckde_logl(X, te.to_numpy(), oracle.bandwidth(X))
    assert relerr(cpd.logl(te), want) < 1e-9


@pytest.mark.parametrize("d", [9, 10])
def test_wide_fast_path_f32_cdf_and_ragged(pbn, d):
    """The d = 9, 10 instantiations: float32, the CDF mode, and sizes that are not multiples of the tiles."""
    tr = util_data.iid_normal(20_011, d, seed=0)
    te = util_data.iid_normal(1_537, d, seed=1)
    variables = list(tr.columns)
    X, T = tr.to_numpy(), te.to_numpy()
    H = oracle.bandwidth(X)
    k = pbn.KDE(variables); k.fit(tr)
    want, want_s = oracle.kde_logl(X, T, H)
    assert relerr(k.logl(te), want) < RTOL64
    assert abs(k.slogl(te) - want_s) <= RTOL64 * abs(want_s)
    assert pbn.default_context().last_fallback_rows() == 0
    cpd = pbn.CKDE(variables[0], variables[1:]); cpd.fit(tr)
    wantc, _ = oracle.ckde_logl(X, T, H)
    assert np.all(np.abs(cpd.logl(te) - wantc) <= 1e-12 + 1e-10 * np.abs(wantc))
    wantcdf = oracle.ckde_cdf(X, T[:300], H)
    assert np.allclose(cpd.cdf(te.iloc[:300]), wantcdf, rtol=1e-10, atol=1e-12)
    tr32, te32 = tr.astype("float32"), te.astype("float32")
    X32, T32 = tr32.to_numpy(), te32.to_numpy()
    k32 = pbn.KDE(variables); k32.fit(tr32)
    want32, _ = oracle.kde_logl(X32, T32, oracle.bandwidth(X32))
    assert np.max(np.abs(k32.logl(te32) - want32) / np.abs(want32)) < RTOL32
    c32 = pbn.CKDE(variables[0], variables[1:]); c32.fit(tr32)
    want32c, _ = oracle.ckde_logl(X32, T32, oracle.bandwidth(X32))
    assert ckde_err32(c32.logl(te32), want32c, X32, T32, oracle.bandwidth(X32)) < RTOL32


def test_config1_shape(pbn):
    """BASELINE.json configs[0]: KDE(['a','b']) fit + logl, 10k train / 10k test, float64."""
    tr = util_data.generate_normal_data(10000, seed=0)
    te = util_data.generate_normal_data(10000, seed=1)
    k = pbn.KDE(["a", "b"]); k.fit(tr)
    X = tr[["a", "b"]].to_numpy()
    want, want_s = oracle.kde_logl(X, te[["a", "b"]].to_numpy(), oracle.bandwidth(X))
    assert relerr(k.logl(te), want) < RTOL64
    assert abs(k.slogl(te) - want_s) < RTOL64 * abs(want_s)


def test_multi_tile_ragged_sizes(pbn):
    """Sizes that are not multiples of the train / test tiles and need several CTAs per test tile."""
    for n, m in [(513, 1), (1025, 511), (5000, 1537), (20011, 2049)]:
        tr = util_data.generate_normal_data(n, seed=2)
        te = util_data.generate_normal_data(m, seed=3)
        for dtype, tol in (("float64", RTOL64), ("float32", RTOL32)):
            cpd = pbn.CKDE("c", ["a", "b"]); cpd.fit(tr.astype(dtype))
            X = tr[["c", "a", "b"]].to_numpy().astype(dtype)
            T = te[["c", "a", "b"]].to_numpy().astype(dtype)
            sub = slice(0, min(m, 200))
            want, _ = oracle.ckde_logl(X, T[sub], oracle.bandwidth(X))
            got = cpd.logl(te.astype(dtype))[sub]
            assert (relerr(got, want) if dtype == "float64" else ckde_err32(got, want, X, T[sub], oracle.bandwidth(X))) < tol


GOLD = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "kde_golden.npz"))


@pytest.mark.parametrize("dt", ["float64", "float32"])
@pytest.mark.parametrize("variables", VARSETS)
@pytest.mark.parametrize("N,m", [(500, 50), (10, 50), (300, 70)])
def test_against_reference_kernel_goldens(pbn, dt, variables, N, m):
    """CUDA path vs the committed outputs of the reference's own OpenCL-C kernels
    (tests/golden/make_golden.py), independent of the oracle library."""
    if N <= len(variables):
        pytest.skip("not enough instances")
    tr = util_data.generate_normal_data(N, 0).astype(dt)
    te = util_data.generate_normal_data(m, 1).astype(dt)
    key = "%s_%s_%d_%d" % (dt, "".join(variables), N, m)
    err = relerr if dt == "float64" else relerr32
    tol = RTOL64 if dt == "float64" else RTOL32
    k = pbn.KDE(variables); k.fit(tr)
    assert relerr(k.bandwidth, GOLD["H_" + key]) < (1e-12 if dt == "float64" else 1e-5)
    assert err(k.logl(te), GOLD["ref_kde_logl_" + key]) < tol
    assert abs(k.slogl(te) - float(GOLD["ref_kde_slogl_" + key])) <= tol * abs(float(GOLD["ref_kde_slogl_" + key]))
    cpd = pbn.CKDE(variables[0], variables[1:]); cpd.fit(tr)
    if dt == "float64":
        assert err(cpd.logl(te), GOLD["ref_ckde_logl_" + key]) < tol
    else:
        Xg, Tg = tr[variables].to_numpy(), te[variables].to_numpy()
        assert ckde_err32(cpd.logl(te), GOLD["ref_ckde_logl_" + key], Xg, Tg, oracle.bandwidth(Xg)) < tol
    if dt == "float64":
        assert np.allclose(k.logl(te), GOLD["scipy_kde_logl_" + key], rtol=1e-9, atol=1e-11)


@pytest.mark.parametrize("d", [4, 8])
def test_dot_product_form_on_heavy_tailed_data(pbn, d):
    """Student-t(4) columns put the bounding box of the whitened rows at 25-50 bandwidths: inside the window where the
    f64 kernel takes the dot-product form of the exponent (pair_kernel.cuh: tile_f64_dot, worst-case cancellation bound
    1e-11 per term).  The rows that matter are the outlying ones; every row must still meet the 1e-10 bar."""
    rng = np.random.default_rng(5)
    n, m = 100_000, 1500
    train = pd.DataFrame({"x%d" % i: rng.standard_t(4, n) for i in range(d)})
    test = pd.DataFrame({"x%d" % i: rng.standard_t(4, m) for i in range(d)})
    # the most outlying training points are test rows too (their own kernel keeps the sum finite)
    far = np.argsort(-np.abs(train.to_numpy()).max(axis=1))[:100]
    test = pd.concat([test, train.iloc[far]], ignore_index=True)
    k = pbn.KDE(list(train.columns))
    k.fit(train)
    got = k.logl(test)
    X, T = train.to_numpy(), test.to_numpy()
    want, want_s = oracle.kde_logl(X, T, oracle.bandwidth(X))
    assert np.all(np.isfinite(want))
    assert relerr(got, want) < RTOL64
    assert abs(k.slogl(test) - want_s) <= RTOL64 * abs(want_s)


def test_edge_cases_empty_small_and_wide(pbn):
    """Empty test sets, all-null test sets, too few instances, the widest supported family."""
    df = util_data.generate_normal_data(300, 0)
    k = pbn.KDE(["a", "b"])
    k.fit(df)
    empty = df.iloc[:0]
    assert k.logl(empty).shape == (0,) and k.slogl(empty) == 0.0
    cpd = pbn.CKDE("c", ["a", "b"])
    cpd.fit(df)
    assert cpd.logl(empty).shape == (0,) and cpd.slogl(empty) == 0.0 and cpd.cdf(empty).shape == (0,)
    allnull = df.iloc[:5].copy()
    allnull["a"] = np.nan
    assert np.all(np.isnan(k.logl(allnull))) and k.slogl(allnull) == 0.0
    assert np.all(np.isnan(cpd.cdf(allnull)))
    # valid_rows <= d: SingularCovarianceData (NormalReferenceRule.hpp:37-60), a ValueError
    with pytest.raises(pbn.SingularCovarianceData):
        pbn.KDE(["a", "b"]).fit(df.iloc[:2])
    with pytest.raises(ValueError):
        pbn.CKDE("c", ["a", "b"]).fit(df.iloc[:3])
    # a training set evaluated on itself (the zero-distance pair contributes exp(0) = 1 exactly)
    X = df[["a", "b"]].to_numpy()
    want, _ = oracle.kde_logl(X, X, oracle.bandwidth(X))
    assert relerr(k.logl(df), want) < RTOL64
    # the widest family the C ABI takes (PBN_MAX_DIM = 32 variables) and one more
    wide = util_data.iid_normal(400, 33, seed=2)
    names = list(wide.columns)
    k32 = pbn.KDE(names[:32])
    k32.fit(wide)
    W = wide[names[:32]].to_numpy()
    want, _ = oracle.kde_logl(W, W[:20], oracle.bandwidth(W))
    assert relerr(k32.logl(wide.iloc[:20]), want) < RTOL64
    c32 = pbn.CKDE(names[0], names[1:32])
    c32.fit(wide)
    wantc, _ = oracle.ckde_logl(W, W[:20], oracle.bandwidth(W))
    assert np.allclose(c32.logl(wide.iloc[:20]), wantc, rtol=1e-9, atol=1e-9)
    wantcdf = oracle.ckde_cdf(W, W[:20], oracle.bandwidth(W))
    assert np.allclose(c32.cdf(wide.iloc[:20]), wantcdf, rtol=1e-9, atol=1e-10)
    with pytest.raises(ValueError):
        pbn.KDE(names).fit(wide)


def test_non_finite_test_values_propagate(pbn):
    df = util_data.generate_normal_data(300, 0)
    k = pbn.KDE(["a", "b"])
    k.fit(df)
    t = df.iloc[:4].copy()
    t.loc[t.index[1], "a"] = np.inf
    t.loc[t.index[2], "b"] = -np.inf
    got = k.logl(t)
    want, _ = oracle.kde_logl(df[["a", "b"]].to_numpy(), t[["a", "b"]].to_numpy(), oracle.bandwidth(df[["a", "b"]].to_numpy()))
    # the reference's arithmetic turns an infinite coordinate into exp(-inf) terms: logl = -inf (or NaN from inf - inf)
    assert np.isfinite(got[0]) and np.isfinite(got[3])
    for i in (1, 2):
        assert not np.isfinite(got[i]) and not np.isfinite(want[i])


@pytest.mark.parametrize("order", ["near_first", "far_first", "shuffled"])
@pytest.mark.parametrize("evidence", [[], ["a"], ["a", "c", "d"]])
def test_exponent_floor_is_invisible(pbn, order, evidence):
    """The f64 kernel clamps the rounded exponent of a term to a floor derived from the row's running sum
    (pair_kernel.cuh: pair_floor), so that far pairs read one table entry.  Two clusters 25 sigma apart make most terms of
    every row fall below any such floor; the training order decides whether the floors are built from the near cluster
    (large sums early), from the far one (sums of ~1e-200 for many tiles, then terms 200 orders larger) or mixed.  The
    result must not depend on it and must meet the 1e-10 bar against the oracle in every order."""
    rng = np.random.default_rng(11)
    n_half = 30_000
    near = util_data.generate_normal_data(n_half, seed=0)
    far = util_data.generate_normal_data(n_half, seed=2) + 25.0
    train = pd.concat([near, far] if order != "far_first" else [far, near], ignore_index=True)
    if order == "shuffled":
        train = train.iloc[rng.permutation(len(train))].reset_index(drop=True)
    # test rows: both clusters, the gap between them, and a few outliers beyond
    test = pd.concat([util_data.generate_normal_data(600, seed=1), util_data.generate_normal_data(600, seed=3) + 25.0,
                      util_data.generate_normal_data(200, seed=4) + 12.5, util_data.generate_normal_data(100, seed=5) * 3.0],
                     ignore_index=True)
    cols = ["b"] + evidence
    cpd = pbn.CKDE("b", evidence) if evidence else pbn.KDE(["b"])
    cpd.fit(train)
    X, T = train[cols].to_numpy(), test[cols].to_numpy()
    H = oracle.bandwidth(X)
    # checker: the long-double evaluation.  For the outlying rows the joint and marginal log-likelihoods are large and
    # nearly cancel, and the reference arithmetic itself (oracle.ckde_logl) is only good to ~1e-13 of THEIR magnitude
    # (3e-10 of the difference on this data), so the 1e-10 bar is taken relative to the terms of the difference.
    joint = oracle.kde_logl_ld(X, T, H)
    marg = oracle.kde_logl_ld(X[:, 1:], T[:, 1:], H[1:, 1:]) if evidence else np.zeros_like(joint)
    want = joint - marg
    scale = np.maximum(np.maximum(np.abs(joint), np.abs(marg)), np.abs(want))
    got = cpd.logl(test)
    assert np.all(np.isfinite(want))
    assert np.max(np.abs(got - want) / scale) < RTOL64
    assert abs(cpd.slogl(test) - want.sum()) <= RTOL64 * abs(want.sum())
    ref, _ = (oracle.ckde_logl if evidence else oracle.kde_logl)(X, T, H)
    assert np.max(np.abs(got - ref) / scale) < RTOL64
