"""GPU parity tests for ProductKDE through the C ABI (pbn_product_kde_fit + pbn_kde_logl), shaped after
the reference's tests/factors/continuous/ProductKDE_test.py.  Checker: the CPU oracle (bit-exact with the reference's
kernels, tests/test_oracle_next.py) plus SciPy the way the reference's tests use it.

Tolerances (BASELINE.json north_star): 1e-10 relative in float64, 1e-4 in float32."""
import pickle

import numpy as np
import pyarrow as pa
import pytest
from scipy.stats import gaussian_kde

import oracle
import util_data

pytestmark = pytest.mark.gpu

RTOL64, RTOL32 = 1e-10, 1e-4
VARSETS = [["a"], ["b", "a"], ["c", "a", "b"], ["d", "a", "b", "c"]]
SIZE = 500
df = util_data.generate_normal_data(SIZE, seed=0)
df_float = df.astype("float32")


@pytest.fixture(scope="module")
def pbn():
    import pybnesian_b200 as pbn
    return pbn


def relerr(got, want, floor=1e-300):
    return np.max(np.abs(got - want) / np.maximum(np.abs(want), floor))


def py_nr_bandwidth(frame, variables):
    cov = np.atleast_2d(frame[variables].cov().to_numpy())
    delta = np.linalg.inv(np.diag(np.diag(cov))).dot(cov)
    delta_inv = np.linalg.inv(delta)
    N, d = frame.shape[0], len(variables)
    k = 4 * d * np.sqrt(np.linalg.det(delta)) / (2 * (delta_inv.dot(delta_inv)).trace() + delta_inv.trace() ** 2)
    return np.power(k / N, 2 / (d + 4)) * np.diag(cov)


def py_scott_bandwidth(frame, variables):
    return np.power(frame.shape[0], -2 / (len(variables) + 4)) * frame[variables].var().to_numpy()


def scipy_product(X, T):
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


def test_check_type(pbn):
    cpd = pbn.ProductKDE(["a"])
    cpd.fit(df)
    for fn in (cpd.logl, cpd.slogl):
        with pytest.raises(ValueError, match="Data type of training and test datasets is different."):
            fn(df_float)
    cpd.fit(df_float)
    for fn in (cpd.logl, cpd.slogl):
        with pytest.raises(ValueError, match="Data type of training and test datasets is different."):
            fn(df)


def test_variables_and_errors(pbn):
    for variables in VARSETS:
        assert pbn.ProductKDE(variables).variables() == variables
    with pytest.raises(ValueError, match="0 variables"):
        pbn.ProductKDE([])
    k = pbn.ProductKDE(["a"])
    with pytest.raises(ValueError, match="not fitted"):
        k.data_type()
    with pytest.raises(ValueError, match="not fitted"):
        k.logl(df)
    k.fit(df)
    assert k.data_type() == pa.float64()
    k.fit(df_float)
    assert k.data_type() == pa.float32()


def test_bandwidth(pbn):
    for variables in VARSETS:
        for instances in [50, 150, 500]:
            cpd = pbn.ProductKDE(variables)
            assert not cpd.fitted()
            cpd.fit(df.iloc[:instances])
            assert cpd.fitted() and cpd.num_instances() == instances and cpd.num_variables() == len(variables)
            assert np.allclose(cpd.bandwidth, py_nr_bandwidth(df[:instances], variables), rtol=1e-9)
            assert relerr(cpd.bandwidth, oracle.diag_bandwidth(df[variables].to_numpy()[:instances])) < 1e-11
            cpd.fit(df_float.iloc[:instances])
            assert np.allclose(cpd.bandwidth, py_nr_bandwidth(df[:instances], variables), atol=0.0005)
            cpd = pbn.ProductKDE(variables, pbn.ScottsBandwidth())
            cpd.fit(df.iloc[:instances])
            assert np.allclose(cpd.bandwidth, py_scott_bandwidth(df[:instances], variables), rtol=1e-11)
            cpd.fit(df_float.iloc[:instances])
            assert np.allclose(cpd.bandwidth, py_scott_bandwidth(df[:instances], variables), atol=0.0005)
    cpd = pbn.ProductKDE(["a"])
    cpd.fit(df)
    cpd.bandwidth = [1]
    assert cpd.bandwidth == np.asarray([1])
    with pytest.raises(ValueError, match="vector with shape"):
        cpd.bandwidth = [1, 2]


def test_python_bandwidth_selector(pbn):
    class UnitaryBandwidth(pbn.BandwidthSelector):
        def diag_bandwidth(self, frame, variables):
            return np.ones((len(variables),))

    kde = pbn.ProductKDE(["a", "b", "c", "d"], UnitaryBandwidth())
    kde.fit(df)
    assert np.all(kde.bandwidth == np.ones(4))
    X = df[["a", "b", "c", "d"]].to_numpy()
    want, _ = oracle.product_kde_logl(X, X[:40], np.ones(4))
    assert relerr(kde.logl(df.iloc[:40]), want) < RTOL64


@pytest.mark.parametrize("variables", VARSETS)
@pytest.mark.parametrize("n_train", [50, 500, 10000])
@pytest.mark.parametrize("rule", ["normal_reference", "scott"])
def test_logl_f64_vs_oracle(pbn, variables, n_train, rule):
    train = util_data.generate_normal_data(n_train, seed=0)
    test = util_data.generate_normal_data(50, seed=1)
    sel = pbn.NormalReferenceRule() if rule == "normal_reference" else pbn.ScottsBandwidth()
    k = pbn.ProductKDE(variables, sel)
    k.fit(train)
    X, T = train[variables].to_numpy(), test[variables].to_numpy()
    want, want_s = oracle.product_kde_logl(X, T, oracle.diag_bandwidth(X, rule))
    got = k.logl(test)
    assert relerr(got, want) < RTOL64
    assert abs(k.slogl(test) - want_s) <= RTOL64 * abs(want_s)
    if rule == "normal_reference":
        assert np.allclose(got, scipy_product(X, T), rtol=1e-9, atol=0)


@pytest.mark.parametrize("variables", VARSETS)
@pytest.mark.parametrize("n_train", [500, 10000])
def test_logl_f32_vs_oracle(pbn, variables, n_train):
    train = util_data.generate_normal_data(n_train, seed=0).astype("float32")
    test = util_data.generate_normal_data(50, seed=1).astype("float32")
    k = pbn.ProductKDE(variables)
    k.fit(train)
    X, T = train[variables].to_numpy(), test[variables].to_numpy()
    want, want_s = oracle.product_kde_logl(X, T, oracle.diag_bandwidth(X))
    assert relerr(k.logl(test), want, floor=1.0) < RTOL32
    assert abs(k.slogl(test) - want_s) <= RTOL32 * abs(want_s)


def test_variable_order_invariance(pbn):
    test = util_data.generate_normal_data(50, seed=1)
    a, b = pbn.ProductKDE(["d", "a", "b", "c"]), pbn.ProductKDE(["a", "c", "d", "b"])
    a.fit(df)
    b.fit(df)
    assert np.allclose(a.logl(test), b.logl(test), rtol=1e-9)
    assert np.isclose(a.slogl(test), b.slogl(test), rtol=1e-10)


def test_nulls(pbn):
    test = util_data.generate_normal_data(50, seed=1)
    rng = np.random.RandomState(0)
    test_null = test.copy()
    for c in "abcd":
        test_null.loc[test_null.index[rng.randint(0, 50, size=10)], c] = np.nan
    train_null = df.copy()
    for c in "abcd":
        train_null.loc[train_null.index[rng.randint(0, SIZE, size=100)], c] = np.nan
    for variables in VARSETS:
        k = pbn.ProductKDE(variables)
        k.fit(train_null)
        X = train_null[variables].dropna().to_numpy()
        assert k.num_instances() == X.shape[0]
        assert np.allclose(k.bandwidth, oracle.diag_bandwidth(X), rtol=1e-11)
        got = k.logl(test_null)
        isnull = test_null[variables].isna().any(axis=1).to_numpy()
        want, want_s = oracle.product_kde_logl(X, test_null[variables].to_numpy()[~isnull], k.bandwidth)
        assert np.all(np.isnan(got[isnull]))
        assert relerr(got[~isnull], want) < RTOL64
        assert abs(k.slogl(test_null) - want_s) <= RTOL64 * abs(want_s)


def test_pickle_and_dataset(pbn):
    k = pbn.ProductKDE(["c", "a", "b"], pbn.ScottsBandwidth())
    k.fit(df)
    test = util_data.generate_normal_data(64, seed=1)
    k2 = pickle.loads(pickle.dumps(k))
    assert k2.fitted() and k2.variables() == ["c", "a", "b"] and k2.num_instances() == SIZE
    assert np.array_equal(k2.bandwidth, k.bandwidth)
    assert np.array_equal(k2.logl(test), k.logl(test))
    assert np.array_equal(k.dataset().to_pandas().to_numpy(), df[["c", "a", "b"]].to_numpy())
    u = pickle.loads(pickle.dumps(pbn.ProductKDE(["a"])))
    assert not u.fitted()
