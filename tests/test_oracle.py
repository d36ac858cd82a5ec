"""CPU tests (no GPU): the oracle (oracle/pbn_oracle.cpp) against
  * the committed golden vectors (tests/golden/kde_golden.npz: outputs of the reference's own
    OpenCL-C kernels run through oracle/ref_shim, SciPy, and libstdc++ shuffles),
  * SciPy, the way the reference's own tests check KDE/CKDE (KDE_test.py:167-203, CKDE_test.py:146-179),
  * a long-double direct evaluation,
  * oracle/_ref (the reference kernels compiled as C++) when it was built in this tree.
"""
import os

import numpy as np
import pytest
from scipy.stats import gaussian_kde, norm

import oracle
import util_data

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "kde_golden.npz"))
VARSETS = [["a"], ["b", "a"], ["c", "a", "b"], ["d", "a", "b", "c"]]
CASES = [(500, 50), (10, 50), (300, 70)]


def nr_factor(s):
    return np.power(4 / (s.d + 2), 1 / (s.d + 4)) * s.scotts_factor()


def data(variables, N, m, dt):
    X = util_data.generate_normal_data(N, 0)[variables].to_numpy().astype(dt)
    T = util_data.generate_normal_data(m, 1)[variables].to_numpy().astype(dt)
    return X, T


@pytest.mark.parametrize("dt", ["float64", "float32"])
@pytest.mark.parametrize("variables", VARSETS)
@pytest.mark.parametrize("N,m", CASES)
def test_oracle_bit_exact_with_reference_kernels_golden(dt, variables, N, m):
    """The standalone restatement reproduces the reference kernels' outputs bit for bit."""
    if N <= len(variables):
        pytest.skip("not enough instances")
    X, T = data(variables, N, m, dt)
    key = "%s_%s_%d_%d" % (dt, "".join(variables), N, m)
    H = oracle.bandwidth(X)
    assert np.array_equal(H, GOLD["H_" + key])
    logl, slogl = oracle.kde_logl(X, T, H)
    assert np.array_equal(logl, GOLD["ref_kde_logl_" + key])
    assert slogl == float(GOLD["ref_kde_slogl_" + key])
    cl, cs = oracle.ckde_logl(X, T, H)
    assert np.array_equal(cl, GOLD["ref_ckde_logl_" + key])
    assert cs == float(GOLD["ref_ckde_slogl_" + key])


@pytest.mark.parametrize("variables", VARSETS)
@pytest.mark.parametrize("N,m", CASES)
def test_oracle_vs_scipy_golden_and_live(variables, N, m):
    if N <= len(variables):
        pytest.skip("not enough instances")
    X, T = data(variables, N, m, "float64")
    key = "float64_%s_%d_%d" % ("".join(variables), N, m)
    H = oracle.bandwidth(X)
    assert np.allclose(H, GOLD["scipy_H_" + key], rtol=1e-12, atol=0)
    logl, _ = oracle.kde_logl(X, T, H)
    assert np.allclose(logl, GOLD["scipy_kde_logl_" + key], rtol=1e-10, atol=1e-12)
    sk = gaussian_kde(X.T, bw_method=nr_factor)
    assert np.allclose(logl, sk.logpdf(T.T), rtol=1e-10, atol=1e-12)
    if len(variables) > 1:
        cl, _ = oracle.ckde_logl(X, T, H)
        assert np.allclose(cl, GOLD["scipy_ckde_logl_" + key], rtol=1e-9, atol=1e-11)


@pytest.mark.parametrize("dt,tol", [("float64", 1e-12), ("float32", 2e-5)])
def test_oracle_vs_long_double(dt, tol):
    for variables in VARSETS:
        X, T = data(variables, 800, 40, dt)
        H = oracle.bandwidth(X)
        a, _ = oracle.kde_logl(X, T, H)
        b = oracle.kde_logl_ld(X, T, H)
        assert np.max(np.abs(a - b) / np.maximum(np.abs(b), 1.0)) < tol


@pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_vs_ref_kernels_live():
    for dt in ("float64", "float32"):
        X, T = data(["c", "a", "b"], 120, 66, dt)
        H = oracle.bandwidth(X)
        a, sa = oracle.kde_logl(X, T, H)
        b, sb = oracle.ref_kde_logl(X, T, H)
        assert np.array_equal(a, b) and sa == sb
        a, sa = oracle.ckde_logl(X, T, H)
        b, sb = oracle.ref_ckde_logl(X, T, H)
        assert np.array_equal(a, b) and sa == sb
        # (the scalar epilogue of the shim path factorises H with LAPACK: last-bit differences allowed)
        u, r = oracle.ucv_score_unconstrained(X, H), oracle.ref_ucv_score_unconstrained(X, H)
        assert abs(u - r) <= (1e-13 if dt == 'float64' else 1e-6) * abs(r)


@pytest.mark.parametrize("dt", ["float64", "float32"])
@pytest.mark.parametrize("variables", VARSETS)
def test_ucv_score_golden(dt, variables):
    X = util_data.generate_normal_data(200, 0)[variables].to_numpy().astype(dt)
    H = oracle.bandwidth(X)
    key = "%s_%s_200" % (dt, "".join(variables))
    tol = 1e-13 if dt == "float64" else 1e-6
    for name, Hs in (("ref_ucv_", H), ("ref_ucv_half_", 0.5 * H)):
        want = float(GOLD[name + key])
        assert abs(oracle.ucv_score_unconstrained(X, Hs) - want) <= tol * abs(want)


def test_ucv_score_vs_direct_formula():
    """N*UCV(H) = e^{c2} + 2/N sum_{i<j} e^{-s/4+c2} - 4/(N-1) sum_{i<j} e^{-s/2+c1} (kde/UCV.cpp:296-358)."""
    X = util_data.generate_normal_data(150, 0)[["a", "b", "c"]].to_numpy()
    H = oracle.bandwidth(X)
    N, d = X.shape
    L = np.linalg.cholesky(H)
    Y = np.linalg.solve(L, X.T).T
    D2 = ((Y[:, None, :] - Y[None, :, :]) ** 2).sum(-1)
    iu = np.triu_indices(N, 1)
    c1 = -np.log(np.diag(L)).sum() - 0.5 * d * np.log(2 * np.pi)
    c2 = c1 - 0.5 * d * np.log(2)
    want = np.exp(c2) + 2 * np.exp(-0.25 * D2[iu] + c2).sum() / N - 4 * np.exp(-0.5 * D2[iu] + c1).sum() / (N - 1)
    assert abs(oracle.ucv_score_unconstrained(X, H) - want) < 1e-10 * abs(want)
    hd = np.diag(H)
    Yd = X / np.sqrt(hd)
    D2 = ((Yd[:, None, :] - Yd[None, :, :]) ** 2).sum(-1)
    c1 = -0.5 * np.log(hd).sum() - 0.5 * d * np.log(2 * np.pi)
    c2 = c1 - 0.5 * d * np.log(2)
    want = np.exp(c2) + 2 * np.exp(-0.25 * D2[iu] + c2).sum() / N - 4 * np.exp(-0.5 * D2[iu] + c1).sum() / (N - 1)
    assert abs(oracle.ucv_score_diagonal(X, hd) - want) < 1e-10 * abs(want)


def test_cv_indices_golden_and_properties():
    for n, k, seed in [(10, 3, 0), (23, 10, 0), (1000, 10, 0), (1000, 7, 123)]:
        idx, lim = oracle.cv_indices(np.arange(n), k, seed)
        assert np.array_equal(idx, GOLD["cv_idx_%d_%d_%d" % (n, k, seed)])
        assert np.array_equal(lim, GOLD["cv_lim_%d_%d_%d" % (n, k, seed)])
        assert sorted(idx) == list(range(n)) and lim[0] == 0 and lim[-1] == n
        sizes = np.diff(lim)
        assert sizes.max() - sizes.min() <= 1 and np.all(np.diff(sizes) <= 0)
    idx, lim = oracle.cv_indices(np.arange(100000), 10, 0)
    assert np.array_equal(idx[:64], GOLD["cv_idx_100000_10_0"])
    assert int(np.sum(idx.astype(np.int64) * np.arange(1, 100001) % 1000003)) == int(GOLD["cv_idxsum_100000_10_0"][0])
    # SURVEY.md §7 known answers (libstdc++ std::shuffle, mt19937{0})
    assert list(oracle.cv_indices(np.arange(10), 3, 0)[0]) == [0, 2, 1, 5, 9, 8, 4, 7, 6, 3]


def test_holdout_indices():
    tr, te = oracle.holdout_indices(np.arange(1000), 0.2, 0)
    assert len(te) == 200 and len(tr) == 800 and sorted(np.concatenate([tr, te])) == list(range(1000))
    idx, _ = oracle.cv_indices(np.arange(1000), 10, 0)  # same shuffle for the same seed
    assert np.array_equal(np.concatenate([tr, te]), idx)
    tr, te = oracle.holdout_indices(np.arange(11), 0.5, 3)
    assert len(te) == 6  # std::round(5.5) == 6


@pytest.mark.parametrize("p", [0, 1, 2, 3])
def test_linear_gaussian_vs_lstsq(p):
    """Same check as the reference's cvlikelihood_test.py:12-49 (numpy lstsq + norm.logpdf)."""
    df = util_data.generate_normal_data(2000, 0)
    te = util_data.generate_normal_data(300, 1)
    y, parents = df["d"].to_numpy(), [df[c].to_numpy() for c in ["a", "b", "c"][:p]]
    beta, var = oracle.lg_fit(y, parents)
    A = np.column_stack([np.ones(len(y))] + parents)
    b_np, res, _, _ = np.linalg.lstsq(A, y, rcond=None)
    assert np.allclose(beta, b_np, rtol=1e-8, atol=1e-10)
    var_np = np.sum((y - A @ b_np) ** 2) / (len(y) - p - 1)
    assert abs(var - var_np) < 1e-10 * var_np
    yt, pt = te["d"].to_numpy(), [te[c].to_numpy() for c in ["a", "b", "c"][:p]]
    logl, slogl = oracle.lg_logl(yt, pt, beta, var)
    At = np.column_stack([np.ones(len(yt))] + pt)
    want = norm(At @ b_np, np.sqrt(var_np)).logpdf(yt)
    assert np.allclose(logl, want, rtol=1e-9, atol=1e-11)
    assert abs(slogl - want.sum()) < 1e-10 * abs(want.sum())


def test_cv_score_composition():
    """CVLikelihood.local_score = sum over folds of slogl(fit on train fold) (cv_likelihood.cpp:16-25)."""
    df = util_data.generate_normal_data(600, 0)
    X = df[["c", "a", "b"]].to_numpy()
    idx, lim = oracle.cv_indices(np.arange(600), 5, 0)
    got = oracle.cv_score(X, idx, lim, "ckde")
    want = 0.0
    for f in range(5):
        te = idx[lim[f]:lim[f + 1]]
        tr = np.concatenate([idx[:lim[f]], idx[lim[f + 1]:]])
        want += oracle.ckde_logl(X[tr], X[te], oracle.bandwidth(X[tr]))[1]
    assert got == want
    got = oracle.cv_score(X, idx, lim, "lg")
    want = 0.0
    for f in range(5):
        te = idx[lim[f]:lim[f + 1]]
        tr = np.concatenate([idx[:lim[f]], idx[lim[f + 1]:]])
        beta, var = oracle.lg_fit(X[tr, 0], [X[tr, 1], X[tr, 2]])
        want += oracle.lg_logl(X[te, 0], [X[te, 1], X[te, 2]], beta, var)[1]
    assert got == want


def test_singular_covariance():
    X = util_data.generate_normal_data(100, 0)[["a", "b"]].to_numpy().copy()
    X[:, 1] = 2 * X[:, 0]
    with pytest.raises(oracle.SingularCovariance):
        oracle.bandwidth(X)
    with pytest.raises(oracle.SingularCovariance):
        oracle.bandwidth(X[:2])


FLOOR_GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "floor_golden.npz"))


@pytest.mark.parametrize("dt", ["float64", "float32"])
@pytest.mark.parametrize("order", ["near_first", "far_first", "shuffled"])
@pytest.mark.parametrize("variables", [["b"], ["b", "a"], ["b", "a", "c", "d"]])
def test_oracle_bit_exact_on_two_cluster_golden(dt, order, variables):
    """Two clusters 25 sigma apart, test rows in both, between and beyond (tests/golden/make_golden_floor.py): almost every
    kernel term of a row is negligible and the joint / marginal log-likelihoods of the outlying rows nearly cancel.  The
    restatement must still reproduce the reference kernels bit for bit, in every training-row order."""
    train, test = util_data.two_cluster_frames(order)
    X, T = train[variables].to_numpy().astype(dt), test[variables].to_numpy().astype(dt)
    H = oracle.bandwidth(X)
    key = "%s_%s_%s" % (dt, order, "".join(variables))
    logl, _ = oracle.kde_logl(X, T, H)
    assert np.array_equal(logl, FLOOR_GOLD["ref_kde_logl_" + key], equal_nan=True)
    if len(variables) > 1:
        cl, _ = oracle.ckde_logl(X, T, H)
        assert np.array_equal(cl, FLOOR_GOLD["ref_ckde_logl_" + key], equal_nan=True)
        if dt == "float64":  # and the long-double value, to the accuracy the reference arithmetic has on these rows
            ld = oracle.kde_logl_ld(X, T, H) - oracle.kde_logl_ld(X[:, 1:], T[:, 1:], H[1:, 1:])
            scale = np.maximum(np.abs(oracle.kde_logl_ld(X, T, H)), 1.0)
            assert np.max(np.abs(cl - ld) / scale) < 1e-11
