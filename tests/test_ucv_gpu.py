"""GPU parity tests for the UCV objective (UCVScorer) and the UCV bandwidth selector.
The reference has NO test for UCV; the objective is pinned by the oracle, which
is bit-exact with the reference's own kernels (tests/test_oracle.py).  The optimiser (NLopt
Nelder-Mead in the reference) is unpinned: only optimality properties are checked."""
import os

import numpy as np
import pytest

import oracle
import util_data

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "kde_golden.npz"))
VARSETS = [["a"], ["b", "a"], ["c", "a", "b"], ["d", "a", "b", "c"]]


@pytest.fixture(scope="module")
def pbn():
    import pybnesian_b200 as pbn
    return pbn


@pytest.mark.parametrize("variables", VARSETS)
@pytest.mark.parametrize("n", [2, 7, 200, 1500, 3001])
def test_ucv_score_f64(pbn, variables, n):
    df = util_data.generate_normal_data(max(n, 50), 0).iloc[:n]
    X = df[variables].to_numpy()
    H = oracle.bandwidth(util_data.generate_normal_data(500, 0)[variables].to_numpy())
    sc = pbn.UCVScorer(df, variables)
    assert sc.num_pairs() == n * (n - 1) // 2
    for Hs in (H, 0.3 * H, 4.0 * H):
        want = oracle.ucv_score_unconstrained(X, Hs)
        got = sc.score_unconstrained(Hs)
        assert abs(got - want) <= 1e-10 * abs(want)
        hd = np.diag(Hs)
        want = oracle.ucv_score_diagonal(X, hd)
        assert abs(sc.score_diagonal(hd) - want) <= 1e-10 * abs(want)


@pytest.mark.parametrize("variables", VARSETS)
def test_ucv_score_golden_reference_kernels(pbn, variables):
    for dt, tol in (("float64", 1e-10), ("float32", 1e-4)):
        df = util_data.generate_normal_data(200, 0).astype(dt)
        H = oracle.bandwidth(df[variables].to_numpy())
        key = "%s_%s_200" % (dt, "".join(variables))
        sc = pbn.UCVScorer(df, variables)
        for name, Hs in (("ref_ucv_", H), ("ref_ucv_half_", 0.5 * H)):
            want = float(GOLD[name + key])
            assert abs(sc.score_unconstrained(Hs) - want) <= tol * abs(want)


def test_ucv_score_f32(pbn):
    df = util_data.generate_normal_data(2500, 0).astype("float32")
    for variables in VARSETS:
        X = df[variables].to_numpy()
        H = oracle.bandwidth(X)
        sc = pbn.UCVScorer(df, variables)
        want = oracle.ucv_score_unconstrained(X, H)
        assert abs(sc.score_unconstrained(H) - want) <= 1e-4 * abs(want)


def test_ucv_pair_sums_partition(pbn):
    """Slices of the tile schedule (the multi-GPU split) add up to the whole."""
    df = util_data.generate_normal_data(5000, 0)
    variables = ["a", "b", "c"]
    H = oracle.bandwidth(df[variables].to_numpy())
    sc = pbn.UCVScorer(df, variables)
    s2, s1 = sc.pair_sums(H)
    for nparts in (2, 3, 8):
        parts = [sc.pair_sums(H, p, nparts) for p in range(nparts)]
        assert abs(sum(p[0] for p in parts) - s2) <= 1e-12 * s2
        assert abs(sum(p[1] for p in parts) - s1) <= 1e-12 * s1


def test_ucv_errors(pbn):
    df = util_data.generate_normal_data(100, 0)
    sc = pbn.UCVScorer(df, ["a", "b"])
    with pytest.raises(ValueError, match="Wrong dimension for bandwidth matrix"):
        sc.score_unconstrained(np.eye(3))
    with pytest.raises(ValueError, match="Wrong dimension for bandwidth vector"):
        sc.score_diagonal([1.0])


@pytest.mark.parametrize("variables", [["a"], ["a", "b"], ["c", "a", "b"]])
def test_ucv_bandwidth_improves_objective(pbn, variables):
    df = util_data.generate_normal_data(1000, 0)
    sel = pbn.UCV()
    H = sel.bandwidth(df, variables)
    d = len(variables)
    assert H.shape == (d, d) and np.allclose(H, H.T) and np.all(np.linalg.eigvalsh(H) > 0)
    sc = pbn.UCVScorer(df, variables)
    H0 = pbn.NormalReferenceRule().bandwidth(df, variables)
    s_opt, s_start = sc.score_unconstrained(H), sc.score_unconstrained(H0)
    assert s_opt <= s_start
    assert sel.last_evaluations > d
    # local optimality at the tolerance the reference asks of NLopt (ftol_rel = xtol_rel = 1e-4)
    L = np.linalg.cholesky(H)
    for eps in (0.02, -0.02):
        Lp = L.copy(); Lp[0, 0] *= (1 + eps)
        assert sc.score_unconstrained(Lp @ Lp.T) >= s_opt - 1e-3 * abs(s_opt)
    # usable as a KDE bandwidth selector
    k = pbn.KDE(variables, pbn.UCV()); k.fit(df)
    assert np.allclose(k.bandwidth, H, rtol=1e-12)
    h = sel.diag_bandwidth(df, variables)
    assert h.shape == (d,) and np.all(h > 0)
    assert sc.score_diagonal(h) <= sc.score_diagonal(pbn.NormalReferenceRule().diag_bandwidth(df, variables)) + 1e-12
