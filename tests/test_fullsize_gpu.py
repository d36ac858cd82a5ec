"""Parity at BASELINE.json's full size (configs[1]: CKDE('d' | a, b, c), float64, 1M train x 1M test rows),
where the oracle cannot evaluate every row: a test-row sub-sample against the oracle, plus properties that do
not depend on the size - additivity of the kernel sum over a partition of the training rows, additivity of
slogl over test shards, CKDE = joint - marginal, and invariance to the order of the training rows."""
import numpy as np
import pytest

import oracle
import util_data

pytestmark = pytest.mark.gpu

N = 1_000_000
VARS = ["d", "a", "b", "c"]


@pytest.fixture(scope="module")
def setup():
    import pybnesian_b200 as pbn
    train = util_data.generate_normal_data(N, seed=0)
    test = util_data.generate_normal_data(N, seed=1)
    ftrain, ftest = pbn.DataFrame(train), pbn.DataFrame(test)
    cpd = pbn.CKDE("d", ["a", "b", "c"])
    cpd.fit(ftrain)
    logl = cpd.logl(ftest)
    return pbn, train, test, ftrain, ftest, cpd, logl


def test_subsample_against_oracle(setup):
    pbn, train, test, ftrain, ftest, cpd, logl = setup
    rows = np.random.default_rng(0).choice(N, 512, replace=False)
    X = train[VARS].to_numpy()
    H = oracle.bandwidth(X)
    assert np.allclose(cpd.kde_joint().bandwidth, H, rtol=1e-11, atol=0)
    T = test.iloc[rows][VARS].to_numpy()
    want, _ = oracle.ckde_logl(X, T, H)
    # a CKDE log-likelihood is the difference of two log-sums of magnitude ~10 and crosses zero (|logl| down to
    # 1e-3 in this sample), so its relative error is unbounded by construction; the two terms are held to the
    # strict relative bar and the difference to 1e-10 relative + 1e-12 absolute (the reference's own tests use
    # np.isclose, i.e. atol = 1e-8)
    joint_want, _ = oracle.kde_logl(X, T, H)
    joint_got = cpd.kde_joint().logl(pbn.DataFrame(test.iloc[rows]))
    assert np.max(np.abs(joint_got - joint_want) / np.abs(joint_want)) < 1e-10
    assert np.all(np.abs(logl[rows] - want) <= 1e-12 + 1e-10 * np.abs(want))
    assert pbn.default_context().last_fallback_rows() == 0


def test_slogl_is_additive_over_test_shards(setup):
    pbn, train, test, ftrain, ftest, cpd, logl = setup
    total = cpd.slogl(ftest)
    assert abs(total - logl.sum()) <= 1e-12 * abs(total)
    cut = 333_337
    parts = cpd.slogl(pbn.DataFrame(test.iloc[:cut])) + cpd.slogl(pbn.DataFrame(test.iloc[cut:]))
    assert abs(total - parts) <= 1e-12 * abs(total)


def test_ckde_is_joint_minus_marginal(setup):
    pbn, train, test, ftrain, ftest, cpd, logl = setup
    sub = pbn.DataFrame(test.iloc[:200_000])
    joint, marg = cpd.kde_joint().logl(sub), cpd.kde_marg().logl(sub)
    assert np.allclose(joint - marg, logl[:200_000], rtol=1e-10, atol=1e-12)


def test_kernel_sum_is_additive_over_training_partition_and_order(setup):
    pbn, train, test, ftrain, ftest, cpd, logl = setup
    sub = pbn.DataFrame(test.iloc[:100_000])
    H = cpd.kde_joint().bandwidth
    full = cpd.kde_joint().logl(sub)
    cut = 400_003
    parts = []
    for chunk in (train.iloc[:cut], train.iloc[cut:]):
        k = pbn.KDE(VARS)
        k.fit(pbn.DataFrame(chunk))
        k.bandwidth = H                      # same kernel, different training rows
        parts.append(k.logl(sub) + np.log(len(chunk)))
    combined = np.logaddexp(parts[0], parts[1]) - np.log(N)
    assert np.allclose(combined, full, rtol=1e-10, atol=1e-12)
    perm = np.random.default_rng(1).permutation(N)
    k = pbn.KDE(VARS)
    k.fit(pbn.DataFrame(train.iloc[perm]))
    k.bandwidth = H
    assert np.allclose(k.logl(sub), full, rtol=1e-11, atol=1e-12)


# ---- CKDE.cdf / CKDE.sample at the full training size ---------------------------------------------
def test_cdf_full_training_set_subsample_and_properties(setup):
    """1M training rows: a 48-row sub-sample against the oracle; the cdf is a probability, monotone in the variable for
    fixed evidence, and tends to 0 / 1 far below / above the data."""
    pbn, train, test, ftrain, ftest, cpd, logl = setup
    sub = test.iloc[np.random.default_rng(3).choice(N, 48, replace=False)]
    got = cpd.cdf(sub)
    X = train[VARS].to_numpy()
    want = oracle.ckde_cdf(X, sub[VARS].to_numpy(), cpd.kde_joint().bandwidth)
    assert np.allclose(got, want, rtol=1e-10, atol=1e-10)
    block = test.iloc[:50_000]
    c = cpd.cdf(block)
    assert np.all((c >= 0) & (c <= 1)) and 0.45 < c.mean() < 0.55
    # fixed evidence, increasing d
    base = test.iloc[:200].copy()
    lo, hi = base.copy(), base.copy()
    lo["d"] = base["d"] - 0.25
    hi["d"] = base["d"] + 0.25
    c0, c1, c2 = cpd.cdf(lo), cpd.cdf(base), cpd.cdf(hi)
    assert np.all(c0 <= c1) and np.all(c1 <= c2) and np.any(c0 < c2)
    far = base.copy()
    far["d"] = base["d"] + 50.0
    assert np.allclose(cpd.cdf(far), 1.0, rtol=0, atol=1e-12)
    far["d"] = base["d"] - 50.0
    assert np.allclose(cpd.cdf(far), 0.0, rtol=0, atol=1e-12)


def test_sample_full_training_set(setup):
    """Indices chosen against 1M training rows equal the oracle's on a 64-row sub-sample; the sampled values follow the
    conditional law of the generator (d | a, b, c is linear-Gaussian with sd 0.5)."""
    pbn, train, test, ftrain, ftest, cpd, logl = setup
    n = 64
    ev = test.iloc[:n][["a", "b", "c"]]
    arr, idx = cpd.sample(n, ev, 123, _return_indices=True)
    X = train[VARS].to_numpy()
    want, want_idx = oracle.ckde_sample(X, cpd.kde_joint().bandwidth, ev.to_numpy(), n, 123)
    assert np.array_equal(idx, want_idx)
    assert np.allclose(arr.to_numpy(), want, rtol=1e-12, atol=1e-10)
    n = 20_000
    ev = test.iloc[:n][["a", "b", "c"]]
    s, idx = cpd.sample(n, ev, 7, _return_indices=True)
    assert idx.min() >= 0 and idx.max() < N
    resid = s.to_numpy() - (1.5 - 0.9 * ev["a"] + 5.6 * ev["b"] + 0.3 * ev["c"]).to_numpy()
    assert abs(resid.mean()) < 0.05 and 0.4 < resid.std() < 0.75
