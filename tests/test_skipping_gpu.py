"""Tile skipping (pybnesian_b200/csrc/spatial.cu; on by default for large single-model calls): Morton-ordered rows, one
bounding box per tile, units whose boxes prove every dropped term of a row below 2^-40 of its sum are not evaluated; in
float64 the units that stay also skip groups of training points whose terms are below 2^-64 of the running sums for every
row of a warp (pair_kernel.cuh: tile_f64_dot_gskip - families with two or more kernel coordinates, exercised here by KDE(b, a),
KDE(a, b, c) and CKDE(c | a, b)).
The reference evaluates every pair (kde/KDE.hpp:592-640); results must agree with the all-pairs evaluation far inside the
1e-10 (float64) / 1e-4 (float32) bars, in the caller's row order, whatever the data look like."""
import numpy as np
import pandas as pd
import pytest

import oracle
import util_data

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pbn():
    """One device for this module: the unit counts asserted below are those of one device's share (on a multi-GPU box the
    default context spreads the test rows, and a share may fall under the size threshold of skipping)."""
    import pybnesian_b200 as pbn
    prev = pbn.default_context()
    pbn.set_default_context(pbn.Context(prev.device) if prev.num_devices > 1 else prev)
    yield pbn
    pbn.default_context().set_skipping(True)
    pbn.set_default_context(prev)


def both(pbn, fn):
    ctx = pbn.default_context()
    ctx.set_skipping(False)
    off = fn()
    ctx.set_skipping(True)
    on = fn()
    return off, on, ctx.skip_stats()


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("kind,variables", [("kde", ["a"]), ("kde", ["b", "a"]), ("kde", ["d", "a", "b", "c"]),
                                            ("ckde", ["b", "a"]), ("ckde", ["d", "a", "b", "c"])])
def test_skipping_matches_all_pairs(pbn, kind, variables, dtype):
    n, m = 300_000, 100_003
    tr = util_data.generate_normal_data(n, 0).astype(dtype)
    te = util_data.generate_normal_data(m, 1).astype(dtype)
    f = pbn.KDE(variables) if kind == "kde" else pbn.CKDE(variables[0], variables[1:])
    ftr, fte = pbn.DataFrame(tr), pbn.DataFrame(te)
    f.fit(ftr)
    off, on, st = both(pbn, lambda: f.logl(fte))
    # something was skipped (families of 1-2 variables; with 4 variables at this size the boxes prove too little and the
    # call falls back to evaluating every pair in the caller's order) ...
    assert st["last_evaluated"] < st["last_total"] if len(variables) <= 2 else st["last_evaluated"] <= st["last_total"], st
    scale = np.maximum(np.abs(off), 1.0)
    # float64: only the order of the partial sums changes.  float32: the Morton order regroups the per-tile FLOAT partial
    # sums (near terms now meet in the same tiles instead of being absorbed one by one into a large partial sum), which
    # moves a row by a few float ulps - still two orders inside the 1e-4 bar
    tol = 1e-12 if dtype == "float64" else 5e-6
    assert np.max(np.abs(on - off) / scale) < tol                  # ... and nothing that matters
    s_off, s_on, _ = both(pbn, lambda: f.slogl(fte))
    assert abs(s_on - s_off) <= (1e-12 if dtype == "float64" else 2e-6) * abs(s_off)
    assert abs(s_on - on.sum()) <= 1e-12 * abs(s_on)
    # and against the oracle on a sub-sample (row order: the caller's)
    rows = np.random.default_rng(1).choice(m, 200, replace=False)
    X, T = tr[variables].to_numpy(), te[variables].to_numpy()[rows]
    if dtype == "float64":
        H = oracle.bandwidth(X)
        want = (oracle.kde_logl if kind == "kde" else oracle.ckde_logl)(X, T, H)[0]
        assert np.all(np.abs(on[rows] - want) <= 1e-12 + 1e-10 * np.abs(want))
    else:
        # at 300k rows the bandwidth of these collinear columns is small enough for the reference's float arithmetic
        # (forward substitution on raw float differences, reproduced by the float oracle) to be ~1e-4 off by itself;
        # the float32 result is held to the 1e-4 bar against the float64 evaluation of the same data and bandwidth
        H = np.asarray(f.bandwidth if kind == "kde" else f.kde_joint().bandwidth, dtype=np.float64)
        want = (oracle.kde_logl if kind == "kde" else oracle.ckde_logl)(X.astype(np.float64), T.astype(np.float64), H)[0]
        assert np.all(np.abs(on[rows] - want) <= 1e-4 * np.maximum(np.abs(want), 1.0))


def test_skipping_clustered_heavy_tailed_and_far_rows(pbn):
    """Data built to stress the bounds: two clusters 40 sigma apart plus Student-t tails, and a test set that mixes
    ordinary rows, rows between the clusters and rows far from everything (these take the shifted second pass, in
    Morton order, and must come back in the caller's order)."""
    rng = np.random.default_rng(3)
    n, m = 400_000, 60_000
    a = np.concatenate([rng.standard_t(3, n // 2), 40.0 + rng.standard_normal(n - n // 2)])
    b = 0.5 * a + np.concatenate([rng.standard_normal(n // 2), 0.2 * rng.standard_t(4, n - n // 2)])
    c = rng.standard_normal(n) - 0.3 * b
    tr = pd.DataFrame({"a": a, "b": b, "c": c}).iloc[rng.permutation(n)].reset_index(drop=True)
    ta = np.concatenate([rng.standard_t(3, m // 3), 40.0 + rng.standard_normal(m // 3), rng.uniform(-150, 200, m - 2 * (m // 3))])
    tb = 0.5 * ta + rng.standard_normal(m)
    tc = rng.standard_normal(m) - 0.3 * tb
    te = pd.DataFrame({"a": ta, "b": tb, "c": tc}).iloc[rng.permutation(m)].reset_index(drop=True)
    ftr, fte = pbn.DataFrame(tr), pbn.DataFrame(te)
    for f in (pbn.KDE(["a", "b", "c"]), pbn.CKDE("c", ["a", "b"]), pbn.KDE(["a"])):
        f.fit(ftr)
        off, on, st = both(pbn, lambda: f.logl(fte))
        assert np.all(np.isfinite(off)) and st["last_evaluated"] < st["last_total"]
        assert np.max(np.abs(on - off) / np.maximum(np.abs(off), 1.0)) < 1e-12
        assert pbn.default_context().last_fallback_rows() > 0
    # a NaN / null test row stays where it was
    te2 = te.copy()
    te2.loc[12345, "a"] = np.nan
    f = pbn.KDE(["a", "b", "c"])
    f.fit(ftr)
    got = f.logl(te2)
    assert np.isnan(got[12345]) and np.isfinite(np.delete(got, 12345)).all()
    assert np.allclose(np.delete(got, 12345), np.delete(on if False else f.logl(te), 12345), rtol=1e-13, atol=1e-13)


def test_skipping_is_off_below_the_size_thresholds_and_switchable(pbn):
    tr = util_data.generate_normal_data(50_000, 0)
    te = util_data.generate_normal_data(20_000, 1)
    k = pbn.KDE(["a", "b"])
    k.fit(tr)
    k.logl(te)
    st = pbn.default_context().skip_stats()
    assert st["last_evaluated"] == st["last_total"]              # 50k training rows: every unit evaluated
    big = util_data.generate_normal_data(300_000, 0)
    k.fit(big)
    ctx = pbn.default_context()
    ctx.set_skipping(False)
    k.logl(te)
    st = ctx.skip_stats()
    assert st["last_evaluated"] == st["last_total"]
    ctx.set_skipping(True)
    k.logl(te)
    st = ctx.skip_stats()
    assert st["last_evaluated"] < st["last_total"]
