"""GPU parity tests for hybrid factors: HCKDE / CLinearGaussianCPD (DiscreteAdaptator over the
segmented multi-job launch: pbn_discrete_slices + pbn_table_take + pbn_kde_logl_multi), hybrid SemiparametricBN
fit / logl, likelihood scores and hill climbing on data with discrete columns.

Checkers: the CPU oracle's DiscreteAdaptator restatement (oracle/hybrid.py), the committed golden vectors of the
reference's own kernels run per configuration (tests/golden/hybrid_golden.npz) and SciPy per configuration.
Tolerances (BASELINE.json north_star): 1e-10 relative in float64, 1e-4 in float32."""
import ctypes
import os
import pickle

import numpy as np
import pyarrow as pa
import pytest

import oracle
from oracle import hybrid as ohy
import util_data

pytestmark = pytest.mark.gpu

RTOL64, RTOL32 = 1e-10, 1e-4
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "hybrid_golden.npz"))
EVIDENCE = [["A", "C", "B"], ["A"], ["B", "A"], ["C", "B"]]
CASES = [(600, 80), (300, 120)]


@pytest.fixture(scope="module")
def pbn():
    import pybnesian_b200 as pbn
    return pbn


def frames(N, m, dt, seeds=(0, 1)):
    tr, te = util_data.generate_hybrid_data(N, seeds[0]), util_data.generate_hybrid_data(m, seeds[1])
    for df in (tr, te):
        df["C"] = df["C"].astype(dt)
        df["D"] = df["D"].astype(dt)
    return tr, te


def close(got, want, rtol):
    got, want = np.asarray(got), np.asarray(want)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    ok = ~np.isnan(want)
    # float32: values near a zero crossing carry the float accumulation error of the reference arithmetic in
    # absolute terms (the reference's own float tests use atol 5e-4, CKDE_test.py)
    assert np.allclose(got[ok], want[ok], rtol=rtol, atol=5e-4 if rtol >= RTOL32 else rtol * 1e-2)


def with_nulls(df, seed):
    rng = np.random.default_rng(seed)
    df = df.copy()
    n = len(df)
    df.loc[rng.choice(n, n // 15, replace=False), "A"] = np.nan
    df.loc[rng.choice(n, n // 20, replace=False), "B"] = np.nan
    df.loc[rng.choice(n, n // 25, replace=False), "C"] = np.nan
    df.loc[rng.choice(n, n // 30, replace=False), "D"] = np.nan
    return df


@pytest.mark.parametrize("dt,rtol", [("float64", RTOL64), ("float32", RTOL32)])
@pytest.mark.parametrize("ev", EVIDENCE)
@pytest.mark.parametrize("N,m", CASES)
def test_hckde_logl_slogl_vs_oracle_and_reference_kernel_golden(pbn, dt, rtol, ev, N, m):
    tr, te = frames(N, m, dt)
    key = "%s_%s_%d_%d" % (dt, "".join(ev), N, m)
    f = pbn.HCKDE("D", ev)
    f.fit(tr)
    assert f.fitted() and f.data_type() == (pa.float64() if dt == "float64" else pa.float32())
    o = ohy.HybridFactor("D", ev).fit(tr)
    got = f.logl(te)
    close(got, o.logl(te), rtol)
    close(got, GOLD["ref_hckde_logl_" + key], rtol)
    if dt == "float64":
        close(got, GOLD["scipy_hckde_logl_" + key], 1e-9)
    s = f.slogl(te)
    assert s == pytest.approx(o.slogl(te), rel=rtol)
    assert s == pytest.approx(float(np.sum(GOLD["ref_hckde_sums_" + key])), rel=rtol)
    # per-configuration factors are the CKDEs the reference would build
    for c, of in enumerate(o.factors):
        a = pbn.Assignment.from_index(c, o.discrete, [list(tr[v].cat.categories) for v in o.discrete], o.card, o.strides)
        cf = f.conditional_factor(a) if o.discrete else f._factors[0]
        assert (cf is None) == (of is None)
        if cf is not None:
            assert cf.num_instances() == of[1].shape[0]
            assert np.allclose(cf.kde_joint().bandwidth, of[2], rtol=1e-9 if dt == "float64" else 1e-4)


@pytest.mark.parametrize("dt,rtol", [("float64", RTOL64), ("float32", RTOL32)])
def test_hckde_nulls_unseen_and_singular_configurations(pbn, dt, rtol):
    tr, te = frames(500, 400, dt, seeds=(2, 3))
    tr, te = with_nulls(tr, 11), with_nulls(te, 12)
    # configuration (a2, b1) never seen in training; (a2, b3) has too few rows for a covariance
    tr = tr[~((tr.A == "a2") & (tr.B == "b1"))]
    few = tr[(tr.A == "a2") & (tr.B == "b3")].index[2:]
    tr = tr.drop(few).reset_index(drop=True)
    f = pbn.HCKDE("D", ["A", "C", "B"])
    f.fit(tr)
    o = ohy.HybridFactor("D", ["A", "C", "B"]).fit(tr)
    assert [x is None for x in f._factors] == [x is None for x in o.factors]
    assert sum(x is None for x in f._factors) == 2
    got, want = f.logl(te), o.logl(te)
    assert np.isnan(want).sum() > 50
    close(got, want, rtol)
    assert f.slogl(te) == pytest.approx(o.slogl(te), rel=rtol)
    assert "not fitted" in str(f)


def test_hckde_other_categories_and_types_are_rejected(pbn):
    tr, te = frames(300, 50, "float64")
    f = pbn.HCKDE("D", ["A", "C"])
    f.fit(tr)
    bad = te.copy()
    bad["A"] = bad["A"].cat.rename_categories({"a1": "zz"})
    with pytest.raises(ValueError, match="Category at index 0 is different for variable A"):
        f.logl(bad)
    bad = te.copy()
    bad["A"] = bad["C"]
    with pytest.raises(ValueError, match="is not categorical"):
        f.slogl(bad)
    bad = te.copy()
    bad["C"] = bad["A"]
    with pytest.raises(ValueError, match='must have "double" or "float" data type'):
        f.logl(bad)
    with pytest.raises(ValueError, match="Data type of training and test datasets is different"):
        f.logl(frames(300, 50, "float32")[1])
    g = pbn.HCKDE("D", ["A"])
    ints = tr.copy()
    ints["A"] = np.arange(len(tr))
    with pytest.raises(ValueError, match="Non valid data type for variable A"):
        g.fit(ints)


def test_hckde_without_discrete_evidence_is_a_ckde(pbn):
    tr, te = frames(400, 60, "float64")
    f = pbn.HCKDE("D", ["C"])
    f.fit(tr)
    c = pbn.CKDE("D", ["C"])
    c.fit(tr)
    assert np.array_equal(f.logl(te), c.logl(te)) and f.slogl(te) == c.slogl(te)


def test_hckde_selectors_per_assignment_and_python_selector(pbn):
    tr, te = frames(700, 90, "float64")

    class Halved(pbn.BandwidthSelector):
        def bandwidth(self, df, variables):
            return 0.5 * pbn.NormalReferenceRule().bandwidth(df, variables)

    f = pbn.HCKDE("D", ["A", "C"], {pbn.Assignment({"A": "a2"}): pbn.ScottsBandwidth()})
    f.fit(tr)
    g = pbn.HCKDE("D", ["A", "C"], Halved())
    g.fit(tr)
    for a, rule in (("a1", "normal_reference"), ("a2", "scott")):
        X = np.asfortranarray(tr[tr.A == a][["D", "C"]].to_numpy())
        H = oracle.bandwidth(X, rule)
        assert np.allclose(f.conditional_factor(pbn.Assignment({"A": a})).kde_joint().bandwidth, H, rtol=1e-10)
        T = np.asfortranarray(te[te.A == a][["D", "C"]].to_numpy())
        close(f.logl(te)[(te.A == a).to_numpy()], oracle.ckde_logl(X, T, H)[0], RTOL64)
        Hh = 0.5 * oracle.bandwidth(X, "normal_reference")
        close(g.logl(te)[(te.A == a).to_numpy()], oracle.ckde_logl(X, T, Hh)[0], RTOL64)


def test_hckde_pickle_round_trip(pbn):
    tr, te = frames(300, 70, "float64")
    f = pbn.HCKDE("D", ["B", "C"])
    f.fit(tr)
    g = pickle.loads(pickle.dumps(f))
    assert g.fitted() and g.evidence() == ["B", "C"]
    assert np.array_equal(g.logl(te), f.logl(te), equal_nan=True)
    assert g.slogl(te) == f.slogl(te)


@pytest.mark.parametrize("dt,rtol", [("float64", RTOL64), ("float32", RTOL32)])
@pytest.mark.parametrize("ev", [["A", "C", "B"], ["B"], ["C", "A"]])
def test_clg_vs_oracle(pbn, dt, rtol, ev):
    tr, te = frames(800, 150, dt)
    tr, te = with_nulls(tr, 21), with_nulls(te, 22)
    f = pbn.CLinearGaussianCPD("D", ev)
    f.fit(tr)
    o = ohy.HybridFactor("D", ev, kind="lg").fit(tr)
    close(f.logl(te), o.logl(te), max(rtol, 1e-9))
    assert f.slogl(te) == pytest.approx(o.slogl(te), rel=max(rtol, 1e-9))
    for cf, of in zip(f._factors, o.factors):
        assert (cf is None) == (of is None)
        if cf is not None:
            assert np.allclose(cf.beta, of[1], rtol=1e-7 if dt == "float64" else 1e-3, atol=1e-9 if dt == "float64" else 1e-4)
    g = pickle.loads(pickle.dumps(f))
    assert np.array_equal(g.logl(te), f.logl(te), equal_nan=True)


def test_clg_given_parameters(pbn):
    tr, te = frames(300, 60, "float64")
    f = pbn.CLinearGaussianCPD("D", ["A", "C"], {pbn.Assignment({"A": "a1"}): ([1.0, 2.0], 0.5)})
    f.fit(tr)
    a1 = f.conditional_factor(pbn.Assignment({"A": "a1"}))
    assert np.array_equal(a1.beta, [1.0, 2.0]) and a1.variance == 0.5
    from scipy.stats import norm
    q = (te.A == "a1").to_numpy()
    assert np.allclose(f.logl(te)[q], norm.logpdf(te.D[q], 1.0 + 2.0 * te.C[q], np.sqrt(0.5)), rtol=1e-10)


def test_table_take_and_multi_logl_through_the_abi(pbn):
    """pbn_table_take gathers exactly the requested rows; pbn_kde_logl_multi equals one pbn_kde_logl per job."""
    from pybnesian_b200 import hybrid as phy
    from pybnesian_b200._lib import Rows, check, int_array, lib
    df = util_data.generate_normal_data(3000, 0)
    frame = pbn.DataFrame(df)
    tbl, cols, _ = frame.device_table(["a", "b", "c"])
    rng = np.random.default_rng(0)
    order = rng.permutation(3000).astype(np.int32)[:2500]
    taken = phy._take_table(tbl, order)
    for c, name in zip(cols, ["a", "b", "c"]):
        assert np.array_equal(taken.download(c), df[name].to_numpy()[order])
    with pytest.raises(ValueError, match="take index out of range"):
        phy._take_table(tbl, np.array([3000], dtype=np.int32))
    # three CKDEs fitted on different row ranges, evaluated on three ranges of the gathered table
    ranges = [(0, 900), (900, 1000), (1000, 2500)]
    kdes = []
    for b, e in ranges:
        k = pbn.CKDE("c", ["a", "b"])
        H = pbn.NormalReferenceRule()._bandwidth_rows(taken, [cols[2], cols[0], cols[1]], Rows.single(b, e))
        k._fit_table(taken, [cols[2], cols[0], cols[1]], Rows.single(b, e), H)
        kdes.append(k)
    test_ranges = [(100, 700), (0, 0), (2000, 2500)]
    handles = (ctypes.c_void_p * 4)(kdes[0]._handle.handle, kdes[1]._handle.handle, kdes[2]._handle.handle, None)
    rows = (Rows * 4)(*[Rows.single(b, e) for b, e in test_ranges + [(700, 710)]])
    total = sum(e - b for b, e in test_ranges) + 10
    vals, sums = np.empty(total), np.zeros(4)
    dp = ctypes.POINTER(ctypes.c_double)
    ccols = int_array([cols[2], cols[0], cols[1]])
    check(lib().pbn_kde_logl_multi(tbl.ctx.handle, handles, 4, taken.handle, ccols, rows, vals.ctypes.data_as(dp),
                                   sums.ctypes.data_as(dp)))
    pos = 0
    for k, (b, e) in zip(kdes, test_ranges):
        one, s = np.empty(e - b), ctypes.c_double()
        check(lib().pbn_kde_logl(tbl.ctx.handle, k._handle.handle, taken.handle, ccols, Rows.single(b, e),
                                 one.ctypes.data_as(dp), ctypes.byref(s)))
        assert np.allclose(vals[pos:pos + e - b], one, rtol=1e-12, atol=0)
        assert sums[kdes.index(k)] == pytest.approx(s.value, rel=1e-12) or e == b
        pos += e - b
    assert np.all(np.isnan(vals[pos:])) and sums[3] == 0.0
    # the grouped evaluation costs a fixed number of launches, whatever the number of configurations
    before = tbl.ctx.counters()["launches"]
    check(lib().pbn_kde_logl_multi(tbl.ctx.handle, handles, 3, taken.handle, ccols, rows, None, sums.ctypes.data_as(dp)))
    assert tbl.ctx.counters()["launches"] - before <= 6  # fill, whiten, pair, finalize, fallback rows, per-job sums


def test_hybrid_semiparametric_bn_fit_logl(pbn):
    tr, te = frames(1500, 300, "float64")
    m = pbn.SemiparametricBN(["A", "B", "C", "D"], [("A", "B"), ("A", "D"), ("B", "D"), ("C", "D")],
                             [("D", pbn.CKDEType())])
    m.fit(tr)
    assert m.node_type("A") == pbn.DiscreteFactorType() or isinstance(m.cpd("A"), pbn.DiscreteFactor)
    assert isinstance(m.cpd("D"), pbn.HCKDE) and isinstance(m.cpd("C"), pbn.LinearGaussianCPD)
    assert isinstance(m.cpd("B"), pbn.DiscreteFactor)
    o = ohy.HybridFactor("D", m.cpd("D").evidence()).fit(tr)
    want = o.logl(te)
    lpA, _, sA = ohy.discrete_factor_logprob(tr, "A", [])
    lpB, _, sB = ohy.discrete_factor_logprob(tr, "B", ["A"])
    want = want + lpA[te.A.cat.codes.to_numpy()] + lpB[te.B.cat.codes.to_numpy() + 3 * te.A.cat.codes.to_numpy()]
    beta, var = oracle.lg_fit(tr.C.to_numpy(), [])
    want = want + oracle.lg_logl(te.C.to_numpy(), [], beta, var)[0]
    close(m.logl(te), want, 1e-9)
    assert m.slogl(te) == pytest.approx(np.sum(want), rel=1e-9)
    # a CLG node
    m2 = pbn.SemiparametricBN(["A", "B", "C", "D"], [("A", "D"), ("C", "D")])
    m2.fit(tr)
    assert isinstance(m2.cpd("D"), pbn.CLinearGaussianCPD)


def test_cv_likelihood_with_discrete_parents_vs_oracle(pbn):
    tr, _ = frames(1200, 10, "float64")
    score = pbn.CVLikelihood(tr, k=5, seed=0)
    m = pbn.SemiparametricBN(["A", "B", "C", "D"], [("A", "D"), ("C", "D")], [("D", pbn.CKDEType())])
    m.set_unknown_node_types(tr)
    got = score.local_score(m, "D")
    idx, limits = oracle.cv_indices(np.arange(len(tr), dtype=np.int32), 5, 0)
    want = 0.0
    ev = m.parents("D")
    for f in range(5):
        test_rows = idx[limits[f]:limits[f + 1]]
        train_rows = np.concatenate([idx[:limits[f]], idx[limits[f + 1]:]])
        o = ohy.HybridFactor("D", ev).fit(tr.iloc[train_rows].reset_index(drop=True))
        want += o.slogl(tr.iloc[test_rows].reset_index(drop=True))
    assert got == pytest.approx(want, rel=1e-10)
    # discrete node and CLG node scores
    m.set_node_type("D", pbn.LinearGaussianCPDType())
    want = 0.0
    for f in range(5):
        test_rows = idx[limits[f]:limits[f + 1]]
        train_rows = np.concatenate([idx[:limits[f]], idx[limits[f + 1]:]])
        o = ohy.HybridFactor("D", ev, kind="lg").fit(tr.iloc[train_rows].reset_index(drop=True))
        want += o.slogl(tr.iloc[test_rows].reset_index(drop=True))
    assert score.local_score(m, "D") == pytest.approx(want, rel=1e-9)
    assert np.isfinite(score.local_score(m, "B", ["A"]))


def test_hill_climbing_on_hybrid_data(pbn):
    tr, _ = frames(1000, 10, "float64")
    score = pbn.CVLikelihood(tr, k=4, seed=0)
    start = pbn.SemiparametricBN(list(tr.columns))
    pool = pbn.OperatorPool([pbn.ArcOperatorSet(), pbn.ChangeNodeTypeSet()])
    model = pbn.GreedyHillClimbing().estimate(pool, score, start, max_indegree=3)
    assert model.node_type("A") == pbn.DiscreteFactorType() and model.node_type("B") == pbn.DiscreteFactorType()
    for s, t in model.arcs():
        # a discrete node never gets a continuous parent (SemiparametricBN.hpp:96-101)
        assert not (model.node_type(t) == pbn.DiscreteFactorType() and model.node_type(s) != pbn.DiscreteFactorType())
    parents_d = set(model.parents("D"))
    assert {"A", "B"} <= parents_d  # D's law differs per (A, B) configuration
    # the greedy search can only have improved the score of the start model
    assert score.score(model) > score.score(start_typed(pbn, tr))


def start_typed(pbn, tr):
    m = pbn.SemiparametricBN(list(tr.columns))
    m.set_unknown_node_types(tr)
    return m


# ---- sampling on hybrid factors / networks ---------------------------------------------------
def test_hckde_sample_per_configuration_vs_oracle(pbn):
    """DiscreteAdaptator::sample (DiscreteAdaptator.hpp:426-520): configuration i is sampled by its CKDE with
    seed + i on its rows, every factor being asked for n draws; replayed here with the oracle's CKDE sampler."""
    tr, _ = frames(1500, 10, "float64")
    f = pbn.HCKDE("D", ["A", "C", "B"])
    f.fit(tr)
    n, seed = 400, 21
    ev = util_data.generate_hybrid_data(n, 5)[["A", "C", "B"]]
    s = f.sample(n, ev, seed)
    assert s.type == pa.float64() and len(s) == n
    got = s.to_numpy()
    cfg = ev["A"].cat.codes.to_numpy() + 2 * ev["B"].cat.codes.to_numpy()
    for i in range(6):
        rows = np.flatnonzero(cfg == i)
        if rows.size == 0:
            continue
        sel = (tr["A"].cat.codes.to_numpy() + 2 * tr["B"].cat.codes.to_numpy()) == i
        X = tr.loc[sel, ["D", "C"]].to_numpy()
        evc = ev["C"].to_numpy()[np.concatenate([rows, np.full(n - rows.size, rows[0])])].reshape(-1, 1)
        want, _ = oracle.ckde_sample(X, oracle.bandwidth(X), evc, n, seed + i)
        close(got[rows], want[:rows.size], 1e-11)
    with pytest.raises(ValueError, match="rows to sample"):
        f.sample(n + 1, ev, 0)
    with pytest.raises(ValueError, match="non-negative"):
        f.sample(-1, ev, 0)


def test_clg_sample_and_hybrid_network_sample(pbn):
    tr, _ = frames(1500, 10, "float64")
    g = pbn.CLinearGaussianCPD("D", ["A", "C"])
    g.fit(tr)
    ev = util_data.generate_hybrid_data(300, 6)[["A", "C"]]
    got = g.sample(300, ev, 9).to_numpy()
    codes = ev["A"].cat.codes.to_numpy()
    for i in range(2):
        rows = np.flatnonzero(codes == i)
        base = g.conditional_factor(pbn.Assignment({"A": "a%d" % (i + 1)}))
        evc = ev["C"].to_numpy()[np.concatenate([rows, np.full(300 - rows.size, rows[0])])]
        want = oracle.lg_sample(base.beta, base.variance, [evc], 300, 9 + i)
        assert np.array_equal(got[rows], want[:rows.size])
    # ancestral sampling through discrete, conditional-linear-Gaussian and hybrid-CKDE nodes
    m = pbn.SemiparametricBN(["A", "B", "C", "D"], [("A", "B"), ("A", "D"), ("B", "D"), ("C", "D")],
                             [("D", pbn.CKDEType())])
    m.fit(tr)
    s = m.sample(2000, 3, ordered=True).to_pandas()
    assert list(s.columns) == ["A", "B", "C", "D"] and len(s) == 2000
    assert str(s["A"].dtype) == "category" and set(s["B"].cat.categories) == {"b1", "b2", "b3"}
    assert abs(np.mean(s["A"] == "a1") - 0.75) < 0.05
    assert np.isfinite(s["D"]).all()
    # the law of D given (a1, b2) is -2 + C + N(0, 2): compare the sampled conditional mean with the training one
    sel_s = (s["A"] == "a1") & (s["B"] == "b2")
    sel_t = (tr["A"] == "a1") & (tr["B"] == "b2")
    assert abs(s.loc[sel_s, "D"].mean() - tr.loc[sel_t, "D"].mean()) < 0.5
    assert s.equals(m.sample(2000, 3, ordered=True).to_pandas())


def test_heterogeneous_bn_fit_and_default_types(pbn):
    """A HeterogeneousBN whose continuous default is CKDE: fit() gives every node the first default type of its data
    type (BayesianNetwork.hpp:965-976), HCKDE / DiscreteFactor included, and equals the SemiparametricBN with the same
    node types."""
    tr, te = frames(1200, 200, "float64")
    types = {pa.float64(): [pbn.CKDEType(), pbn.LinearGaussianCPDType()],
             pa.dictionary(pa.int8(), pa.string()): [pbn.DiscreteFactorType()]}
    arcs = [("A", "B"), ("A", "D"), ("C", "D")]
    m = pbn.HeterogeneousBN(types, ["A", "B", "C", "D"], arcs)
    m.fit(tr)
    assert m.node_type("C") == pbn.CKDEType() and m.node_type("A") == pbn.DiscreteFactorType()
    assert isinstance(m.cpd("D"), pbn.HCKDE) and isinstance(m.cpd("C"), pbn.CKDE)
    s = pbn.SemiparametricBN(["A", "B", "C", "D"], arcs, [("C", pbn.CKDEType()), ("D", pbn.CKDEType())])
    s.fit(tr)
    close(m.logl(te), s.logl(te), 1e-12)
    assert len(m.sample(50, 1)) == 50
