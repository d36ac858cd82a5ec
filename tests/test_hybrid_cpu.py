"""CPU tests (no GPU) for SURVEY §8 row f1 (hybrid factors): the oracle's DiscreteAdaptator restatement against
the committed golden vectors of the reference's own kernels and against SciPy; the bit-exact host integer logic
behind the C ABI (pbn_discrete_slices); DiscreteFactor (host only) and the Assignment / network-type mirrors."""
import os
import pickle

import numpy as np
import pandas as pd
import pyarrow as pa
import pytest

import oracle
from oracle import hybrid as ohy
import pybnesian_b200 as pbn
from pybnesian_b200 import hybrid as phy
import util_data

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "hybrid_golden.npz"))
EVIDENCE = [["A", "C", "B"], ["A"], ["B", "A"], ["C", "B"]]
CASES = [(600, 80), (300, 120)]


def frames(N, m, dt):
    tr, te = util_data.generate_hybrid_data(N, 0), util_data.generate_hybrid_data(m, 1)
    for df in (tr, te):
        df["C"] = df["C"].astype(dt)
        df["D"] = df["D"].astype(dt)
    return tr, te


@pytest.mark.parametrize("dt", ["float64", "float32"])
@pytest.mark.parametrize("ev", EVIDENCE)
@pytest.mark.parametrize("N,m", CASES)
def test_oracle_hckde_bit_exact_with_reference_kernels_golden(dt, ev, N, m):
    tr, te = frames(N, m, dt)
    key = "%s_%s_%d_%d" % (dt, "".join(ev), N, m)
    f = ohy.HybridFactor("D", ev).fit(tr)
    assert np.array_equal(f.logl(te), GOLD["ref_hckde_logl_" + key], equal_nan=True)
    sums = GOLD["ref_hckde_sums_" + key]
    total = 0.0
    for s in sums:
        total += s
    assert f.slogl(te) == total


@pytest.mark.parametrize("ev", EVIDENCE)
@pytest.mark.parametrize("N,m", CASES)
def test_oracle_hckde_vs_scipy_per_configuration(ev, N, m):
    tr, te = frames(N, m, "float64")
    key = "float64_%s_%d_%d" % ("".join(ev), N, m)
    f = ohy.HybridFactor("D", ev).fit(tr)
    assert np.allclose(f.logl(te), GOLD["scipy_hckde_logl_" + key], rtol=1e-9, atol=1e-12, equal_nan=True)


def test_oracle_clg_vs_lstsq_per_configuration():
    """The reference's LinearGaussianCPD tests fit with lstsq and score with norm.logpdf (cvlikelihood_test.py:12-49)."""
    from scipy.stats import norm
    tr, te = frames(800, 150, "float64")
    f = ohy.HybridFactor("D", ["A", "C", "B"], kind="lg").fit(tr)
    got = f.logl(te)
    want = np.full(len(te), np.nan)
    for a in ("a1", "a2"):
        for b in ("b1", "b2", "b3"):
            s = tr[(tr.A == a) & (tr.B == b)]
            X = np.column_stack([np.ones(len(s)), s.C])
            beta, res, _, _ = np.linalg.lstsq(X, s.D, rcond=None)
            var = res[0] / (len(s) - 2)
            q = (te.A == a) & (te.B == b)
            want[q.to_numpy()] = norm.logpdf(te.D[q], beta[0] + beta[1] * te.C[q], np.sqrt(var))
    assert np.allclose(got, want, rtol=1e-9, atol=0)
    assert f.slogl(te) == pytest.approx(np.nansum(want), rel=1e-12)


def _with_nulls(df, seed):
    rng = np.random.default_rng(seed)
    df = df.copy()
    n = len(df)
    df.loc[rng.choice(n, n // 15, replace=False), "A"] = np.nan
    df.loc[rng.choice(n, n // 20, replace=False), "B"] = np.nan
    df.loc[rng.choice(n, n // 25, replace=False), "C"] = np.nan
    return df


@pytest.mark.parametrize("nulls", [False, True])
@pytest.mark.parametrize("discrete", [["A"], ["B"], ["A", "B"], ["B", "A"]])
def test_discrete_slices_abi_bit_exact_with_restatement(discrete, nulls):
    """pbn_discrete_slices (host, integer) == discrete_slice_indices as written in the reference."""
    df = util_data.generate_hybrid_data(1500, 3)
    if nulls:
        df = _with_nulls(df, 5)
    frame = pbn.DataFrame(df)
    card, strides = phy.create_cardinality_strides(frame, discrete)
    ocard, ostrides = ohy.cardinality_strides(df, discrete)
    assert card.tolist() == ocard and strides.tolist() == ostrides
    F = int(card.prod())
    order, offsets = phy.discrete_slices(frame, discrete, strides, F)
    want = ohy.slice_indices(df, discrete, ostrides, F)
    assert offsets.tolist() == np.concatenate([[0], np.cumsum([len(w) for w in want])]).tolist()
    for c in range(F):
        assert order[offsets[c]:offsets[c + 1]].tolist() == want[c]
    assert order.dtype == np.int32


def test_discrete_slices_rejects_out_of_range_configuration():
    df = util_data.generate_hybrid_data(50, 0)
    frame = pbn.DataFrame(df)
    with pytest.raises(ValueError, match="out of range"):
        phy.discrete_slices(frame, ["A", "B"], np.array([1, 2], dtype=np.int32), 3)


def test_discrete_slices_empty_frame_and_int8_codes():
    df = util_data.generate_hybrid_data(40, 0)
    rb = pa.RecordBatch.from_pandas(df, None, False)
    assert rb.schema.field("A").type.index_type == pa.int8()
    frame = pbn.DataFrame(rb.slice(0, 0))
    order, offsets = phy.discrete_slices(frame, ["A", "B"], np.array([1, 2], dtype=np.int32), 6)
    assert order.size == 0 and offsets.tolist() == [0] * 7


@pytest.mark.parametrize("nulls", [False, True])
@pytest.mark.parametrize("evidence", [[], ["A"], ["B", "A"]])
def test_discrete_factor_host_vs_oracle(evidence, nulls):
    tr, te = util_data.generate_hybrid_data(900, 0), util_data.generate_hybrid_data(200, 1)
    if nulls:
        tr, te = _with_nulls(tr, 1), _with_nulls(te, 2)
    var = "B" if "B" not in evidence else "A"
    evidence = [e for e in evidence if e != var]
    f = pbn.DiscreteFactor(var, evidence)
    with pytest.raises(ValueError, match="not fitted"):
        f.logl(te)
    f.fit(tr)
    logprob, card, strides = ohy.discrete_factor_logprob(tr, var, evidence)
    assert np.array_equal(f._logprob, logprob)
    codes = [te[v].cat.codes.to_numpy() for v in [var] + evidence]
    ok = np.all([c >= 0 for c in codes], axis=0)
    idx = sum(c.astype(np.int64) * s for c, s in zip(codes, strides))
    want = np.where(ok, logprob[np.where(ok, idx, 0)], np.nan)
    assert np.array_equal(f.logl(te), want, equal_nan=True)
    total = 0.0
    for v in want[ok]:
        total += v
    assert f.slogl(te) == total
    g = pickle.loads(pickle.dumps(f))
    assert np.array_equal(g.logl(te), want, equal_nan=True)
    assert f.type() == pbn.DiscreteFactorType() and str(f.type()) == "DiscreteFactor"


def test_discrete_factor_rejects_other_categories_and_continuous_data():
    tr = util_data.generate_hybrid_data(100, 0)
    f = pbn.DiscreteFactor("A", ["B"])
    f.fit(tr)
    other = tr.copy()
    other["B"] = other["B"].cat.rename_categories({"b1": "zz"})
    with pytest.raises(ValueError, match="Category at index 0 is different for variable B"):
        f.logl(other)
    other["B"] = tr["B"].cat.add_categories(["b4"])
    with pytest.raises(ValueError, match="does not contain the same categories"):
        f.slogl(other)
    with pytest.raises(ValueError, match="Categorical data is expected"):
        pbn.DiscreteFactor("C", []).fit(tr)


def test_assignment_mirror():
    a = pbn.Assignment({"A": "a1", "B": "b2"})
    assert a.value("A") == "a1" and a.size() == 2 and not a.empty() and a.has_variables(["A", "B"])
    assert not a.has_variables(["Z"])
    with pytest.raises(ValueError, match="not found in the assignment"):
        a.value("Z")
    assert a == pbn.Assignment({"B": "b2", "A": "a1"}) and hash(a) == hash(pbn.Assignment({"B": "b2", "A": "a1"}))
    assert a != pbn.Assignment({"A": "a1"})
    vals = [["a1", "a2"], ["b1", "b2", "b3"]]
    assert a.index(["A", "B"], vals, [1, 2]) == 2
    for i in range(6):
        b = pbn.Assignment.from_index(i, ["A", "B"], vals, [2, 3], [1, 2])
        assert b.index(["A", "B"], vals, [1, 2]) == i
    with pytest.raises(ValueError, match="is not valid for variable"):
        pbn.Assignment({"A": "zz", "B": "b1"}).index(["A", "B"], vals, [1, 2])
    a.insert("C", 1)
    assert a.value("C") == 1.0
    a.remove("C")
    assert pickle.loads(pickle.dumps(a)) == a
    assert str(pbn.Assignment()) == "[]"


def test_semiparametric_type_with_discrete_nodes():
    """models/SemiparametricBN.hpp:43-101."""
    t = pbn.SemiparametricBNType()
    dict_t = pa.dictionary(pa.int8(), pa.string())
    assert t.data_default_node_type(dict_t) == [pbn.DiscreteFactorType()]
    assert t.data_default_node_type(pa.float64()) == [pbn.LinearGaussianCPDType(), pbn.CKDEType()]
    with pytest.raises(ValueError, match="not compatible with SemiparametricBNType"):
        t.data_default_node_type(pa.int32())
    m = pbn.SemiparametricBN(["A", "B", "C", "D"], [("A", pbn.DiscreteFactorType()), ("B", pbn.DiscreteFactorType()),
                                                     ("C", pbn.LinearGaussianCPDType()), ("D", pbn.CKDEType())])
    assert m.can_add_arc("A", "B") and m.can_add_arc("A", "D") and m.can_add_arc("C", "D")
    assert not m.can_add_arc("C", "A") and not m.can_add_arc("D", "B")
    m.add_arc("A", "D")
    assert not m.can_flip_arc("A", "D")
    m.add_arc("C", "D")
    with pytest.raises(ValueError, match="Wrong factor type"):
        m.set_node_type("D", pbn.DiscreteFactorType())  # a discrete node cannot keep the continuous parent C ...
    m2 = pbn.SemiparametricBN(["A", "C"], [("A", pbn.DiscreteFactorType()), ("C", pbn.CKDEType())])
    m2.add_arc("A", "C")
    m2.set_node_type("C", pbn.DiscreteFactorType())  # ... when every parent is discrete
    assert m2.node_type("C") == pbn.DiscreteFactorType()
    assert t.alternative_node_type(m, "A") == []
    # a discrete parent turns the continuous factor types into their hybrid adaptors (CKDE.cpp:15-41,
    # LinearGaussianCPD.cpp:33-57)
    assert isinstance(pbn.CKDEType().new_factor(m, "D", ["A", "C"]), pbn.HCKDE)
    assert isinstance(pbn.LinearGaussianCPDType().new_factor(m, "C", ["B"]), pbn.CLinearGaussianCPD)
    assert type(pbn.CKDEType().new_factor(m, "D", ["C"])) is pbn.CKDE
    assert pbn.HCKDE("D", ["A"]).type() == pbn.CKDEType()
    assert pbn.CLinearGaussianCPD("D", ["A"]).type() == pbn.LinearGaussianCPDType()


def test_hybrid_factor_constructor_and_unfitted_errors():
    f = pbn.HCKDE("D", ["A", "C"])
    assert f.variable() == "D" and f.evidence() == ["A", "C"] and not f.fitted()
    assert str(f) == "[HCKDE] P(D | A, C) not fitted."
    with pytest.raises(ValueError, match="not fitted"):
        f.logl(util_data.generate_hybrid_data(10, 0))
    with pytest.raises(RuntimeError, match="Bandwidth selector procedure must be non-null"):
        pbn.HCKDE("D", ["A"], None)
    g = pbn.HCKDE("D", ["A"], {pbn.Assignment({"A": "a1"}): pbn.ScottsBandwidth()})
    assert g._specific and isinstance(g._initialize(pbn.Assignment({"A": "a1"})).bandwidth_type(), pbn.ScottsBandwidth)
    assert isinstance(g._initialize(pbn.Assignment({"A": "a2"})).bandwidth_type(), pbn.NormalReferenceRule)
    h = pickle.loads(pickle.dumps(g))
    assert h._specific and h.evidence() == ["A"]
    c = pbn.CLinearGaussianCPD("D", ["A", "C"], [1.0, 2.0], 0.5)
    c._continuous_evidence = ["C"]  # what fit() derives from the data types
    assert c._initialize(pbn.Assignment()).fitted()


# ---- sampling (SURVEY §8 f3 on hybrid networks): DiscreteFactor is host only ------------------------------
def _discrete_sample_restated(logprob, c0, parent_rows, n, seed):
    """DiscreteFactor::sample_indices (DiscreteFactor.hpp:144-207), line by line."""
    u = oracle.uniform_real(n, seed, np.float64)
    out = np.empty(n, dtype=np.int64)
    for i in range(n):
        off = int(parent_rows[i]) * c0
        acc, index = 0.0, c0 - 1
        for j in range(c0 - 1):
            acc = np.exp(logprob[off + j]) if j == 0 else acc + np.exp(logprob[off + j])
            if u[i] < acc:
                index = j
                break
        out[i] = index
    return out


def test_discrete_factor_sample():
    tr = util_data.generate_hybrid_data(2000, 0)
    f = pbn.DiscreteFactor("B", ["A"])
    f.fit(tr)
    ev = util_data.generate_hybrid_data(500, 3)[["A"]]
    s = f.sample(500, ev, 17)
    assert pa.types.is_dictionary(s.type) and s.type == f.data_type() and len(s) == 500
    assert s.dictionary.to_pylist() == ["b1", "b2", "b3"]
    parent = ev["A"].cat.codes.to_numpy()
    want = _discrete_sample_restated(f._logprob, 3, parent, 500, 17)
    assert np.array_equal(s.indices.to_numpy(), want)
    # without evidence; frequencies follow the fitted table
    g = pbn.DiscreteFactor("A", [])
    g.fit(tr)
    s = g.sample(4000, None, 1)
    assert np.array_equal(s.indices.to_numpy(), _discrete_sample_restated(g._logprob, 2, np.zeros(4000, dtype=int), 4000, 1))
    assert abs(np.mean(s.indices.to_numpy() == 0) - np.exp(g._logprob[0])) < 0.03
    with pytest.raises(ValueError, match="non-negative"):
        f.sample(-1, ev, 0)
    with pytest.raises(ValueError, match="rows to sample"):
        f.sample(10, ev, 0)
    with pytest.raises(ValueError, match="not present"):
        f.sample(500, None, 0)
    bad = ev.copy()
    bad["A"] = bad["A"].cat.rename_categories({"a1": "zz"})
    with pytest.raises(ValueError):
        f.sample(500, bad, 0)


def test_pbn_uniform_real_is_the_libstdcxx_stream():
    import ctypes
    from pybnesian_b200 import _lib
    for dt, code in ((np.float64, 0), (np.float32, 1)):
        out = np.empty(33, dtype=dt)
        assert _lib.lib().pbn_uniform_real(33, 7, code, out.ctypes.data_as(ctypes.c_void_p)) == 0
        assert np.array_equal(out, oracle.uniform_real(33, 7, dt))


def test_mle_discrete_factor():
    tr = util_data.generate_hybrid_data(2000, 0)
    p = pbn.MLE(pbn.DiscreteFactorType()).estimate(tr, "B", ["A"])
    assert p.logprob.shape == (3, 2)
    assert np.allclose(np.exp(p.logprob).sum(axis=0), 1.0)
    tab = pd.crosstab(tr["B"], tr["A"], normalize="columns").to_numpy()
    assert np.allclose(np.exp(p.logprob), tab, rtol=1e-12)


def test_heterogeneous_bn_type_equality_like_the_reference_test():
    """tests/models/HeterogeneousBN_test.py:5-55."""
    nodes = ["a", "b", "c", "d"]
    het_single = pbn.HeterogeneousBN([pbn.CKDEType(), pbn.LinearGaussianCPDType()], nodes)
    het2_single = pbn.HeterogeneousBN([pbn.CKDEType(), pbn.LinearGaussianCPDType()], nodes)
    assert het_single.type() == het2_single.type()
    het3_single = pbn.HeterogeneousBN([pbn.LinearGaussianCPDType(), pbn.CKDEType()], nodes)
    assert het_single.type() != het3_single.type()
    cont = [pbn.CKDEType(), pbn.LinearGaussianCPDType()]
    het_dt = pbn.HeterogeneousBN({pa.float64(): cont, pa.float32(): cont,
                                  pa.dictionary(pa.int8(), pa.string()): [pbn.DiscreteFactorType()]}, nodes)
    het2_dt = pbn.HeterogeneousBN({pa.dictionary(pa.int8(), pa.string()): [pbn.DiscreteFactorType()],
                                   pa.float32(): cont, pa.float64(): cont}, nodes)
    assert het_dt.type() == het2_dt.type()  # the order of the map is not relevant
    het3_dt = pbn.HeterogeneousBN({pa.dictionary(pa.int8(), pa.string()): [pbn.DiscreteFactorType()],
                                   pa.float32(): [pbn.LinearGaussianCPDType(), pbn.CKDEType()], pa.float64(): cont}, nodes)
    assert het_dt.type() != het3_dt.type()  # the order of the default FactorTypes is relevant
    assert het_single.type() != pbn.HeterogeneousBN({pa.float64(): cont}, nodes).type()
    # the rest of the type's surface
    t = het_dt.type()
    assert not t.is_homogeneous() and not t.single_default() and het_single.type().single_default()
    assert t.data_default_node_type(pa.dictionary(pa.int32(), pa.string())) == [pbn.DiscreteFactorType()]  # by type id
    assert t.data_default_node_type(pa.float32()) == cont
    with pytest.raises(ValueError, match="Not valid FactorType"):
        pbn.HeterogeneousBNType({pa.float64(): cont}).data_default_node_type(pa.float32())
    with pytest.raises(ValueError, match="cannot be empty"):
        pbn.HeterogeneousBNType([])
    with pytest.raises(RuntimeError):
        t.default_node_type()
    r = pickle.loads(pickle.dumps(het_dt))
    assert type(r) is pbn.HeterogeneousBN and r.type() == t and r.nodes() == nodes
    assert type(t.new_bn(["x", "y"])) is pbn.HeterogeneousBN
    assert all(het_dt.node_type(n) == pbn.UnknownFactorType() for n in nodes)
    m = pbn.HeterogeneousBN(cont, nodes, [("a", "b")], [("a", pbn.LinearGaussianCPDType())])
    assert m.has_arc("a", "b") and m.node_type("a") == pbn.LinearGaussianCPDType()
