"""CPU tests (no GPU) of the host-side logic around the path: fold generation (bit-exact, libstdc++
known answers), the CrossValidation / HoldOut mirrors (same assertions as the reference's
tests/dataset/crossvalidation_test.py and holdout_test.py), the DAG, and the hill-climbing operator
logic - run against a CPU score built on the oracle and compared with the oracle's own serial search."""
import ctypes

import numpy as np
import pandas as pd
import pytest

import oracle
from oracle import hc as oracle_hc
import util_data
import pybnesian_b200 as pbn
from pybnesian_b200 import _lib

SIZE = 2000
df = util_data.generate_normal_data(SIZE)


def _split(n, k, seed):
    idx = np.arange(n, dtype=np.int32)
    lim = np.empty(k + 1, dtype=np.int32)
    p = ctypes.POINTER(ctypes.c_int32)
    _lib.check(_lib.lib().pbn_cv_split(idx.ctypes.data_as(p), n, k, seed, lim.ctypes.data_as(p)))
    return idx, lim


def test_cv_split_libstdcxx_known_answers():
    # SURVEY.md §7 "Bit-exact folds": std::shuffle(std::mt19937{0}) of iota(n), recorded from libstdc++ (g++ 13)
    assert _split(10, 2, 0)[0].tolist() == [0, 2, 1, 5, 9, 8, 4, 7, 6, 3]
    assert _split(23, 2, 0)[0].tolist() == [10, 4, 6, 12, 22, 1, 9, 17, 7, 20, 11, 18, 19, 15, 21, 3, 14, 13, 2, 0, 16, 8, 5]
    assert _split(1000, 10, 0)[0][:10].tolist() == [882, 396, 136, 545, 569, 298, 709, 664, 519, 504]
    # fold limits: the first n % k folds get one more row
    assert _split(23, 5, 0)[1].tolist() == [0, 5, 10, 15, 19, 23]


@pytest.mark.parametrize("n,k,seed", [(1000, 10, 0), (1003, 7, 123), (17, 17, 5)])
def test_cv_split_matches_oracle(n, k, seed):
    idx, lim = _split(n, k, seed)
    oidx, olim = oracle.cv_indices(np.arange(n), k, seed)
    assert np.array_equal(idx, oidx) and np.array_equal(lim, olim)


def test_cv_disjoint_indices_and_folds():
    cv = pbn.CrossValidation(df)
    for i, ((train_df, test_df), (train_indices, test_indices)) in enumerate(zip(cv, cv.indices())):
        nptrain, nptest = np.asarray(train_indices), np.asarray(test_indices)
        assert np.all(np.sort(np.hstack((nptrain, nptest))) == np.arange(SIZE))
        assert np.all(train_df.to_pandas().to_numpy() == df.iloc[train_indices, :].to_numpy())
        assert np.all(test_df.to_pandas().to_numpy() == df.iloc[test_indices, :].to_numpy())
        assert np.setdiff1d(nptrain, nptest).shape == nptrain.shape
        train_fold, test_fold = cv.fold(i)
        assert train_fold.equals(train_df) and test_fold.equals(test_df)


def test_cv_seed_and_num_folds():
    a, b, c = list(pbn.CrossValidation(df, seed=0)), list(pbn.CrossValidation(df, seed=0)), list(pbn.CrossValidation(df, seed=1))
    for (tr, te), (tr2, te2), (tr3, te3) in zip(a, b, c):
        assert tr.equals(tr2) and te.equals(te2)
        assert not tr.equals(tr3) and not te.equals(te3)
    assert len(a) == 10 and len(list(pbn.CrossValidation(df, 5))) == 5
    with pytest.raises(ValueError, match="Cannot split"):
        pbn.CrossValidation(df, SIZE + 1)
    with pytest.raises(ValueError, match="Cannot split"):
        pbn.CrossValidation(df, 1)


def test_cv_loc():
    cv = pbn.CrossValidation(df)
    for sel, names in (("a", ["a"]), (1, ["b"]), (["b", "d"], ["b", "d"]), ([0, 2], ["a", "c"])):
        for train_df, test_df in cv.loc(sel):
            assert train_df.schema.names == names and test_df.schema.names == names


def test_cv_null():
    np.random.seed(0)
    nulls = {c: np.random.randint(0, SIZE, size=100) for c in "abcd"}
    df_null = df.copy()
    for c, rows in nulls.items():
        df_null.loc[df_null.index[rows], c] = np.nan
    valid = np.setdiff1d(np.arange(SIZE), np.concatenate(list(nulls.values())))
    cv = pbn.CrossValidation(df_null, seed=3)
    for (train_df, test_df), (tri, tei) in zip(cv, cv.indices()):
        assert train_df.num_rows + test_df.num_rows == valid.size
        assert np.all(np.sort(np.hstack((tri, tei))) == valid)
        assert np.all(train_df.to_pandas().to_numpy() == df_null.iloc[tri, :].to_numpy())
    # identical to the oracle's shuffle of the valid rows
    oidx, olim = oracle.cv_indices(valid, 10, 3)
    assert np.array_equal(cv._prop.indices, oidx) and np.array_equal(cv._prop.limits, olim)
    cv_all = pbn.CrossValidation(df_null, include_null=True)
    for train_df, test_df in cv_all:
        assert train_df.num_rows + test_df.num_rows == SIZE


def test_holdout():
    h = pbn.HoldOut(df, 0.2, 0)
    assert h.training_data().num_rows == 1600 and h.test_data().num_rows == 400
    tr, te = oracle.holdout_indices(np.arange(SIZE), 0.2, 0)
    assert np.array_equal(h.train_indices, tr) and np.array_equal(h.test_indices, te)
    assert np.all(h.training_data().to_pandas().to_numpy() == df.iloc[tr, :].to_numpy())
    h3 = pbn.HoldOut(df, 0.3, 1)
    assert h3.test_data().num_rows == round(SIZE * 0.3)
    for bad in (0, 1, -0.5, 1.5):
        with pytest.raises(ValueError, match="test_ratio must be a number"):
            pbn.HoldOut(df, bad)
    df_null = df.copy()
    df_null.loc[df_null.index[:50], "a"] = np.nan
    hn = pbn.HoldOut(df_null, 0.2, 0)
    assert hn.training_data().num_rows + hn.test_data().num_rows == SIZE - 50


def test_sort_desc_is_std_sort():
    rng = np.random.default_rng(0)
    delta = rng.normal(size=400)
    delta[rng.integers(0, 400, 80)] = 1.5   # ties: the unstable std::sort decides
    idx = np.arange(400, dtype=np.int32)
    want = idx.copy()
    L = oracle.lib()
    L.orc_sort_desc.argtypes = [ctypes.POINTER(ctypes.c_int), ctypes.c_int64, ctypes.POINTER(ctypes.c_double)]
    dp = delta.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    for _ in range(3):  # persistent vector re-sorted, as find_max does
        _lib.check(_lib.lib().pbn_sort_desc(idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), idx.size, dp))
        L.orc_sort_desc(want.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), want.size, dp)
        assert np.array_equal(idx, want)
        assert np.all(np.diff(delta[idx]) <= 0)
        delta[rng.integers(0, 400, 30)] += 0.25


def test_dag_parent_order_and_acyclicity():
    g = pbn.Dag(["a", "b", "c", "d", "e"])
    for s, t in (("a", "e"), ("b", "e"), ("c", "e"), ("d", "e")):
        g.add_arc(s, t)
    # libstdc++ unordered_set<int>: most recently inserted first
    assert g.parents("e") == ["d", "c", "b", "a"]
    c = g.clone()
    c.remove_arc("c", "e")
    assert c.parents("e") == ["d", "b", "a"] and g.parents("e") == ["d", "c", "b", "a"]
    g2 = pbn.Dag(["a", "b", "c", "d"], [("a", "b"), ("b", "c"), ("a", "c")])
    assert not g2.can_add_arc("c", "a") and g2.can_add_arc("c", "d") and g2.can_add_arc("d", "a")
    assert g2.can_flip_arc("a", "b") is False or g2.can_flip_arc("a", "b") is True
    assert not g2.can_flip_arc("a", "c")      # a -> b -> c would close a cycle
    assert g2.can_flip_arc("b", "c")
    with pytest.raises(ValueError):
        g2.add_arc("c", "a")
    # brute force: can_add_arc(s, t) <=> adding keeps the graph acyclic
    rng = np.random.default_rng(1)
    names = ["n%d" % i for i in range(7)]
    g3 = pbn.Dag(names)
    for _ in range(60):
        s, t = rng.choice(names, 2, replace=False)
        if g3.has_arc(s, t):
            continue
        ok = g3.can_add_arc(s, t)
        assert ok == (not g3.has_path(t, s))
        if ok and rng.random() < 0.6:
            g3.add_arc(s, t)
            g3.topological_sort()


class OracleCVScore(pbn.Score):
    """CVLikelihood computed on the CPU by the oracle: exercises the product's operator / hill-climbing host
    logic without a GPU."""

    def __init__(self, frame, k, seed):
        self._df = frame
        self._X = frame.to_numpy()
        self._cols = {c: i for i, c in enumerate(frame.columns)}
        self._idx, self._lim = oracle.cv_indices(np.arange(len(frame)), k, seed)
        self.calls = 0

    def data(self):
        return pbn.DataFrame(self._df)

    def local_score(self, model, variable, evidence=None):
        evidence = model.parents(variable) if evidence is None else evidence
        return self.local_score_node_type(model, model.underlying_node_type(self._df, variable), variable, evidence)

    def local_score_node_type(self, model, variable_type, variable, evidence):
        self.calls += 1
        cols = [self._cols[variable]] + [self._cols[e] for e in evidence]
        X = np.asfortranarray(self._X[:, cols])
        return oracle.cv_score(X, self._idx, self._lim, "ckde" if variable_type == pbn.CKDEType() else "lg")


def nonlinear_data(rows, seed=0):
    rng = np.random.default_rng(seed)
    a = rng.normal(0, 1, rows)
    b = 0.8 * a + rng.normal(0, 0.6, rows)
    c = np.sin(2.0 * a) + rng.normal(0, 0.25, rows)
    d = 0.5 * b * b + rng.normal(0, 0.4, rows)
    e = -1.1 * c + 0.7 * d + rng.normal(0, 0.5, rows)
    return pd.DataFrame({"a": a, "b": b, "c": c, "d": d, "e": e})


def _as_tuples(model_ops, names):
    out = []
    for op in model_ops:
        if isinstance(op, pbn.ChangeNodeType):
            out.append(("ChangeNodeType", names.index(op.node()), str(op.node_type()), op.delta()))
        else:
            out.append((type(op).__name__, names.index(op.source()), names.index(op.target()), op.delta()))
    return out


@pytest.mark.parametrize("max_indegree", [0, 2])
def test_hill_climbing_host_logic_matches_oracle(max_indegree):
    data = nonlinear_data(300, 0)
    names = list(data.columns)
    want_ops, want_arcs, want_types, _ = oracle_hc.hill_climb(data.to_numpy(), k=5, seed=0, max_indegree=max_indegree)
    assert any(o[0] == "ChangeNodeType" for o in want_ops), "test data should trigger a node-type change"
    score = OracleCVScore(data, 5, 0)
    pool = pbn.OperatorPool([pbn.ArcOperatorSet(), pbn.ChangeNodeTypeSet()])
    ghc = pbn.GreedyHillClimbing()
    best = ghc.estimate(pool, score, pbn.SemiparametricBN(names), max_indegree=max_indegree)
    got_ops = _as_tuples(ghc.last_run["operators"], names)
    assert [o[:3] for o in got_ops] == [o[:3] for o in want_ops]
    assert np.allclose([o[3] for o in got_ops], [o[3] for o in want_ops], rtol=1e-12, atol=0)
    assert sorted((names.index(s), names.index(t)) for s, t in best.arcs()) == want_arcs
    assert [str(best.node_type(n)) for n in names] == want_types


def test_hill_climbing_callback_blacklist_whitelist():
    data = nonlinear_data(200, 1)
    names = list(data.columns)
    score = OracleCVScore(data, 5, 0)

    class Rec(pbn.Callback):
        def __init__(self):
            self.seen = []

        def call(self, model, op, score, iteration):
            self.seen.append((iteration, None if op is None else str(op)))

    rec = Rec()
    best = pbn.GreedyHillClimbing().estimate(pbn.ArcOperatorSet(), score, pbn.SemiparametricBN(names),
                                             arc_blacklist=[("a", "b")], arc_whitelist=[("e", "a")], callback=rec, max_iters=3)
    assert best.has_arc("e", "a") and not best.has_arc("a", "b")
    assert rec.seen[0] == (0, None) and rec.seen[-1][1] is None and len(rec.seen) >= 3
    with pytest.raises(ValueError, match="blacklist and whitelist"):
        pbn.GreedyHillClimbing().estimate(pbn.ArcOperatorSet(), score, pbn.SemiparametricBN(names),
                                          arc_blacklist=[("a", "b")], arc_whitelist=[("a", "b")])
    with pytest.raises(ValueError, match="not initialized"):
        pbn.ArcOperatorSet().find_max(pbn.SemiparametricBN(names))
    with pytest.raises(ValueError, match="non-homogeneous"):
        pbn.ChangeNodeTypeSet().cache_scores(pbn.GaussianNetwork(names), score)


def test_hill_climbing_patience_tabu_runs():
    data = nonlinear_data(200, 2)
    names = list(data.columns)
    score = OracleCVScore(data, 5, 0)
    ghc = pbn.GreedyHillClimbing()
    m0 = ghc.estimate(pbn.OperatorPool([pbn.ArcOperatorSet(), pbn.ChangeNodeTypeSet()]), score, pbn.SemiparametricBN(names))
    ops0 = [str(o) for o in ghc.last_run["operators"]]
    m1 = ghc.estimate(pbn.OperatorPool([pbn.ArcOperatorSet(), pbn.ChangeNodeTypeSet()]), score, pbn.SemiparametricBN(names),
                      patience=2)
    # a plain Score never lowers the validation delta below its own operator delta: same search, same result
    assert [str(o) for o in ghc.last_run["operators"]][:len(ops0)] == ops0
    assert sorted(m0.arcs()) == sorted(m1.arcs())


def test_arguments_lookup_order():
    A = pbn.Arguments({"a": (pbn.ScottsBandwidth(),), pbn.CKDEType(): (pbn.NormalReferenceRule(),),
                       ("b", pbn.CKDEType()): pbn.Kwargs(bandwidth_selector=pbn.ScottsBandwidth())})
    assert isinstance(A.args("a", pbn.CKDEType())[0][0], pbn.ScottsBandwidth)
    assert isinstance(A.args("c", pbn.CKDEType())[0][0], pbn.NormalReferenceRule)
    assert isinstance(A.args("b", pbn.CKDEType())[1]["bandwidth_selector"], pbn.ScottsBandwidth)
    assert A.args("c", pbn.LinearGaussianCPDType()) == ((), {})
    with pytest.raises(ValueError):
        pbn.Arguments({3: ()})


def test_remove_and_add_node_keep_the_reference_index_semantics():
    """graph/generic_graph.hpp:509-580 and util/bidirectionalmap_index.hpp:59-65: raw indices are stable, a freed slot is
    reused by the next add_node, the collapsed order closes a gap with the LAST node.  The reference's
    hillclimbing_test.py:13-20,108-112 builds its start models this way."""
    import pybnesian_b200 as pbn
    g = pbn.GaussianNetwork(["a", "e", "b", "f", "c", "d"], [("a", "e"), ("e", "b"), ("b", "c"), ("f", "d")])
    g.remove_node("e")
    assert g.nodes() == ["a", "d", "b", "f", "c"] and g.num_nodes() == 5 and g.num_arcs() == 2
    assert g.index("d") == 5 and g.collapsed_index("d") == 1 and g.collapsed_name(1) == "d" and g.name(5) == "d"
    assert not g.contains_node("e") and g.arcs() == [("b", "c"), ("f", "d")]
    with pytest.raises(IndexError):
        g.index("e")
    g.remove_node("f")
    assert g.nodes() == ["a", "d", "b", "c"] and g.num_arcs() == 1
    assert sorted(g.graph().roots()) == ["a", "b", "d"]
    assert sorted(g.graph().topological_sort()) == ["a", "b", "c", "d"]
    assert g.add_node("z") == 3 and g.add_node("y") == 1 and g.add_node("x") == 6   # last freed slot first
    assert g.nodes() == ["a", "d", "b", "c", "z", "y", "x"]
    with pytest.raises(ValueError, match="same name"):
        g.add_node("a")
    g.add_arc("z", "a")
    assert g.can_add_arc("a", "b") and not g.can_add_arc("a", "z")
    c = g.clone()
    c.remove_node("z")
    assert g.has_arc("z", "a") and c.nodes() == ["a", "d", "b", "c", "x", "y"]
    import pickle
    r = pickle.loads(pickle.dumps(g))
    assert r.nodes() == g.nodes() and r.arcs() == g.arcs()
    sp = pbn.SemiparametricBN(["a", "b", "c"], node_types=[("b", pbn.CKDEType())])
    sp.remove_node("a")
    assert not sp.has_unknown_node_types() or sp.node_type("c") == pbn.UnknownFactorType()
    assert sp.node_type("b") == pbn.CKDEType() and sp.nodes() == ["c", "b"]


def test_clone_keeps_python_side_state_of_derived_networks():
    """reference hillclimbing_test.py:178-207 (NewBN.extra_data through __getstate_extra__ / __setstate_extra__)."""
    import pybnesian_b200 as pbn

    class NewType(pbn.BayesianNetworkType):
        def is_homogeneous(self):
            return True

        def default_node_type(self):
            return pbn.LinearGaussianCPDType()

        def new_bn(self, nodes):
            return NewBN(nodes)

    class NewBN(pbn.BayesianNetwork):
        def __init__(self, variables):
            pbn.BayesianNetwork.__init__(self, NewType(), variables)
            self.extra_data = "extra"

        def __getstate_extra__(self):
            return self.extra_data

        def __setstate_extra__(self, extra):
            self.extra_data = extra

    m = NewBN(["a", "b"])
    m.extra_data = "changed"
    c = m.clone()
    assert type(c) is NewBN and c.extra_data == "changed" and c.nodes() == ["a", "b"]


def test_arc_operator_set_uses_collapsed_indices_after_node_removal():
    """operators.cpp:36-52 (collapsed_from_index): the reference's hillclimbing_test.py:13-45 runs an ArcOperatorSet with an
    arc blacklist on a network whose nodes 'e' and 'f' were removed (raw indices 4, 5 of the survivors >= num_nodes)."""
    import pybnesian_b200 as pbn
    g = pbn.GaussianNetwork(["a", "e", "b", "f", "c", "d"])
    g.remove_node("e")
    g.remove_node("f")
    ops = pbn.ArcOperatorSet(blacklist=[("c", "d")], whitelist=[])
    ops.update_valid_ops(g)
    n = g.num_nodes()
    valid = ops._valid.reshape(n, n, order="F")
    ci = g.collapsed_index
    assert not valid[ci("c"), ci("d")] and valid[ci("d"), ci("c")] and not valid[ci("a"), ci("a")]
    assert int(valid.sum()) == n * n - n - 1 == len(ops._sorted_idx)
