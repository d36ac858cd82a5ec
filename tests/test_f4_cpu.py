"""CPU tests (no GPU) for SURVEY §8 f4: model save / load with and without CPDs (BNGeneric::__getstate__,
models/BayesianNetwork.hpp:1127-1167; the reference's tests/serialization/serialize_models_test.py), add_cpds /
check_compatible_cpd (863-940) and the SaveModel callback (callbacks/save_model.hpp)."""
import os
import pickle

import numpy as np
import pytest

import pybnesian_b200 as pbn


def small_spbn():
    return pbn.SemiparametricBN(["a", "b", "c", "d"], [("a", "b"), ("b", "c"), ("a", "c")],
                                [("c", pbn.CKDEType()), ("a", pbn.LinearGaussianCPDType())])


def test_save_load_structure_only(tmp_path):
    m = small_spbn()
    m.save(str(tmp_path / "model"))
    assert os.path.exists(tmp_path / "model.pickle")
    r = pbn.load(str(tmp_path / "model.pickle"))
    assert type(r) is pbn.SemiparametricBN and r.type() == pbn.SemiparametricBNType()
    assert r.nodes() == m.nodes() and sorted(r.arcs()) == sorted(m.arcs())
    assert r.node_type("c") == pbn.CKDEType() and r.node_type("a") == pbn.LinearGaussianCPDType()
    assert r.node_type("b") == pbn.UnknownFactorType()
    assert not r.fitted() and not r.include_cpd()
    g = pbn.GaussianNetwork(["a", "b"], [("a", "b")])
    r = pickle.loads(pickle.dumps(g))
    assert type(r) is pbn.GaussianNetwork and r.arcs() == [("a", "b")] and r.node_type("a") == pbn.LinearGaussianCPDType()


def test_add_cpds_and_include_cpd_roundtrip(tmp_path):
    m = pbn.GaussianNetwork(["a", "b", "c"], [("a", "b"), ("a", "c"), ("b", "c")])
    cpds = [pbn.LinearGaussianCPD("a", [], [1.0], 0.5), pbn.LinearGaussianCPD("b", ["a"], [0.5, 2.0], 1.5),
            pbn.LinearGaussianCPD("c", ["b", "a"], [0.0, -1.0, 3.0], 0.25)]
    m.add_cpds(cpds)
    assert m.fitted() and m.cpd("b").variance == 1.5
    m.save(str(tmp_path / "with_cpd"), include_cpd=True)
    r = pbn.load(str(tmp_path / "with_cpd.pickle"))
    assert r.fitted() and r.include_cpd()
    for n in "abc":
        assert r.cpd(n).evidence() == m.cpd(n).evidence()
        assert np.array_equal(r.cpd(n).beta, m.cpd(n).beta) and r.cpd(n).variance == m.cpd(n).variance
    m.save(str(tmp_path / "without_cpd"), include_cpd=False)
    r = pbn.load(str(tmp_path / "without_cpd.pickle"))
    assert not r.fitted()
    with pytest.raises(ValueError, match="not added"):
        r.cpd("a")


def test_check_compatible_cpd_errors():
    m = pbn.GaussianNetwork(["a", "b", "c"], [("a", "b")])
    with pytest.raises(ValueError, match="not present in the model"):
        m.add_cpds([pbn.LinearGaussianCPD("z", [], [0.0], 1.0)])
    with pytest.raises(ValueError, match="is not present in the model"):
        m.add_cpds([pbn.LinearGaussianCPD("b", ["z"], [0.0, 1.0], 1.0)])
    with pytest.raises(ValueError, match="parent set as evidence"):
        m.add_cpds([pbn.LinearGaussianCPD("b", ["c"], [0.0, 1.0], 1.0)])
    with pytest.raises(ValueError, match="parent set as evidence"):
        m.add_cpds([pbn.LinearGaussianCPD("b", [], [0.0], 1.0)])
    s = pbn.SemiparametricBN(["a", "b"], [("a", "b")], [("b", pbn.CKDEType())])
    with pytest.raises(ValueError, match="Bayesian network expects type"):
        s.add_cpds([pbn.LinearGaussianCPD("b", ["a"], [0.0, 1.0], 1.0)])
    s.add_cpds([pbn.LinearGaussianCPD("a", [], [0.0], 1.0)])  # unknown type: adopts the CPD's
    assert s.node_type("a") == pbn.LinearGaussianCPDType()


def test_save_model_callback(tmp_path):
    cb = pbn.SaveModel(str(tmp_path))
    m = small_spbn()
    cb.call(m, None, None, 0)
    m.add_arc("c", "d")
    cb.call(m, pbn.AddArc("c", "d", 1.0), None, 7)
    assert sorted(os.listdir(tmp_path)) == ["000000.pickle", "000007.pickle"]
    assert pbn.load(str(tmp_path / "000007.pickle")).has_arc("c", "d")
    assert not pbn.load(str(tmp_path / "000000.pickle")).has_arc("c", "d")
