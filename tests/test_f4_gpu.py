"""GPU tests for SURVEY §8 f4: the callers and on-disk format around the path - KDENetwork / SemiparametricBN
default hc() flow (util/validate_options.cpp:16-91: ValidatedLikelihood + ArcOperatorSet [+ ChangeNodeTypeSet]),
save / load of models with GPU-resident factors (models/BayesianNetwork.hpp:1127-1167, KDE.hpp:642-666,
CKDE.cpp:164-218; the reference's tests/serialization/*) and the SaveModel callback."""
import os

import numpy as np
import pytest

import util_data

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pbn():
    import pybnesian_b200 as pbn
    return pbn


def test_fitted_spbn_save_load_keeps_gpu_factors(pbn, tmp_path):
    df = util_data.generate_normal_data(800, 0)
    test = util_data.generate_normal_data(100, 1)
    m = pbn.SemiparametricBN(["a", "b", "c", "d"], [("a", "b"), ("a", "c"), ("b", "c"), ("c", "d")],
                             [("b", pbn.CKDEType()), ("c", pbn.CKDEType())])
    m.fit(df)
    m.save(str(tmp_path / "spbn"), include_cpd=True)
    r = pbn.load(str(tmp_path / "spbn.pickle"))
    assert r.fitted() and r.node_type("c") == pbn.CKDEType() and type(r.cpd("c")) is pbn.CKDE
    assert r.cpd("c").num_instances() == 800 and r.cpd("c").evidence() == m.cpd("c").evidence()
    assert np.allclose(r.logl(test), m.logl(test), rtol=1e-13, atol=0)
    assert r.slogl(test) == pytest.approx(m.slogl(test), rel=1e-13)
    assert np.allclose(r.cpd("c").cdf(test), m.cpd("c").cdf(test), rtol=1e-13, atol=1e-15)
    assert r.sample(50, 3, ordered=True).equals(m.sample(50, 3, ordered=True))
    m.save(str(tmp_path / "bare"))
    assert not pbn.load(str(tmp_path / "bare.pickle")).fitted()


def test_kdenetwork_default_hc_flow_with_save_model(pbn, tmp_path):
    df = util_data.generate_normal_data(300, 0)
    cb = pbn.SaveModel(str(tmp_path))
    model = pbn.hc(df, bn_type=pbn.KDENetworkType(), seed=0, num_folds=5, max_indegree=2, callback=cb)
    assert type(model) is pbn.KDENetwork
    assert all(model.node_type(n) == pbn.CKDEType() for n in model.nodes())
    assert model.num_arcs() >= 3  # a -> b -> c -> d chain of the generator is recoverable from 300 rows
    files = sorted(os.listdir(tmp_path))
    assert files[0] == "000000.pickle" and len(files) >= model.num_arcs() + 1
    # the last saved model is the returned structure
    assert sorted(pbn.load(str(tmp_path / files[-1])).arcs()) == sorted(model.arcs())
    # the defaults are ValidatedLikelihood(test_ratio 0.2, k) + ArcOperatorSet: the explicit call gives the same DAG
    score = pbn.ValidatedLikelihood(df, test_ratio=0.2, k=5, seed=0)
    start = pbn.KDENetwork(list(df.columns))
    explicit = pbn.GreedyHillClimbing().estimate(pbn.ArcOperatorSet(max_indegree=2), score, start, max_indegree=2)
    assert sorted(explicit.arcs()) == sorted(model.arcs())
    model.fit(df)
    assert np.isfinite(model.slogl(df))
    assert len(model.sample(20, 1)) == 20


def test_spbn_default_hc_flow(pbn):
    df = util_data.generate_normal_data(300, 0)
    model = pbn.hc(df, bn_type=pbn.SemiparametricBNType(), seed=1, num_folds=3, max_indegree=2, max_iters=6)
    assert type(model) is pbn.SemiparametricBN and model.num_arcs() >= 2
    assert all(model.node_type(n) in (pbn.CKDEType(), pbn.LinearGaussianCPDType()) for n in model.nodes())
