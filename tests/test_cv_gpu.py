"""GPU parity tests of the batched likelihood scores and of hill climbing on top of them.

The checker is the CPU oracle (oracle.cv_score: one fit + slogl per fold, exactly the reference's
serial CVLikelihood::local_score) and, for the shapes of the reference's own
tests/learning/scores/cvlikelihood_test.py / holdoutlikelihood_test.py, SciPy / lstsq as an
independent check.  float64: 1e-10 relative; float32: 1e-4 relative (BASELINE.json north_star).
Fold indices and the hill-climbing operator sequence must be identical.
"""
import numpy as np
import pandas as pd
import pytest
from scipy.stats import gaussian_kde, norm

import oracle
from oracle import hc as oracle_hc
import util_data
from test_host_logic import nonlinear_data, _as_tuples

pytestmark = pytest.mark.gpu

SIZE = 1000
df = util_data.generate_normal_data(SIZE)
seed = 0


@pytest.fixture(scope="module")
def pbn():
    import pybnesian_b200 as pbn
    return pbn


def oracle_cv(data, variable, evidence, factor, k=10, sd=0, rule="normal_reference"):
    X = np.asfortranarray(data[[variable] + evidence].to_numpy())
    idx, lim = oracle.cv_indices(np.arange(len(data)), k, sd)
    return oracle.cv_score(X, idx, lim, factor, rule)


def numpy_local_score(pbn, node_type, data, variable, evidence):
    """tests/learning/scores/cvlikelihood_test.py:12-49 of the reference, folds from our CrossValidation mirror."""
    cv = pbn.CrossValidation(data, 10, seed)
    loglik = 0
    for train_df, test_df in cv:
        node_data = train_df.to_pandas().loc[:, [variable] + evidence].dropna()
        test_node_data = test_df.to_pandas().loc[:, [variable] + evidence].dropna()
        if node_type == pbn.LinearGaussianCPDType():
            N, d = node_data.shape[0], len(evidence)
            A = np.column_stack((np.ones(N), node_data.loc[:, evidence].to_numpy()))
            beta, res, _, _ = np.linalg.lstsq(A, node_data.loc[:, variable].to_numpy(), rcond=None)
            var = res / (N - d - 1)
            means = beta[0] + np.sum(beta[1:] * test_node_data.loc[:, evidence], axis=1)
            loglik += norm.logpdf(test_node_data.loc[:, variable], means, np.sqrt(var)).sum()
        else:
            kj = gaussian_kde(node_data.to_numpy().T, bw_method=lambda s: np.power(4 / (s.d + 2), 1 / (s.d + 4)) * s.scotts_factor())
            if evidence:
                km = gaussian_kde(node_data.loc[:, evidence].to_numpy().T, bw_method=kj.covariance_factor())
                loglik += np.sum(kj.logpdf(test_node_data.to_numpy().T) - km.logpdf(test_node_data.loc[:, evidence].to_numpy().T))
            else:
                loglik += np.sum(kj.logpdf(test_node_data.to_numpy().T))
    return loglik


FAMILIES = [("a", []), ("b", ["a"]), ("c", ["a", "b"]), ("d", ["a", "b", "c"]), ("d", ["c", "a"]), ("a", ["d", "c", "b"])]


def test_cvl_create(pbn):
    s = pbn.CVLikelihood(df)
    assert len(list(s.cv)) == 10
    assert len(list(pbn.CVLikelihood(df, 5).cv)) == 5
    s, s2 = pbn.CVLikelihood(df, 10, 0), pbn.CVLikelihood(df, 10, 0)
    for (tr, te), (tr2, te2) in zip(s.cv, s2.cv):
        assert tr.equals(tr2) and te.equals(te2)
    with pytest.raises(ValueError, match="Cannot split"):
        pbn.CVLikelihood(df, SIZE + 1)


@pytest.mark.parametrize("variable,evidence", FAMILIES)
def test_cvl_local_score_f64(pbn, variable, evidence):
    cvl = pbn.CVLikelihood(df, 10, seed)
    gbn = pbn.GaussianNetwork(list(df.columns))
    kdn = pbn.KDENetwork(list(df.columns))
    got_lg = cvl.local_score(gbn, variable, evidence)
    got_ck = cvl.local_score(kdn, variable, evidence)
    want_lg = oracle_cv(df, variable, evidence, "lg")
    want_ck = oracle_cv(df, variable, evidence, "ckde")
    assert abs(got_lg - want_lg) <= 1e-10 * abs(want_lg)
    assert abs(got_ck - want_ck) <= 1e-10 * abs(want_ck)
    # the reference's own (SciPy / lstsq) checks
    assert np.isclose(got_lg, numpy_local_score(pbn, pbn.LinearGaussianCPDType(), df, variable, evidence))
    assert np.isclose(got_ck, numpy_local_score(pbn, pbn.CKDEType(), df, variable, evidence))


def test_cvl_batch_equals_single_and_memo(pbn):
    cvl = pbn.CVLikelihood(df, 10, seed)
    spbn = pbn.SemiparametricBN(list(df.columns))
    reqs = [(pbn.CKDEType(), v, e) for v, e in FAMILIES] + [(pbn.LinearGaussianCPDType(), v, e) for v, e in FAMILIES]
    batch = cvl.local_score_batch(spbn, reqs)
    fresh = pbn.CVLikelihood(df, 10, seed)
    single = [fresh.local_score_node_type(spbn, t, v, e) for t, v, e in reqs]
    # same kernels on the same jobs; only the unit split (hence the order of partial sums) may differ
    assert np.allclose(batch, single, rtol=1e-13, atol=0)
    again = cvl.local_score_batch(spbn, reqs)
    assert again == batch and cvl._scorer.stats["memo_hits"] >= len(reqs)
    # evidence order is part of the key (it changes rounding in the reference too)
    a = cvl.local_score_node_type(spbn, pbn.CKDEType(), "d", ["a", "b"])
    b = cvl.local_score_node_type(spbn, pbn.CKDEType(), "d", ["b", "a"])
    assert abs(a - b) <= 1e-10 * abs(a)


def test_cvl_local_score_f32(pbn):
    df32 = df.astype(np.float32)
    cvl = pbn.CVLikelihood(df32, 10, seed)
    spbn = pbn.SemiparametricBN(list(df.columns))
    for variable, evidence in FAMILIES:
        for t, f in ((pbn.CKDEType(), "ckde"), (pbn.LinearGaussianCPDType(), "lg")):
            got = cvl.local_score_node_type(spbn, t, variable, evidence)
            want = oracle_cv(df32, variable, evidence, f)
            want64 = oracle_cv(df, variable, evidence, f)
            # float32 bar of the north star; the oracle itself (reference arithmetic in float) is only that close
            # to the float64 value
            assert abs(got - want) <= 1e-4 * abs(want), (variable, evidence, f, got, want)
            assert abs(got - want64) <= 2e-4 * abs(want64)


def test_cvl_scott_and_python_selector_and_ucv(pbn):
    spbn = pbn.SemiparametricBN(list(df.columns))
    args = pbn.Arguments({pbn.CKDEType(): (pbn.ScottsBandwidth(),)})
    cvl = pbn.CVLikelihood(df, 10, seed, args)
    got = cvl.local_score_node_type(spbn, pbn.CKDEType(), "c", ["a", "b"])
    want = oracle_cv(df, "c", ["a", "b"], "ckde", rule="scott")
    assert abs(got - want) <= 1e-10 * abs(want)

    class MyNR(pbn.BandwidthSelector):  # Python-derived selector: the reference's generic per-fold loop
        def bandwidth(self, d, variables):
            return pbn.NormalReferenceRule().bandwidth(d, variables)

    cvl2 = pbn.CVLikelihood(df, 10, seed, pbn.Arguments({"c": (MyNR(),)}))
    got2 = cvl2.local_score_node_type(spbn, pbn.CKDEType(), "c", ["a", "b"])
    want2 = oracle_cv(df, "c", ["a", "b"], "ckde")
    assert abs(got2 - want2) <= 1e-10 * abs(want2)
    assert cvl2._scorer.stats["host_items"] == 1


def test_cv_train_moments_match_numpy(pbn):
    import ctypes
    from pybnesian_b200 import _lib
    cvl = pbn.CVLikelihood(df, 10, seed)
    spbn = pbn.SemiparametricBN(list(df.columns))
    cvl.local_score_node_type(spbn, pbn.LinearGaussianCPDType(), "a", [])
    handle, index = cvl._scorer._device(_lib.PBN_F64)
    cols = ["d", "b", "a"]
    vars_ = _lib.int_array([index[c] for c in cols])
    for f, (train, _) in enumerate(cvl.cv.indices()):
        mean, cov = np.empty(3), np.empty((3, 3), order="F")
        dp = ctypes.POINTER(ctypes.c_double)
        _lib.check(_lib.lib().pbn_cv_train_moments(handle.h, f, vars_, 3, mean.ctypes.data_as(dp), cov.ctypes.data_as(dp)))
        X = df.iloc[train][cols].to_numpy()
        assert np.allclose(mean, X.mean(axis=0), rtol=1e-13, atol=0)
        assert np.allclose(cov, np.cov(X.T), rtol=1e-12, atol=0)


def test_cvl_null_rows_are_excluded(pbn):
    np.random.seed(0)
    df_null = df.copy()
    for c in "abcd":
        df_null.loc[df_null.index[np.random.randint(0, SIZE, size=40)], c] = np.nan
    cvl = pbn.CVLikelihood(df_null, 10, seed)
    spbn = pbn.SemiparametricBN(list(df.columns))
    clean = df_null.dropna()
    valid = np.flatnonzero(~df_null.isna().any(axis=1).to_numpy())
    idx, lim = oracle.cv_indices(valid, 10, seed)
    for t, f in ((pbn.CKDEType(), "ckde"), (pbn.LinearGaussianCPDType(), "lg")):
        X = np.asfortranarray(df_null[["c", "a", "b"]].to_numpy())
        want = oracle.cv_score(X, idx, lim, f)
        got = cvl.local_score_node_type(spbn, t, "c", ["a", "b"])
        assert abs(got - want) <= 1e-10 * abs(want)
    assert len(clean) == valid.size


def test_cvl_singular_covariance(pbn):
    tiny = util_data.generate_normal_data(12)
    spbn = pbn.SemiparametricBN(list(df.columns))
    cvl = pbn.CVLikelihood(tiny, 6, seed)   # training folds of 10 rows
    cvl.local_score_node_type(spbn, pbn.CKDEType(), "a", ["b"])
    small = util_data.generate_normal_data(5)
    cvl = pbn.CVLikelihood(small, 5, seed)  # 4 training rows, 4 variables: rows <= d
    with pytest.raises(pbn.SingularCovarianceData):
        cvl.local_score_node_type(spbn, pbn.CKDEType(), "d", ["a", "b", "c"])
    dup = df.copy()
    dup["b"] = dup["a"] * 2.0
    cvl = pbn.CVLikelihood(dup, 10, seed)
    with pytest.raises(pbn.SingularCovarianceData, match="positive-definite"):
        cvl.local_score_node_type(spbn, pbn.CKDEType(), "c", ["a", "b"])


def test_cvl_wide_family_uses_generic_path(pbn):
    wide = util_data.iid_normal(400, 10, 3)
    wide["x0"] = wide["x0"] + 0.5 * wide["x1"] - 0.3 * wide["x9"]
    cvl = pbn.CVLikelihood(wide, 4, seed)
    kdn = pbn.KDENetwork(list(wide.columns))
    ev = ["x%d" % i for i in range(1, 10)]
    got = cvl.local_score(kdn, "x0", ev)   # 10 variables > 8: per-fold launches through pbn_ckde_fit / pbn_kde_logl
    want = oracle_cv(wide, "x0", ev, "ckde", k=4)
    assert abs(got - want) <= 1e-10 * abs(want)


def test_holdout_and_validated_likelihood(pbn):
    spbn = pbn.SemiparametricBN(list(df.columns))
    hl = pbn.HoldoutLikelihood(df, 0.2, seed)
    assert hl.training_data().num_rows == 0.8 * SIZE and hl.test_data().num_rows == 0.2 * SIZE
    tr, te = oracle.holdout_indices(np.arange(SIZE), 0.2, seed)
    for variable, evidence in FAMILIES:
        cols = [variable] + evidence
        Xtr, Xte = df.iloc[tr][cols].to_numpy(), df.iloc[te][cols].to_numpy()
        want_ck = oracle.ckde_logl(Xtr, Xte, oracle.bandwidth(Xtr))[1]
        got_ck = hl.local_score_node_type(spbn, pbn.CKDEType(), variable, evidence)
        assert abs(got_ck - want_ck) <= 1e-10 * abs(want_ck)
        beta, var = oracle.lg_fit(Xtr[:, 0], [Xtr[:, j] for j in range(1, len(cols))])
        want_lg = oracle.lg_logl(Xte[:, 0], [Xte[:, j] for j in range(1, len(cols))], beta, var)[1]
        got_lg = hl.local_score_node_type(spbn, pbn.LinearGaussianCPDType(), variable, evidence)
        assert abs(got_lg - want_lg) <= 1e-10 * abs(want_lg)
    with pytest.raises(ValueError, match="test_ratio must be a number"):
        pbn.HoldoutLikelihood(df, 1.2, seed)

    vl = pbn.ValidatedLikelihood(df, 0.2, 10, seed)
    train_part = df.iloc[tr].reset_index(drop=True)
    for variable, evidence in FAMILIES[1:4]:
        want = oracle_cv(train_part, variable, evidence, "ckde")
        got = vl.local_score_node_type(spbn, pbn.CKDEType(), variable, evidence)
        assert abs(got - want) <= 1e-10 * abs(want)
        assert vl.vlocal_score_node_type(spbn, pbn.CKDEType(), variable, evidence) == \
            hl.local_score_node_type(spbn, pbn.CKDEType(), variable, evidence)
    assert vl.training_data().num_rows == 800 and vl.validation_data().num_rows == 200


def test_linear_gaussian_cpd_fit_logl(pbn):
    test = util_data.generate_normal_data(200, seed=1)
    for variable, evidence in FAMILIES + [("d", ["a", "b", "c"])]:
        cpd = pbn.LinearGaussianCPD(variable, evidence)
        assert not cpd.fitted()
        cpd.fit(df)
        cols = [variable] + evidence
        X, T = df[cols].to_numpy(), test[cols].to_numpy()
        beta, var = oracle.lg_fit(X[:, 0], [X[:, j] for j in range(1, len(cols))])
        assert np.allclose(cpd.beta, beta, rtol=1e-9, atol=1e-12) and abs(cpd.variance - var) <= 1e-10 * var
        want, want_s = oracle.lg_logl(T[:, 0], [T[:, j] for j in range(1, len(cols))], beta, var)
        got = cpd.logl(test)
        assert np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-300)) < 1e-9
        assert abs(cpd.slogl(test) - want_s) <= 1e-10 * abs(want_s)
    tn = test.copy()
    tn.loc[tn.index[[3, 17]], "a"] = np.nan
    cpd = pbn.LinearGaussianCPD("b", ["a"])
    cpd.fit(df)
    l = cpd.logl(tn)
    assert np.isnan(l[[3, 17]]).all() and np.isfinite(np.delete(l, [3, 17])).all()
    assert abs(cpd.slogl(tn) - np.nansum(l)) <= 1e-10 * abs(np.nansum(l))


@pytest.mark.parametrize("rows,k,max_indegree", [(400, 5, 0), (2000, 10, 3)])
def test_hill_climbing_operator_sequence_matches_oracle(pbn, rows, k, max_indegree):
    """SURVEY.md §8d config 4, reduced: identical folds, identical operator list, deltas at 1e-10 of the scores."""
    data = nonlinear_data(rows, 0)
    names = list(data.columns)
    want_ops, want_arcs, want_types, _ = oracle_hc.hill_climb(data.to_numpy(), k=k, seed=0, max_indegree=max_indegree)
    score = pbn.CVLikelihood(data, k, 0)
    pool = pbn.OperatorPool([pbn.ArcOperatorSet(), pbn.ChangeNodeTypeSet()])
    ghc = pbn.GreedyHillClimbing()
    best = ghc.estimate(pool, score, pbn.SemiparametricBN(names), max_indegree=max_indegree)
    got_ops = _as_tuples(ghc.last_run["operators"], names)
    assert [o[:3] for o in got_ops] == [o[:3] for o in want_ops]
    scale = abs(score.score(best))
    assert np.allclose([o[3] for o in got_ops], [o[3] for o in want_ops], rtol=0, atol=1e-10 * scale)
    assert sorted((names.index(s), names.index(t)) for s, t in best.arcs()) == want_arcs
    assert [str(best.node_type(n)) for n in names] == want_types
    st = score._scorer.stats
    assert st["device_items"] > 0 and st["batches"] < st["device_items"]


def test_hc_convenience_function(pbn):
    data = nonlinear_data(300, 1)
    m = pbn.hc(data, bn_type=pbn.SemiparametricBNType(), score="cv-lik", seed=0, num_folds=5, max_indegree=2)
    want_ops, want_arcs, want_types, _ = oracle_hc.hill_climb(data.to_numpy(), k=5, seed=0, max_indegree=2)
    names = list(data.columns)
    assert sorted((names.index(s), names.index(t)) for s, t in m.arcs()) == want_arcs
    assert [str(m.node_type(n)) for n in names] == want_types
    v = pbn.hc(data, start=pbn.SemiparametricBN(names), seed=0, num_folds=5, max_iters=4)   # default score: validated-lik
    assert v.num_arcs() >= 1
    with pytest.raises(ValueError):
        pbn.hc(data)


def test_multi_gpu_shards_match_single_gpu():
    """Needs >= 2 GPUs (run with `gpurun --gpus 2`): test-row shards, dealt score batches, UCV tile slices and the
    replicated hill climbing give the single-GPU results (tools/multi_gpu_check.py under torchrun)."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(root, "tools", "multi_gpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "MULTI_GPU_CHECK_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


def test_cvl_folds_with_isolated_outliers_take_the_shifted_pass(pbn):
    """Isolated outliers: in the fold that tests them they are hundreds of bandwidths from every training row, so their
    unshifted sums underflow inside the batched launch; they are re-evaluated by the shifted second pass (runtime.cu:
    pbn_shift_pass, job by job) and the score still equals the oracle's serial fit + slogl per fold."""
    data = util_data.generate_normal_data(3000, 0)
    rng = np.random.default_rng(4)
    # geometric spacing: the bandwidth grows with the largest outlier, so evenly spaced ones would stay within a few
    # bandwidths of each other; here the two or three largest of each column are > 40 bandwidths from any other row
    idx = rng.choice(3000, 24, replace=False)
    data.loc[idx[:12], "a"] += 200.0 * 3.0 ** np.arange(12) * rng.choice([-1.0, 1.0], 12)
    data.loc[idx[12:], "c"] -= 300.0 * 3.0 ** np.arange(12)
    cvl = pbn.CVLikelihood(data, 10, seed)
    kdn = pbn.KDENetwork(list(data.columns))
    ctx = pbn.default_context()
    for variable, evidence in [("a", []), ("c", ["a"]), ("c", ["a", "b"]), ("d", ["a", "b", "c"])]:
        got = cvl.local_score(kdn, variable, evidence)
        assert ctx.last_fallback_rows() > 0
        want = oracle_cv(data, variable, evidence, "ckde")
        assert np.isfinite(got) and abs(got - want) <= 1e-10 * abs(want), (variable, evidence, got, want)
    # batch = single, outliers or not
    reqs = [(pbn.CKDEType(), "c", ["a", "b"]), (pbn.CKDEType(), "d", ["a"]), (pbn.LinearGaussianCPDType(), "b", ["a"])]
    spbn = pbn.SemiparametricBN(list(data.columns))
    batch = pbn.CVLikelihood(data, 10, seed).local_score_batch(spbn, reqs)
    single = [pbn.CVLikelihood(data, 10, seed).local_score_node_type(spbn, t, v, e) for t, v, e in reqs]
    assert np.allclose(batch, single, rtol=1e-13, atol=0)


def test_in_process_multi_device_context_matches_single_device():
    """Needs >= 2 GPUs: ONE process, a context over all devices (pbn_ctx_create_multi, no torchrun) - sharded logl / slogl /
    cdf, dealt CV jobs, per-device UCV slices and hill climbing against the single-device results
    (tools/inproc_multi_gpu_check.py)."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = {k: v for k, v in os.environ.items() if k not in ("PBN_CUDA_DEVICE", "PBN_CUDA_DEVICES", "LOCAL_RANK")}
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "inproc_multi_gpu_check.py"), "--no-timing"],
                         capture_output=True, text=True, timeout=1200, env=env)
    assert out.returncode == 0 and "INPROC_MULTI_GPU_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


def test_mle_linear_gaussian_like_the_reference_test(pbn):
    """tests/learning/parameters/mle_test.py: MLE(LinearGaussianCPDType()).estimate against numpy lstsq;
    MLE(CKDEType()) is not available."""
    df = util_data.generate_normal_data(10000, 0)
    with pytest.raises(ValueError, match="MLE not available"):
        pbn.MLE(pbn.CKDEType())
    mle = pbn.MLE(pbn.LinearGaussianCPDType())
    for variable, evidence in [("a", []), ("b", ["a"]), ("c", ["a", "b"]), ("d", ["a", "b", "c"])]:
        p = mle.estimate(df, variable, evidence)
        A = np.column_stack([np.ones(len(df))] + [df[e].to_numpy() for e in evidence])
        beta, res, _, _ = np.linalg.lstsq(A, df[variable].to_numpy(), rcond=None)
        assert np.allclose(p.beta, beta, rtol=1e-8, atol=1e-10)
        assert np.isclose(p.variance, res[0] / (len(df) - len(evidence) - 1), rtol=1e-9)
    dfn = df.copy()
    dfn.loc[5, "a"] = np.nan
    p = mle.estimate(dfn, "b", ["a"])
    sub = dfn.dropna()
    A = np.column_stack([np.ones(len(sub)), sub["a"].to_numpy()])
    beta, _, _, _ = np.linalg.lstsq(A, sub["b"].to_numpy(), rcond=None)
    assert np.allclose(p.beta, beta, rtol=1e-8)


def _numpy_bic(data, variable, evidence):
    """tests/learning/scores/bic_test.py:10-30."""
    from scipy.stats import norm
    nd = data[[variable] + evidence].dropna()
    N, d = len(nd), len(evidence)
    A = np.column_stack([np.ones(N)] + [nd[e].to_numpy() for e in evidence])
    beta, res, _, _ = np.linalg.lstsq(A, nd[variable].to_numpy(), rcond=None)
    var = res[0] / (N - d - 1)
    means = A @ beta
    return norm.logpdf(nd[variable].to_numpy(), means, np.sqrt(var)).sum() - np.log(N) * 0.5 * (d + 2)


def test_bic_like_the_reference_test_and_gaussian_hc(pbn):
    df = util_data.generate_normal_data(10000, 0)
    gbn = pbn.GaussianNetwork(["a", "b", "c", "d"], [("a", "b"), ("a", "c"), ("a", "d"), ("b", "c"), ("b", "d"), ("c", "d")])
    bic = pbn.BIC(df)
    for v, e in [("a", []), ("b", ["a"]), ("c", ["a", "b"]), ("d", ["a", "b", "c"])]:
        assert np.isclose(bic.local_score(gbn, v, e), _numpy_bic(df, v, e), rtol=1e-9)
        assert bic.local_score(gbn, v) == bic.local_score(gbn, v, gbn.parents(v))
    assert np.isclose(bic.local_score(gbn, "d", ["b", "c", "a"]), _numpy_bic(df, "d", ["a", "b", "c"]), rtol=1e-9)
    assert np.isclose(bic.score(gbn), sum(_numpy_bic(df, v, gbn.parents(v)) for v in "abcd"), rtol=1e-9)
    dfn = df.copy()
    rng = np.random.default_rng(0)
    for c in "abcd":
        dfn.loc[rng.integers(0, len(df), 100), c] = np.nan
    bn = pbn.BIC(dfn)
    assert np.isclose(bn.local_score(gbn, "c", ["a", "b"]), _numpy_bic(dfn, "c", ["a", "b"]), rtol=1e-9)
    # hillclimbing_test.py:8-60 with the GaussianNetwork defaults (score "bic", arc operators)
    small = util_data.generate_normal_data(1000, 0)
    start = pbn.GaussianNetwork(list(small.columns))
    sb = pbn.BIC(small)
    res = pbn.GreedyHillClimbing().estimate(pbn.ArcOperatorSet(), sb, start, max_iters=1)
    assert res.num_arcs() == 1
    added = res.arcs()[0]
    delta = sb.score(res) - sb.score(start)
    assert np.isclose(delta, sb.local_score(res, added[1], [added[0]]) - sb.local_score(res, added[1], []))
    # BIC is score equivalent: blacklisting the arc adds its reverse
    res2 = pbn.GreedyHillClimbing().estimate(pbn.ArcOperatorSet(), sb, start, max_iters=1, arc_blacklist=[added])
    assert res2.arcs()[0] == added[::-1]
    assert pbn.GreedyHillClimbing().estimate(pbn.ArcOperatorSet(), sb, start, epsilon=delta + 0.01).num_arcs() == 0
    full = pbn.hc(small, bn_type=pbn.GaussianNetworkType())
    assert type(full) is pbn.GaussianNetwork and full.num_arcs() >= 5
    assert pbn.hc(small, bn_type=pbn.GaussianNetworkType(), score="bic").num_arcs() == full.num_arcs()
