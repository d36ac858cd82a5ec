"""CPU tests of the multi-rank plumbing with world_size 2 over gloo (no GPU): shard partitions, the
all-reduce of score vectors, and the candidate dealing of the batched scores (the device call is replaced
by the oracle so that the product's sharding logic itself runs on the CPU)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch.distributed as dist
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % os.environ["PORT"], rank=int(os.environ["RANK"]), world_size=2)
import oracle, util_data
import pybnesian_b200 as pbn
from pybnesian_b200 import parallel, _lib
from pybnesian_b200.scores import _FoldScorer

rank = parallel.rank()
assert parallel.active() and parallel.world_size() == 2
# 1. every element produced by one rank -> exact sum
v = np.zeros(7); v[rank::2] = np.arange(7)[rank::2] + 0.1
tot = parallel.all_reduce_sum(v)
assert np.array_equal(tot, np.arange(7) + 0.1)
# 2. shards tile [0, n)
b, e = parallel.shard_range(11)
sizes = parallel.all_reduce_sum(np.array([e - b if rank == 0 else 0.0, e - b if rank == 1 else 0.0]))
assert sizes.tolist() == [6.0, 5.0] and (b, e) == ((0, 6) if rank == 0 else (6, 11))
# 3. the batched score engine: items dealt over ranks, summed by all-reduce == serial evaluation
data = util_data.generate_normal_data(300, 0)
idx, lim = oracle.cv_indices(np.arange(300), 5, 0)

class CpuScorer(_FoldScorer):
    ran = []
    _fold_parallel = False  # the oracle scores whole items; the (item, fold) dealing is covered on GPUs (tools/multi_gpu_check.py)
    def _ctx(self, code):
        return None
    def _run_items(self, code, items):
        out = []
        for key, factor, rule, variables in items:
            CpuScorer.ran.append(tuple(variables))
            X = np.asfortranarray(data[variables].to_numpy())
            out.append(oracle.cv_score(X, idx, lim, "ckde" if factor == _lib.FACTOR_CKDE else "lg"))
        return np.array(out)

sc = CpuScorer(pbn.DataFrame(data), idx, lim, 0, 5, pbn.Arguments())
model = pbn.SemiparametricBN(list(data.columns))
fam = [("a", []), ("b", ["a"]), ("c", ["a", "b"]), ("d", ["a", "b", "c"]), ("d", ["c"]), ("c", ["d", "a"])]
reqs = [(pbn.CKDEType(), v, e) for v, e in fam] + [(pbn.LinearGaussianCPDType(), v, e) for v, e in fam]
got = sc.score_batch(model, reqs)
want = [oracle.cv_score(np.asfortranarray(data[[v] + e].to_numpy()), idx, lim, "ckde" if t == pbn.CKDEType() else "lg")
        for t, v, e in reqs]
assert got == want, (got, want)
n_mine = len(CpuScorer.ran)
counts = parallel.all_reduce_sum(np.array([n_mine if rank == 0 else 0.0, n_mine if rank == 1 else 0.0]))
assert counts.sum() == len(reqs) and abs(counts[0] - counts[1]) <= 1, counts
# 3b. (item, fold) dealing: jobs of one item land on both ranks, the folds are added in fold order
def fold_value(variables, f):
    return 1.0 / (3.0 + len(variables) + 7.0 * f) + 0.001 * sum(ord(c) for v in variables for c in v)

class FoldScorer(_FoldScorer):
    calls = []
    def _ctx(self, code):
        return None
    def _run_jobs(self, code, items, job_item, job_fold):
        FoldScorer.calls.append(len(job_item))
        assert list(zip(job_item, job_fold)) == sorted(zip(job_item, job_fold))
        return np.array([fold_value(items[i][3], f) for i, f in zip(job_item, job_fold)])

fs = FoldScorer(pbn.DataFrame(data), idx, lim, 0, 5, pbn.Arguments())
got = fs.score_batch(model, reqs)
for (t, v, e), g in zip(reqs, got):
    want_f = 0.0
    for f in range(5):
        want_f += fold_value([v] + e, f)
    assert g == want_f, (v, e, g, want_f)
assert len(FoldScorer.calls) == 1   # ONE batched call per rank however its jobs spread over the folds
jobs = sum(FoldScorer.calls)
tot = parallel.all_reduce_sum(np.array([jobs if rank == 0 else 0.0, jobs if rank == 1 else 0.0]))
assert tot.sum() == 5 * len(reqs) and abs(tot[0] - tot[1]) <= 1, tot
# 3c. a failure on ONE rank (an item only it was dealt) raises on EVERY rank after the collective - no rank is left
# waiting in the all-reduce (ADVICE round 1: the reference raises cleanly from its single process)
class FailingScorer(FoldScorer):
    def _run_jobs(self, code, items, job_item, job_fold):
        if rank == 1:
            raise _lib.SingularCovarianceData("Covariance matrix for variables [a, b] is not positive-definite.")
        return super()._run_jobs(code, items, job_item, job_fold)

bad = FailingScorer(pbn.DataFrame(data), idx, lim, 0, 5, pbn.Arguments())
try:
    bad.score_batch(model, reqs)
    raise AssertionError("no exception on rank %d" % rank)
except _lib.SingularCovarianceData as ex:
    assert "not positive-definite" in str(ex), str(ex)
with parallel.guard() as g:
    if rank == 0:
        raise ValueError("boom")
try:
    parallel.all_reduce_sum(np.zeros(3), error=g.error)
    raise AssertionError("no exception on rank %d" % rank)
except ValueError as ex:
    assert "boom" in str(ex)
assert parallel.all_reduce_sum(np.ones(2)).tolist() == [2.0, 2.0]   # the group is still usable afterwards
# 4. hill climbing on top of the sharded engine makes the same decisions on every rank
class ShardedScore(pbn.CVLikelihood):
    pass
score = pbn.CVLikelihood(data, 5, 0)
score._scorer = CpuScorer(pbn.DataFrame(data), idx, lim, 0, 5, pbn.Arguments())
ghc = pbn.GreedyHillClimbing()
best = ghc.estimate(pbn.OperatorPool([pbn.ArcOperatorSet(), pbn.ChangeNodeTypeSet()]), score, model, max_indegree=2)
ops = [str(o) for o in ghc.last_run["operators"]]
print("RESULT " + json.dumps({"rank": rank, "ops": ops, "arcs": sorted(best.arcs())}))
dist.barrier()
dist.destroy_process_group()
'''


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_world_size_2_gloo():
    import json
    port = _free_port()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), PORT=str(port), OMP_NUM_THREADS="2")
        procs.append(subprocess.Popen([sys.executable, "-c", "ROOT = %r\n" % ROOT + WORKER], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        out, _ = p.communicate(timeout=600)
        outs.append(out)
        assert p.returncode == 0, out
    res = [json.loads([l for l in o.splitlines() if l.startswith("RESULT ")][0][7:]) for o in outs]
    assert res[0]["ops"] == res[1]["ops"] and res[0]["arcs"] == res[1]["arcs"] and len(res[0]["ops"]) >= 3
    # and the same as the serial oracle search
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import util_data
    from oracle import hc as oracle_hc
    data = util_data.generate_normal_data(300, 0)
    want_ops, want_arcs, _, _ = oracle_hc.hill_climb(data.to_numpy(), k=5, seed=0, max_indegree=2)
    names = list(data.columns)
    assert [tuple(a) for a in res[0]["arcs"]] == [(names[s], names[t]) for s, t in want_arcs]
    assert len(res[0]["ops"]) == len(want_ops)


def test_deal_and_shard_partition():
    from pybnesian_b200 import parallel
    costs = [5, 1, 0, 5, 3, 0, 2, 4]
    for w in (1, 2, 3, 8):
        owned = [parallel.deal(costs, r, w) for r in range(w)]
        assert sorted(i for o in owned for i in o) == list(range(len(costs)))
        loads = [sum(costs[i] for i in o) for o in owned]
        assert max(loads) - min(loads) <= max(costs)
        spans = [parallel.shard_range(10, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == 10 and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
