"""Run under torchrun on N GPUs of one box: the sharded paths (slogl / logl test-row shards, dealt score
batches, UCV pair-tile slices, replicated hill climbing) must give the single-GPU results.
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch
import torch.distributed as dist

lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
os.environ["PBN_CUDA_DEVICE"] = str(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))

import util_data
import pybnesian_b200 as pbn
from pybnesian_b200 import parallel
from test_host_logic import nonlinear_data

rank, world = parallel.rank(), parallel.world_size()
report = {"world": world}

train, test = util_data.generate_normal_data(20000, 0), util_data.generate_normal_data(5003, 1)
cpd = pbn.CKDE("d", ["a", "b", "c"])
cpd.fit(train)
s_sharded, l_sharded = cpd.slogl(test), cpd.logl(test)
parallel.enable(False)
s_single, l_single = cpd.slogl(test), cpd.logl(test)
parallel.enable(True)
report["slogl_rel_diff"] = abs(s_sharded - s_single) / abs(s_single)
report["logl_max_rel_diff"] = float(np.max(np.abs(l_sharded - l_single) / np.abs(l_single)))
# a different unit split changes the order of the partial sums: equal to rounding, not bit for bit
assert report["slogl_rel_diff"] < 1e-12 and report["logl_max_rel_diff"] < 1e-12, report
c_sharded = cpd.cdf(test)
smp_sharded = cpd.sample(1000, test[["a", "b", "c"]], 5).to_numpy()  # replicated: every rank draws the same stream
parallel.enable(False)
c_single = cpd.cdf(test)
smp_single = cpd.sample(1000, test[["a", "b", "c"]], 5).to_numpy()
parallel.enable(True)
report["cdf_max_abs_diff"] = float(np.max(np.abs(c_sharded - c_single)))
assert report["cdf_max_abs_diff"] < 1e-13 and np.array_equal(smp_sharded, smp_single), report

data = nonlinear_data(3000, 0)
names = list(data.columns)
model = pbn.SemiparametricBN(names)
fam = [("a", []), ("b", ["a"]), ("c", ["a", "b"]), ("d", ["a", "b", "c"]), ("e", ["c", "d"]), ("e", ["a", "b", "c", "d"])]
reqs = [(pbn.CKDEType(), v, e) for v, e in fam] + [(pbn.LinearGaussianCPDType(), v, e) for v, e in fam]
sharded = pbn.CVLikelihood(data, 10, 0).local_score_batch(model, reqs)
parallel.enable(False)
single = pbn.CVLikelihood(data, 10, 0).local_score_batch(model, reqs)
parallel.enable(True)
report["cv_scores_max_rel_diff"] = max(abs(a - b) / abs(b) for a, b in zip(sharded, single))
assert report["cv_scores_max_rel_diff"] < 1e-12, report

sc = pbn.UCVScorer(train.iloc[:6000], ["a", "b", "c", "d"])
H = pbn.NormalReferenceRule().bandwidth(train.iloc[:6000], ["a", "b", "c", "d"])
u_sharded = sc.score_unconstrained(H)
parallel.enable(False)
u_single = sc.score_unconstrained(H)
parallel.enable(True)
report["ucv_rel_diff"] = abs(u_sharded - u_single) / abs(u_single)
assert report["ucv_rel_diff"] < 1e-12, report

ghc = pbn.GreedyHillClimbing()
best = ghc.estimate(pbn.OperatorPool([pbn.ArcOperatorSet(), pbn.ChangeNodeTypeSet()]), pbn.CVLikelihood(data, 10, 0), model,
                    max_indegree=3)
ops = [str(o) for o in ghc.last_run["operators"]]
parallel.enable(False)
best1 = ghc.estimate(pbn.OperatorPool([pbn.ArcOperatorSet(), pbn.ChangeNodeTypeSet()]), pbn.CVLikelihood(data, 10, 0), model,
                     max_indegree=3)
parallel.enable(True)
assert ops == [str(o) for o in ghc.last_run["operators"]] and sorted(best.arcs()) == sorted(best1.arcs())
gathered = [None] * world
dist.all_gather_object(gathered, ops)
assert all(g == ops for g in gathered)
report["hc_ops"] = len(ops)
if rank == 0:
    print("MULTI_GPU_CHECK_OK " + json.dumps(report))
dist.barrier()
dist.destroy_process_group()
