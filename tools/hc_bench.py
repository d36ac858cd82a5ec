"""BASELINE.json configs[3]: SemiparametricBN GreedyHillClimbing with
CVLikelihood(k=10) on a 20-node synthetic continuous data set of 100k rows.  Reports HC-CV seconds per
iteration (mean), the first-iteration cache_scores time, the operator list and the score-engine counters.
With torchrun (N ranks) the candidates of every batch are dealt over the GPUs (pybnesian_b200.parallel).

usage: python tools/hc_bench.py [--rows 100000] [--nodes 20] [--max-iters 0] [--max-indegree 4] [--json out.json]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import pandas as pd


def config4_data(rows=100_000, nodes=20, seed=0, dtype=np.float64):
    """Random DAG in index order (each node <= 3 parents among earlier nodes), weights U(-2, 2), noise sd
    U(0.5, 2); about half of the nodes get a non-linear term (sin / square of one parent)."""
    rng = np.random.default_rng(seed)
    cols = {}
    truth = []
    for i in range(nodes):
        k = int(rng.integers(0, min(i, 3) + 1))
        parents = sorted(rng.choice(i, size=k, replace=False).tolist()) if k else []
        w = rng.uniform(-2, 2, size=k)
        sd = rng.uniform(0.5, 2.0)
        x = rng.normal(0, sd, rows)
        nonlinear = bool(rng.random() < 0.5) and k > 0
        for j, p in enumerate(parents):
            z = cols["x%d" % p]
            z = (z - z.mean()) / z.std()
            if nonlinear and j == 0:
                x = x + w[j] * (np.sin(2.5 * z) if rng.random() < 0.5 else z * z)
            else:
                x = x + w[j] * z
        cols["x%d" % i] = x
        truth.append((i, parents, nonlinear))
    return pd.DataFrame(cols).astype(dtype), truth


def run(rows, nodes, max_iters, max_indegree, seed=0, k=10, verbose=0):
    import pybnesian_b200 as pbn
    data, truth = config4_data(rows, nodes, seed)
    names = list(data.columns)
    ctx = pbn.default_context()
    t0 = time.perf_counter()
    score = pbn.CVLikelihood(data, k, seed)
    pool = pbn.OperatorPool([pbn.ArcOperatorSet(), pbn.ChangeNodeTypeSet()])
    ghc = pbn.GreedyHillClimbing()
    from pybnesian_b200 import _lib
    tdev0 = time.perf_counter()
    score._scorer._device(_lib.PBN_F64)   # upload + shuffled-order table + fold statistics (pbn_cv_create)
    ctx.synchronize()
    t_device = time.perf_counter() - tdev0
    tw0 = time.perf_counter()
    warm = pbn.CVLikelihood(data.iloc[:2000], k, seed)   # loads the kernels once (CUDA lazy module loading)
    warm.local_score_node_type(pbn.SemiparametricBN(names), pbn.CKDEType(), names[0], [names[1]])
    t_warm = time.perf_counter() - tw0
    ctx.set_timing(True)
    ctx.pair_kernel_time(reset=True)
    c0 = ctx.counters()
    t1 = time.perf_counter()
    best = ghc.estimate(pool, score, pbn.SemiparametricBN(names), max_indegree=max_indegree,
                        max_iters=max_iters if max_iters > 0 else 2 ** 31 - 1, verbose=verbose)
    ctx.synchronize()
    t2 = time.perf_counter()
    kern_ms, kern_launches, kern_pairs = ctx.pair_kernel_time(reset=True)
    ctx.set_timing(False)
    c1 = ctx.counters()
    run_ = ghc.last_run
    its = run_["iteration_s"]
    return {
        "rows": rows, "nodes": nodes, "k": k, "max_indegree": max_indegree, "iterations": len(its),
        "hc_cv_s_per_iter_mean": float(np.mean(its)) if its else None,
        "hc_cv_s_per_iter_max": float(np.max(its)) if its else None,
        "cache_scores_s": run_["cache_scores_s"], "setup_s": t1 - t0, "cv_create_s": t_device, "kernel_warmup_s": t_warm, "total_s": t2 - t1,
        "pair_kernel_ms": kern_ms, "pair_kernel_launches": kern_launches, "pair_evals": kern_pairs,
        "pair_evals_per_s_in_kernel": kern_pairs / (kern_ms * 1e-3) if kern_ms else None,
        "gpu_launches": c1["launches"] - c0["launches"], "engine": dict(score._scorer.stats),
        "operators": [str(o) for o in run_["operators"]],
        "final_arcs": best.num_arcs(), "ckde_nodes": sum(1 for n in names if str(best.node_type(n)) == "CKDEFactor"),
        "final_score": score.score(best),
    }


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=100_000)
    ap.add_argument("--nodes", type=int, default=20)
    ap.add_argument("--max-iters", type=int, default=0)
    ap.add_argument("--max-indegree", type=int, default=4)
    ap.add_argument("--verbose", type=int, default=0)
    ap.add_argument("--json", default="")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        import torch
        import torch.distributed as dist
        lr = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(lr)
        os.environ["PBN_CUDA_DEVICE"] = str(lr)
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    res = run(a.rows, a.nodes, a.max_iters, a.max_indegree, verbose=a.verbose)
    res["n_gpus"] = world
    if int(os.environ.get("RANK", "0")) == 0:
        print(json.dumps(res))
        if a.json:
            with open(a.json, "w") as f:
                json.dump(res, f, indent=1)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
