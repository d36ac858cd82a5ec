"""ONE process, ALL visible GPUs (no torchrun): a multi-device context (pbn_ctx_create_multi) must give the single-device
results for every sharding entry point - KDE / CKDE logl, slogl and cdf (test rows sharded), CV scores ((candidate, fold)
jobs dealt), the UCV objective (pair tiles cut per device) - and hill climbing must select the identical operator
sequence.  Also times the in-process strong scaling of one CKDE slogl.

   python tools/inproc_multi_gpu_check.py [--devices 0,1] [--json out.json]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))

import numpy as np


def run_all(pbn, tag, n_dev):
    import util_data
    from hc_bench import config4_data
    out = {}
    train, test = util_data.generate_normal_data(200_000, 0), util_data.generate_normal_data(100_003, 1)
    cpd = pbn.CKDE("d", ["a", "b", "c"])
    cpd.fit(train)
    out["ckde_slogl"], out["ckde_logl"] = cpd.slogl(test), cpd.logl(test)
    out["ckde_cdf"] = cpd.cdf(test.iloc[:40_000])
    k32 = pbn.KDE(["a", "b"])
    k32.fit(train.astype("float32"))
    out["kde32_logl"] = k32.logl(test.astype("float32"))
    far = test.iloc[:30_000].copy()
    far["a"] += 6.0
    out["ckde_far_logl"] = cpd.logl(far)            # every row takes the shifted second pass, on every device
    data, _ = config4_data(60_000, 8, 0)
    names = list(data.columns)
    model = pbn.SemiparametricBN(names)
    fam = [("x1", []), ("x2", ["x0"]), ("x3", ["x0", "x1"]), ("x5", ["x1", "x2", "x4"]), ("x7", ["x3", "x6"]),
           ("x6", ["x0", "x2", "x3", "x5"])]
    reqs = [(pbn.CKDEType(), v, e) for v, e in fam] + [(pbn.LinearGaussianCPDType(), v, e) for v, e in fam]
    out["cv_scores"] = np.array(pbn.CVLikelihood(data, 10, 0).local_score_batch(model, reqs))
    ucv_df = train.iloc[:40_000]
    H = pbn.NormalReferenceRule().bandwidth(ucv_df, ["a", "b", "c", "d"])
    sc = pbn.UCVScorer(ucv_df, ["a", "b", "c", "d"])
    out["ucv"] = np.array([sc.score_unconstrained(H), sc.score_diagonal(np.diag(H))])
    small, _ = config4_data(20_000, 8, 1)
    ghc = pbn.GreedyHillClimbing()
    best = ghc.estimate(pbn.OperatorPool([pbn.ArcOperatorSet(), pbn.ChangeNodeTypeSet()]), pbn.CVLikelihood(small, 10, 0),
                        pbn.SemiparametricBN(list(small.columns)), max_indegree=3)
    out["hc_ops"] = [str(o) for o in ghc.last_run["operators"]]
    out["hc_arcs"] = sorted(best.arcs())
    return out


def strong_scaling(pbn, n=1_000_000):
    import util_data
    train, test = pbn.DataFrame(util_data.generate_normal_data(n, 0)), pbn.DataFrame(util_data.generate_normal_data(n, 1))
    cpd = pbn.CKDE("d", ["a", "b", "c"])
    cpd.fit(train)
    s = cpd.slogl(test)
    ctx = pbn.default_context()
    ctx.synchronize()
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        s = cpd.slogl(test)
    ctx.synchronize()
    dt = (time.perf_counter() - t0) / reps
    return {"slogl": s, "ms": 1e3 * dt, "pair_evals_per_s": 2.0 * n * n / dt}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--devices", default="all")
    ap.add_argument("--json", default="")
    ap.add_argument("--no-timing", action="store_true")
    a = ap.parse_args()
    os.environ.pop("PBN_CUDA_DEVICE", None)
    import pybnesian_b200 as pbn
    from pybnesian_b200 import _lib
    import ctypes
    n = ctypes.c_int()
    _lib.check(_lib.lib().pbn_device_count(ctypes.byref(n)))
    devices = list(range(n.value)) if a.devices == "all" else [int(x) for x in a.devices.split(",")]
    report = {"devices": devices}
    pbn.set_default_context(pbn.Context(devices[0]))
    single = run_all(pbn, "single", 1)
    t1 = None if a.no_timing else strong_scaling(pbn)
    multi_ctx = pbn.set_default_context(pbn.Context(devices))
    assert multi_ctx.num_devices == len(devices)
    multi = run_all(pbn, "multi", len(devices))
    tn = None if a.no_timing else strong_scaling(pbn)
    for key in ("ckde_logl", "ckde_cdf", "kde32_logl", "ckde_far_logl", "cv_scores", "ucv"):
        x, y = np.asarray(single[key]), np.asarray(multi[key])
        # a CKDE log-likelihood is a difference of two log-sums and crosses zero: relative to max(|value|, 1)
        err = float(np.max(np.abs(x - y) / np.maximum(np.abs(x), 1.0)))
        report[key + "_max_rel_diff"] = err
        # sharding changes where the partial sums of a row are cut (unit schedule), never what is summed
        assert err < (1e-12 if key != "kde32_logl" else 1e-6), (key, err)
    report["slogl_rel_diff"] = abs(single["ckde_slogl"] - multi["ckde_slogl"]) / abs(single["ckde_slogl"])
    assert report["slogl_rel_diff"] < 1e-12
    assert single["hc_ops"] == multi["hc_ops"] and single["hc_arcs"] == multi["hc_arcs"], (single["hc_ops"], multi["hc_ops"])
    report["hc_ops"] = len(single["hc_ops"])
    c = multi_ctx.counters()
    report["multi_ctx_counters"] = c
    if t1 and tn:
        report["strong_scaling_1M_x_1M"] = {"single": t1, "multi": tn, "speedup": t1["ms"] / tn["ms"],
                                            "efficiency": t1["ms"] / tn["ms"] / len(devices)}
        assert abs(t1["slogl"] - tn["slogl"]) <= 1e-12 * abs(t1["slogl"])
    print("INPROC_MULTI_GPU_OK " + json.dumps(report))
    if a.json:
        with open(a.json, "w") as f:
            json.dump(report, f, indent=1)
