"""CKDE.cdf / CKDE.sample throughput on one GPU: wall time of the public call with the training
set resident (tables cached on the DataFrame wrapper), in train x test pairs per second.
    python tools/f3_bench.py [n_train] [n_test] [--cpu]      (--cpu: also time the oracle port on a small sample)"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, util_data
import pybnesian_b200 as pbn

args = [a for a in sys.argv[1:] if not a.startswith("--")]
N = int(args[0]) if len(args) > 0 else 1000000
M = int(args[1]) if len(args) > 1 else 100000
out = {"n_train": N, "n_test": M}
QUICK = "--quick" in sys.argv  # one float64 d=4 pass (for ncu)
for dt in (("float64",) if QUICK else ("float64", "float32")):
    train = pbn.DataFrame(util_data.generate_normal_data(N, 0).astype(dt))
    test = pbn.DataFrame(util_data.generate_normal_data(M, 1).astype(dt))
    for variable, evidence in ((("d", ["a", "b", "c"]),) if QUICK else (("a", []), ("d", ["a", "b", "c"]))):
        cpd = pbn.CKDE(variable, evidence)
        cpd.fit(train)
        cpd.cdf(test)  # warm-up: uploads the test table, sets the kernel attributes
        ts = []
        for _ in range(1 if QUICK else 3):
            t0 = time.perf_counter(); c = cpd.cdf(test); ts.append(time.perf_counter() - t0)
        t = min(ts)
        key = "cdf_%s_d%d" % (dt, 1 + len(evidence))
        out[key] = {"s": t, "pairs_per_s": N * M / t, "mean_cdf": float(np.nanmean(c))}
        print(key, "%.4f s  %.3e pairs/s  mean %.6f" % (t, N * M / t, np.nanmean(c)), flush=True)
        if evidence:
            evp = util_data.generate_normal_data(M, 1).astype(dt)[evidence]
            cpd.sample(M, evp, 0)
            ts = []
            for _ in range(1 if QUICK else 3):
                t0 = time.perf_counter(); s, idx = cpd.sample(M, evp, 0, _return_indices=True); ts.append(time.perf_counter() - t0)
            t = min(ts)
            key = "sample_%s_d%d" % (dt, 1 + len(evidence))
            # pairs visited: one full pass for the totals + the scan up to the last selected row of each CTA (<= N)
            out[key] = {"s": t, "pairs_per_s_lower_bound": N * M / t, "mean_index": float(idx.mean())}
            print(key, "%.4f s  >= %.3e pairs/s  mean index %.1f" % (t, N * M / t, idx.mean()), flush=True)
if "--cpu" in sys.argv:
    import oracle
    X = util_data.generate_normal_data(N, 0)[["d", "a", "b", "c"]].to_numpy()
    T = util_data.generate_normal_data(256, 1)[["d", "a", "b", "c"]].to_numpy()
    H = oracle.bandwidth(X)
    t0 = time.perf_counter(); oracle.ckde_cdf(X, T, H); t = time.perf_counter() - t0
    out["cpu_cdf_float64_d4"] = {"s": t, "pairs_per_s": N * 256 / t, "cores": oracle.num_threads(), "sample": "256 test rows"}
    print("oracle cdf f64 d4: %.3e pairs/s on %d threads" % (N * 256 / t, oracle.num_threads()))
print(json.dumps(out))
