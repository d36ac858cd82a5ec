"""BASELINE.json configs[2]: UCV objective and UCV().bandwidth on N rows, d = 4, both dtypes.
usage: python tools/ucv_bench.py [n_rows]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, util_data
import pybnesian_b200 as pbn
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
for dt in ("float64", "float32"):
    df = pbn.DataFrame(util_data.generate_normal_data(n, 0).astype(dt))
    v = ["a", "b", "c", "d"]
    ctx = pbn.default_context()
    sc = pbn.UCVScorer(df, v)
    H = pbn.NormalReferenceRule().bandwidth(df, v)
    sc.score_unconstrained(H)
    ctx.set_timing(True); ctx.pair_kernel_time(reset=True)
    t0 = time.perf_counter(); s = sc.score_unconstrained(H); t1 = time.perf_counter()
    ms, nl, pe = ctx.pair_kernel_time(reset=True); ctx.set_timing(False)
    print(dt, "N=%d d=4 UCV objective %.10g: %.4f s wall, kernel %.2f ms, %.3e pairs/s" % (n, s, t1 - t0, ms, pe / (ms * 1e-3)), flush=True)
for dt in ("float64", "float32"):  # BASELINE configs[2]: the whole bandwidth selection (Nelder-Mead over vech(chol H))
    df = pbn.DataFrame(util_data.generate_normal_data(n, 0).astype(dt))
    t0 = time.perf_counter(); sel = pbn.UCV(); Hopt = sel.bandwidth(df, v); t1 = time.perf_counter()
    print("UCV.bandwidth (%s, N=%d, d=4): %d objective evaluations, %.2f s" % (dt, n, sel.last_evaluations, t1 - t0), flush=True)
