"""Throughput of KDE.logl when EVERY test row is far from the data (all unshifted sums underflow, every row takes the
shifted second pass) next to the ordinary case, and of the d = 9, 10 kernels.  Prints one JSON line per case.
usage: python tools/far_bench.py [--n 1000000] [--m 200000]"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import util_data
import pybnesian_b200 as pbn


def timed(f, reps=3):
    f()
    ctx = pbn.default_context(); ctx.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        f()
    ctx.synchronize()
    return (time.perf_counter() - t0) / reps


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--m", type=int, default=200_000)
    a = ap.parse_args()
    ctx = pbn.default_context()
    for dtype in ("float64", "float32"):
        tr = util_data.generate_normal_data(a.n, 0).astype(dtype)
        near = util_data.generate_normal_data(a.m, 1).astype(dtype)
        far = near.copy(); far["a"] += np.asarray(6.0, dtype=dtype); far["d"] -= np.asarray(40.0, dtype=dtype)
        ftr, fnear, ffar = pbn.DataFrame(tr), pbn.DataFrame(near), pbn.DataFrame(far)
        for kind in ("kde", "ckde"):
            f = pbn.KDE(["d", "a", "b", "c"]) if kind == "kde" else pbn.CKDE("d", ["a", "b", "c"])
            f.fit(ftr)
            pe = a.n * a.m * (2 if kind == "ckde" else 1)
            t_near = timed(lambda: f.slogl(fnear)); fb_near = ctx.last_fallback_rows()
            t_far = timed(lambda: f.slogl(ffar)); fb_far, rk = ctx.last_fallback_rows(), ctx.last_row_kernel_rows()
            print(json.dumps({"case": "%s d=4 %s, %d x %d" % (kind, dtype, a.n, a.m), "near_pair_evals_per_s": pe / t_near,
                              "far_pair_evals_per_s": pe / t_far, "far_over_near_time": t_far / t_near,
                              "flagged_near": fb_near, "flagged_far": fb_far, "row_kernel_rows_far": rk}), flush=True)
    for d in (8, 9, 10, 12):
        for dtype in ("float64", "float32"):
            n = min(a.n, 300_000) if d <= 10 else 100_000
            m = n if d <= 10 else 2_000
            tr = util_data.iid_normal(n, d, 0, dtype); te = util_data.iid_normal(m, d, 1, dtype)
            cols = list(tr.columns)
            ftr, fte = pbn.DataFrame(tr), pbn.DataFrame(te)
            for kind in ("kde", "ckde"):
                f = pbn.KDE(cols) if kind == "kde" else pbn.CKDE(cols[0], cols[1:])
                f.fit(ftr)
                t = timed(lambda: f.slogl(fte))
                pe = n * m * (2 if kind == "ckde" else 1)
                print(json.dumps({"case": "%s d=%d %s, %d x %d" % (kind, d, dtype, n, m), "pair_evals_per_s": pe / t,
                                  "path": "pair_kernel" if d <= 10 else "row_kernel"}), flush=True)
