"""Small-size run of every kernel added in round 2 for compute-sanitizer (memcheck / racecheck are 10-50x slower, so the
size thresholds of tile skipping are lowered through PBN_SKIP_MIN_TRAIN / PBN_SKIP_MIN_TEST):
   PBN_SKIP_MIN_TRAIN=1000 PBN_SKIP_MIN_TEST=500 compute-sanitizer --tool memcheck python tools/memcheck_r2.py
Checks the values against the oracle as well."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle, util_data
import pybnesian_b200 as pbn

ctx = pbn.default_context()
ctx.warmup()
n, m = 6011, 2503
for dtype, tol in (("float64", 1e-10), ("float32", 1e-4)):
    tr = util_data.generate_normal_data(n, 0).astype(dtype)
    te = util_data.generate_normal_data(m, 1)
    te.loc[:199, "a"] += 6.0                    # far rows: shifted second pass, in Morton order
    te.loc[200:209, "a"] += 1.0e4               # beyond the shift: per-row kernel
    te = te.astype(dtype)
    for kind, variables in (("kde", ["a"]), ("kde", ["b", "a"]), ("ckde", ["c", "a", "b"]), ("ckde", ["d", "a", "b", "c"])):
        f = pbn.KDE(variables) if kind == "kde" else pbn.CKDE(variables[0], variables[1:])
        f.fit(tr)
        for skipping in (True, False):
            ctx.set_skipping(skipping)
            got = f.logl(te)
            st = ctx.skip_stats()
            X, T = tr[variables].to_numpy().astype(np.float64), te[variables].to_numpy().astype(np.float64)
            H = np.asarray(f.bandwidth if kind == "kde" else f.kde_joint().bandwidth, dtype=np.float64)
            want = (oracle.kde_logl if kind == "kde" else oracle.ckde_logl)(X, T, H)[0]
            err = np.max(np.abs(got - want) / np.maximum(np.abs(want), 1.0))
            print(dtype, kind, variables, "skipping", skipping, "units", st["last_evaluated"], "/", st["last_total"], "fallback",
                  ctx.last_fallback_rows(), ctx.last_row_kernel_rows(), "err %.2e" % err, flush=True)
            assert err < tol, err
            assert abs(f.slogl(te) - got.sum()) <= 1e-11 * abs(got.sum())
    # d = 9, 10 and a wide family
    for d in (9, 10, 12):
        w = util_data.iid_normal(3000, d, 0, dtype); wt = util_data.iid_normal(700, d, 1, dtype)
        k = pbn.KDE(list(w.columns)); k.fit(w)
        got = k.logl(wt)
        want = oracle.kde_logl(w.to_numpy().astype(np.float64), wt.to_numpy().astype(np.float64), np.asarray(k.bandwidth))[0]
        assert np.max(np.abs(got - want) / np.abs(want)) < (1e-9 if dtype == "float64" else 1e-4)
ctx.set_skipping(True)
# enough tiles in one and two dimensions for list B (units dropped, explicit unit list, pass B on top of pass A)
for dtype, tol in (("float64", 1e-10), ("float32", 1e-4)):
    big = util_data.generate_normal_data(60_011, 0).astype(dtype)
    bt = util_data.generate_normal_data(20_003, 1).astype(dtype)
    for kind, variables in (("kde", ["a"]), ("ckde", ["b", "a"]), ("kde", ["b", "a"])):
        f = pbn.KDE(variables) if kind == "kde" else pbn.CKDE(variables[0], variables[1:])
        f.fit(big)
        got = f.logl(bt)
        st = ctx.skip_stats()
        assert st["last_evaluated"] <= st["last_total"], st   # (2-d float32 at this size: nothing to drop, falls back)
        rows = np.arange(0, 20_003, 97)
        X, T = big[variables].to_numpy().astype(np.float64), bt[variables].to_numpy().astype(np.float64)[rows]
        H = np.asarray(f.bandwidth if kind == "kde" else f.kde_joint().bandwidth, dtype=np.float64)
        want = (oracle.kde_logl if kind == "kde" else oracle.ckde_logl)(X, T, H)[0]
        err = np.max(np.abs(got[rows] - want) / np.maximum(np.abs(want), 1.0))
        print(dtype, kind, variables, "list B: units", st["last_evaluated"], "/", st["last_total"], "err %.2e" % err, flush=True)
        assert err < tol, err
# batched scores (job lists) and UCV
data = util_data.generate_normal_data(3000, 0)
cv = pbn.CVLikelihood(data, 5, 0)
model = pbn.SemiparametricBN(list(data.columns))
reqs = [(pbn.CKDEType(), "d", ["a", "b"]), (pbn.CKDEType(), "a", []), (pbn.LinearGaussianCPDType(), "c", ["a"])]
print("cv", cv.local_score_batch(model, reqs))
sc = pbn.UCVScorer(data, ["a", "b"])
print("ucv", sc.score_unconstrained(np.asarray(pbn.NormalReferenceRule().bandwidth(data, ["a", "b"]))))
print("MEMCHECK_RUN_OK")
