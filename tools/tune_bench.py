"""Kernel tuning helper: pair-kernel throughput (CUDA events inside the library) for a few
shapes, for whichever build PBN_CUDA_LIB points at.  Not part of the product or the tests."""
import os, sys, ctypes, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import util_data
import pybnesian_b200 as pbn
from pybnesian_b200 import _lib

def run(kind, d, n, dtype, reps=3):
    ctx = pbn.default_context()
    ctx.set_skipping(os.environ.get("TUNE_SKIPPING", "0") == "1")   # kernel tuning: every pair evaluated unless asked
    tr = util_data.iid_normal(n, d, 0, dtype); te = util_data.iid_normal(n, d, 1, dtype)
    cols = list(tr.columns)
    f = pbn.KDE(cols) if kind == 'kde' else pbn.CKDE(cols[0], cols[1:])
    ftr, fte = pbn.DataFrame(tr), pbn.DataFrame(te)
    f.fit(ftr)
    s = f.slogl(fte)
    ctx.set_timing(True); ctx.pair_kernel_time(reset=True)
    for _ in range(reps): s = f.slogl(fte)
    ms, nl, pe = ctx.pair_kernel_time(reset=True)
    ctx.set_timing(False)
    return pe / (ms * 1e-3), s

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    n = int(os.environ.get("TUNE_N", "300000"))
    shapes = []
    if which in ("all", "f64"):
        shapes += [("ckde", 4, "float64"), ("kde", 1, "float64"), ("kde", 2, "float64"), ("kde", 4, "float64"), ("kde", 8, "float64")]
    if which in ("all", "f32"):
        shapes += [("ckde", 4, "float32"), ("kde", 1, "float32"), ("kde", 2, "float32"), ("kde", 4, "float32"), ("kde", 8, "float32")]
    if os.environ.get("TUNE_SHAPES"):  # e.g. "ckde:2:float64,kde:5:float64"
        shapes = [(k, int(d), dt) for k, d, dt in (x.split(":") for x in os.environ["TUNE_SHAPES"].split(","))]
    out = {}
    for kind, d, dt in shapes:
        v, s = run(kind, d, n, dt)
        out["%s_d%d_%s" % (kind, d, dt[-2:])] = "%.3e" % v
        out["%s_d%d_%s_slogl" % (kind, d, dt[-2:])] = "%.15g" % s
    print(os.path.basename(os.environ.get("PBN_CUDA_LIB", "default")), json.dumps(out), flush=True)
