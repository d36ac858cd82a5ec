import sys, time, ctypes
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, util_data
import pybnesian_b200 as pbn
from pybnesian_b200 import _lib
import torch
ctx = pbn.default_context()
for dtype in ('float64', 'float32'):
  for (kind, d, n) in [('kde', 1, 200000), ('kde', 2, 200000), ('kde', 4, 200000), ('kde', 8, 200000), ('ckde', 4, 200000), ('ckde',2,200000)]:
    tr = util_data.iid_normal(n, d, 0, dtype); te = util_data.iid_normal(n, d, 1, dtype)
    cols = list(tr.columns)
    f = pbn.KDE(cols) if kind == 'kde' else pbn.CKDE(cols[0], cols[1:])
    ftr, fte = pbn.DataFrame(tr), pbn.DataFrame(te)
    f.fit(ftr)
    f.slogl(fte)
    ctx.synchronize()
    t0 = time.perf_counter(); s = f.slogl(fte); ctx.synchronize(); t1 = time.perf_counter()
    pairs = n * n * (2 if (kind == 'ckde' and d > 1) else 1)
    print(dtype, kind, d, n, 'slogl', s, 'time %.4f s' % (t1 - t0), 'pair-evals/s %.3e' % (pairs / (t1 - t0)), flush=True)
