"""The memory-bound helpers of the path at sizes larger than L2 (SURVEY §8d "memory-bound pieces": covariance, the
shuffled-order column store, fold statistics, whitening, LinearGaussian logl / slogl reduction).  Run under
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
        --log-file gpurun_out/hbm_kernels.csv python tools/hbm_bench.py
and summarise with tools/hbm_summary.py (algorithmic bytes / launch duration against the measured HBM peak).
Each operation is run twice; the summary uses the second launch of every kernel."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import util_data
import pybnesian_b200 as pbn
from pybnesian_b200 import _lib
from pybnesian_b200._lib import lib, check, int_array

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8_000_000
C, K = 8, 10
df = util_data.iid_normal(N, C, 0, np.float64)
frame = pbn.DataFrame(df)
names = list(df.columns)
tbl, cols, _ = frame.device_table(names)
ctx = tbl.ctx
dp = ctypes.POINTER(ctypes.c_double)
for rep in range(2):
    # covariance of 8 columns: colsum_kernel + cov_tile_kernel (3 tile pairs of 4 x 4 columns)
    mean, cov = np.empty(C), np.empty((C, C), order="F")
    check(lib().pbn_table_moments(ctx.handle, tbl.handle, int_array(cols), C, tbl.rows(), mean.ctypes.data_as(dp), cov.ctypes.data_as(dp)))
    # whitening of a d = 4 KDE (colsum for the centring + whiten_kernel with row norms)
    k = pbn.KDE(names[:4]); k.fit(frame)
    # LinearGaussianCPD logl + slogl of x0 | x1, x2, x3: lg_logl_kernel + the sum reduction
    cpd = pbn.LinearGaussianCPD(names[0], names[1:4], [0.1, 0.2, -0.3, 0.4], 1.5)
    cpd.slogl(frame)
    # cross-validation store: shuffled gather of every column + per-fold Gram statistics of all column pairs
    cv = pbn.CVLikelihood(frame, K, 0)
    sc = cv.local_score(pbn.GaussianNetwork(names), names[0], names[1:3])
    del cv, k
print(json.dumps({"N": N, "C": C, "K": K, "dtype": "float64", "check_score": sc}))
