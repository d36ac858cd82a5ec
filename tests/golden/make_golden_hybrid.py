#!/usr/bin/env python3
"""Golden vectors for hybrid factors, same conventions as make_golden.py: per discrete
configuration, the outputs of the REFERENCE'S OWN OpenCL-C kernels run through oracle/_ref (`ref_*`) on the
rows the reference's DiscreteAdaptator would hand them (factors/discrete/DiscreteAdaptator.hpp:201-325), and the
per-configuration SciPy conditional log-density the reference's CKDE tests use (`scipy_*`,
tests/factors/continuous/CKDE_test.py:146-179).  Inputs are regenerated from seeds by tests/util_data.py.

    python tests/golden/make_golden_hybrid.py        (needs /root/reference for oracle/_ref)
"""
import os
import sys

import numpy as np
from scipy.stats import gaussian_kde

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from oracle import hybrid  # noqa: E402
import util_data  # noqa: E402

EVIDENCE = [["A", "C", "B"], ["A"], ["B", "A"], ["C", "B"]]
CASES = [(600, 80), (300, 120)]


def frames(N, m, dt):
    tr, te = util_data.generate_hybrid_data(N, 0), util_data.generate_hybrid_data(m, 1)
    for df in (tr, te):
        df["C"] = df["C"].astype(dt)
        df["D"] = df["D"].astype(dt)
    return tr, te


def scipy_conditional(X, T):
    """log p(x0 | x1..) with the normal-reference factor, built from two gaussian_kde like CKDE_test.py does;
    the marginal reuses the JOINT factor, which makes its bandwidth the sub-block H[1:,1:] (CKDE.hpp:187-199)."""
    sk = gaussian_kde(X.T, bw_method=lambda g: np.power(4 / (g.d + 2), 1 / (g.d + 4)) * g.scotts_factor())
    joint = sk.logpdf(T.T)
    if X.shape[1] == 1:
        return joint
    km = gaussian_kde(X[:, 1:].T, bw_method=sk.covariance_factor())
    return joint - km.logpdf(T[:, 1:].T)


def main():
    assert oracle.ref_available(), "build oracle/_ref first (make -C oracle ref)"
    out = {}
    for dt in ("float64", "float32"):
        for N, m in CASES:
            tr, te = frames(N, m, dt)
            for ev in EVIDENCE:
                key = "%s_%s_%d_%d" % (dt, "".join(ev), N, m)
                f = hybrid.HybridFactor("D", ev).fit(tr)
                slices = hybrid.slice_indices(te, f.discrete, f.strides, f.num_factors)
                ref = np.full(m, np.nan)
                sci = np.full(m, np.nan)
                sums = np.zeros(f.num_factors)
                for c, (fac, rows) in enumerate(zip(f.factors, slices)):
                    if fac is None or not rows:
                        continue
                    T = np.asfortranarray(te.iloc[rows][f.variables].to_numpy())
                    ll, s = oracle.ref_ckde_logl(fac[1], T, fac[2])
                    ref[rows] = ll
                    sums[c] = s
                    if dt == "float64":
                        sci[rows] = scipy_conditional(fac[1], T)
                out["ref_hckde_logl_" + key] = ref
                out["ref_hckde_sums_" + key] = sums
                if dt == "float64":
                    out["scipy_hckde_logl_" + key] = sci
    np.savez_compressed(os.path.join(HERE, "hybrid_golden.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
