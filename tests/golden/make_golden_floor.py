#!/usr/bin/env python3
"""Golden vectors for the two-cluster data of tests/test_kde_gpu.py::test_exponent_floor_is_invisible at a size the
reference kernels finish quickly: outputs of the REFERENCE'S OWN OpenCL-C kernels (oracle/_ref, see make_golden.py) for
KDE / CKDE logl on training sets made of two clusters 25 sigma apart, in three row orders, with test rows in both clusters,
between them and beyond.  Most kernel terms of every row are negligible here and the joint / marginal log-likelihoods of
the outlying rows nearly cancel, which is where a restatement (or a GPU kernel) that treats small terms differently would
show.  Run in the build container:  python tests/golden/make_golden_floor.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
import util_data  # noqa: E402

VARSETS = [["b"], ["b", "a"], ["b", "a", "c", "d"]]
ORDERS = ["near_first", "far_first", "shuffled"]


def main():
    assert oracle.ref_available(), "build oracle/_ref first (make -C oracle ref)"
    out = {}
    for dt in ("float64", "float32"):
        for order in ORDERS:
            train, test = util_data.two_cluster_frames(order)  # ~0.1 ms per pair in the kernel emulation: keep it small
            for variables in VARSETS:
                X = train[variables].to_numpy().astype(dt)
                T = test[variables].to_numpy().astype(dt)
                H = oracle.bandwidth(X)
                key = "%s_%s_%s" % (dt, order, "".join(variables))
                logl, slogl = oracle.ref_kde_logl(X, T, H)
                out["ref_kde_logl_" + key] = logl
                if len(variables) > 1:
                    cl, cs = oracle.ref_ckde_logl(X, T, H)
                    out["ref_ckde_logl_" + key] = cl
    np.savez_compressed(os.path.join(HERE, "floor_golden.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
