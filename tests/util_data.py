"""Synthetic data generators, same recipes as the reference's tests/helpers/util_test.py:5-37."""
import numpy as np
import pandas as pd


def generate_normal_data(size, seed=0):
    np.random.seed(seed)
    a = np.random.normal(3, 0.5, size=size)
    b = 2.5 + 1.65 * a + np.random.normal(0, 2, size=size)
    c = -4.2 - 1.2 * a + 3.2 * b + np.random.normal(0, 0.75, size=size)
    d = 1.5 - 0.9 * a + 5.6 * b + 0.3 * c + np.random.normal(0, 0.5, size=size)
    return pd.DataFrame({"a": a, "b": b, "c": c, "d": d})


def generate_normal_data_indep(size, seed=0):
    np.random.seed(seed)
    a = np.random.normal(3, 0.5, size=size)
    b = np.random.normal(2.5, 2, size=size)
    c = -4.2 - 1.2 * a + 3.2 * b + np.random.normal(0, 0.75, size=size)
    d = 1.5 - 0.3 * c + np.random.normal(0, 0.5, size=size)
    return pd.DataFrame({"a": a, "b": b, "c": c, "d": d})


def iid_normal(size, d, seed=0, dtype=np.float64):
    rng = np.random.default_rng(seed)
    return pd.DataFrame({"x%d" % i: rng.standard_normal(size).astype(dtype) for i in range(d)})
