"""Synthetic data generators, same recipes as the reference's tests/helpers/util_test.py:5-37, 99-136."""
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


def generate_hybrid_data(size, seed=0):
    """Discrete A (2 categories), B (3), continuous C, and D | A, B, C: a different linear law of C per (A, B)
    configuration (the recipe of the reference's generate_hybrid_data)."""
    np.random.seed(seed)
    a_cats = np.asarray(["a1", "a2"])
    a = a_cats[np.random.choice(a_cats.size, size, p=[0.75, 0.25])]
    b_cats = np.asarray(["b1", "b2", "b3"])
    b = b_cats[np.random.choice(b_cats.size, size, p=[0.3, 0.4, 0.3])]
    c = -4.2 + np.random.normal(0, 0.75, size=size)
    d = np.empty_like(c)
    # (A, B) -> (intercept, slope on C, noise sd); drawn in this order
    laws = [("a1", "b1", 1.0, 0.0, 0.75), ("a1", "b2", -2.0, 1.0, 2.0), ("a1", "b3", -1.0, 3.0, 0.25),
            ("a2", "b1", 2.0, 0.0, 1.0), ("a2", "b2", 3.5, -1.2, 1.0), ("a2", "b3", 4.8, -2.0, 1.5)]
    for av, bv, b0, b1, sd in laws:
        sel = np.logical_and(a == av, b == bv)
        if b1 == 0.0:
            d[sel] = np.random.normal(b0, sd, size=sel.sum())
        else:
            d[sel] = b0 + b1 * c[sel] + np.random.normal(0, sd, size=sel.sum())
    return pd.DataFrame({"A": pd.Series(a, dtype="category"), "B": pd.Series(b, dtype="category"), "C": c, "D": d})


def two_cluster_frames(order, n_half=300, m_sizes=(12, 12, 6, 6)):
    """Training set of two clusters 25 sigma apart in the given row order ("near_first", "far_first", "shuffled") and test
    rows in both clusters, between them and beyond (tests/golden/make_golden_floor.py, tests/test_oracle.py)."""
    rng = np.random.default_rng(11)
    near = generate_normal_data(n_half, seed=0)
    far = generate_normal_data(n_half, seed=2) + 25.0
    train = pd.concat([near, far] if order != "far_first" else [far, near], ignore_index=True)
    if order == "shuffled":
        train = train.iloc[rng.permutation(len(train))].reset_index(drop=True)
    test = pd.concat([generate_normal_data(m_sizes[0], seed=1), generate_normal_data(m_sizes[1], seed=3) + 25.0,
                      generate_normal_data(m_sizes[2], seed=4) + 12.5, generate_normal_data(m_sizes[3], seed=5) * 3.0],
                     ignore_index=True)
    return train, test
