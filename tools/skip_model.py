"""Planning model (CPU, numpy) for tile skipping in the pair kernel: with rows in a spatial order, which share of the
(test tile x training tile) units of BASELINE configs[1] lies entirely below the exponent floor (pair_floor in
pair_kernel.cuh) and could be skipped from the tiles' bounding boxes alone?

    python tools/skip_model.py [n_rows] [floor_bits]            config 2: CKDE d = 4 (joint and marginal)
    python tools/skip_model.py kde N d [floor_bits] [n_test]    KDE of d i.i.d. normal columns (config 5 / CV-fold shapes)

Not part of the product or the tests; DESIGN.md section 9 quotes its output."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle, util_data

KDE_MODE = len(sys.argv) > 1 and sys.argv[1] == "kde"
args = sys.argv[2:] if KDE_MODE else sys.argv[1:]
N = int(args[0]) if len(args) > 0 else 1_000_000
if KDE_MODE:
    D = int(args[1])
    BITS = float(args[2]) if len(args) > 2 else 80.0
    M = int(args[3]) if len(args) > 3 else N
    TB, TILE = (1024 if D <= 5 else 768), (512 if D <= 6 else 256)  # pair_rows / pair_tile of the f64 KDE kernels
    X = util_data.iid_normal(N, D, 0, "float64").to_numpy()
    T = util_data.iid_normal(M, D, 1, "float64").to_numpy()
else:
    BITS = float(args[1]) if len(args) > 1 else 80.0
    TB, TILE = 768, 512
    cols = ["d", "a", "b", "c"]
    X = util_data.generate_normal_data(N, seed=0)[cols].to_numpy()
    T = util_data.generate_normal_data(N, seed=1)[cols].to_numpy()
H = oracle.bandwidth(X)


def whiten(A, Hm, mu):
    return (A - mu) @ np.linalg.inv(np.linalg.cholesky(Hm)).T


def morton(Y, bits=10):
    lo, hi = Y.min(0), Y.max(0)
    q = np.clip(((Y - lo) / (hi - lo) * (1 << bits)).astype(np.int64), 0, (1 << bits) - 1)
    code = np.zeros(len(Y), dtype=np.int64)
    for b in range(bits):
        for c in range(Y.shape[1]):
            code |= ((q[:, c] >> b) & 1) << (b * Y.shape[1] + c)
    return code


def boxes(Y, size):
    n = (len(Y) + size - 1) // size
    lo = np.stack([Y[i * size:(i + 1) * size].min(0) for i in range(n)])
    hi = np.stack([Y[i * size:(i + 1) * size].max(0) for i in range(n)])
    return lo, hi


def model(name, Hm, Xc, Tc, order):
    mu = Xc.mean(0)
    Y, Z = whiten(Xc, Hm, mu), whiten(Tc, Hm, mu)
    if order == "coord0":
        Y, Z = Y[np.argsort(Y[:, 0])], Z[np.argsort(Z[:, 0])]
    elif order == "morton":
        both = np.concatenate([Y, Z])
        lo, hi = both.min(0), both.max(0)
        Y, Z = Y[np.argsort(morton((Y - lo) / (hi - lo)))], Z[np.argsort(morton((Z - lo) / (hi - lo)))]
    # log2 of the row sums of a sample of test rows (exact, against all training rows) -> floor of a test tile
    rng = np.random.default_rng(0)
    samp = rng.choice(len(Z), min(200, len(Z)), replace=False)
    l2 = []
    for z in Z[samp]:
        e = -0.5 * ((Y - z) ** 2).sum(1) * np.log2(np.e)
        l2.append(np.log2(np.exp2(e - e.max()).sum()) + e.max())
    floor = np.percentile(l2, 1) - BITS  # a tile is held back by its smallest row sum: take the 1st percentile
    tlo, thi = boxes(Z, TB)
    xlo, xhi = boxes(Y, TILE)
    alive = 0
    for i in range(len(tlo)):
        gap = np.maximum(0.0, np.maximum(xlo - thi[i], tlo[i] - xhi)).astype(np.float64)
        emax = -0.5 * (gap ** 2).sum(1) * np.log2(np.e)  # largest exponent any pair of the unit can have
        alive += int((emax >= floor).sum())
    total = len(tlo) * len(xlo)
    print("%-9s %-7s floor 2^%.0f: %5.1f%% of %d units alive -> %.2fx" % (name, order, floor, 100.0 * alive / total, total,
                                                                          total / max(alive, 1)))
    return alive, total


for order in ("none", "coord0", "morton"):
    if KDE_MODE:
        model("KDE d=%d" % D, H, X, T, order)
    else:
        model("joint d=4", H, X, T, order)
        model("marg d=3", H[1:, 1:], X[:, 1:], T[:, 1:], order)
