"""CPU oracle for hybrid factors — TEST INFRASTRUCTURE ONLY.

Plain restatement of the reference's DiscreteAdaptator logic on pandas frames:
  factors/discrete/discrete_indices.cpp:166-201  (discrete_slice_indices: one ascending row-id list per
                                                   configuration of the discrete parents)
  factors/discrete/DiscreteAdaptator.hpp:201-325 (fit / logl / slogl: base factor per configuration, run on the
                                                   `take` of each configuration's rows)
  factors/continuous/CKDE.hpp:747-770            (CKDEFitter: SingularCovarianceData -> configuration unfitted)
  factors/continuous/LinearGaussianCPD.hpp:127-138 (LinearGaussianFitter: variance < machine_tol or inf -> unfitted)
  learning/parameters/mle_DiscreteFactor.cpp:5-40 (DiscreteFactor MLE)
The per-configuration arithmetic is the oracle's CKDE / LinearGaussianCPD port (pinned bit for bit against the
reference's own kernels, tests/test_oracle.py).  **Parity of the adaptor itself is unpinned**: the reference has
no test that checks HCKDE / CLinearGaussianCPD numbers; the tests compare with per-configuration SciPy results.
"""
import numpy as np

import oracle

MACHINE_TOL = 1.4901161193847656e-08  # util/math_constants.hpp:30


def cardinality_strides(df, discrete):
    """discrete_indices.cpp:118-140."""
    card = [len(df[v].cat.categories) for v in discrete]
    strides = [1] * len(discrete)
    for i in range(1, len(discrete)):
        strides[i] = strides[i - 1] * card[i - 1]
    return card, strides


def slice_indices(df, discrete, strides, num_factors):
    """discrete_indices.cpp:166-201, as written: push each valid row id onto its configuration's list."""
    codes = [df[v].cat.codes.to_numpy() for v in discrete]  # -1 marks a null
    slices = [[] for _ in range(num_factors)]
    for r in range(len(df)):
        if any(c[r] < 0 for c in codes):
            continue
        idx = 0
        for c, s in zip(codes, strides):
            idx += int(c[r]) * s
        slices[idx].append(r)
    return slices


def _split_evidence(df, evidence):
    discrete = [e for e in evidence if str(df[e].dtype) == "category"]
    continuous = [e for e in evidence if str(df[e].dtype) != "category"]
    return discrete, continuous


def _dense(df, rows, variables):
    sub = df.iloc[rows][variables]
    keep = ~sub.isna().any(axis=1).to_numpy()
    return np.asfortranarray(sub.to_numpy()[keep]), keep


class HybridFactor:
    """DiscreteAdaptator<CKDE> (kind='ckde') or DiscreteAdaptator<LinearGaussianCPD> (kind='lg')."""

    def __init__(self, variable, evidence, kind="ckde", rule="normal_reference"):
        self.variable, self.evidence, self.kind, self.rule = variable, list(evidence), kind, rule

    def fit(self, df):
        self.discrete, self.continuous = _split_evidence(df, self.evidence)
        self.variables = [self.variable] + self.continuous
        self.card, self.strides = cardinality_strides(df, self.discrete)
        self.num_factors = int(np.prod(self.card)) if self.discrete else 1
        slices = slice_indices(df, self.discrete, self.strides, self.num_factors) if self.discrete else [list(range(len(df)))]
        self.factors = []
        for rows in slices:
            if not rows:
                self.factors.append(None)
                continue
            X, _ = _dense(df, rows, self.variables)
            self.factors.append(self._fit_base(X))
        return self

    def _fit_base(self, X):
        d = X.shape[1]
        if self.kind == "ckde":
            try:
                H = oracle.bandwidth(X, self.rule)
            except oracle.SingularCovariance:
                return None
            return ("ckde", X, H)
        beta, var = oracle.lg_fit(X[:, 0], [X[:, j] for j in range(1, d)])
        if var < MACHINE_TOL or np.isinf(var):
            return None
        return ("lg", beta, var)

    def _eval_base(self, f, T):
        if f[0] == "ckde":
            return oracle.ckde_logl(f[1], T, f[2])[0] if T.shape[1] > 1 else oracle.kde_logl(f[1], T, f[2])[0]
        return oracle.lg_logl(T[:, 0], [T[:, j] for j in range(1, T.shape[1])], f[1], f[2])[0]

    def logl(self, df):
        out = np.full(len(df), np.nan)
        slices = slice_indices(df, self.discrete, self.strides, self.num_factors) if self.discrete else [list(range(len(df)))]
        for f, rows in zip(self.factors, slices):
            if not rows or f is None:
                continue
            T, keep = _dense(df, rows, self.variables)
            if T.shape[0]:
                out[np.asarray(rows)[keep]] = self._eval_base(f, T)
        return out

    def slogl(self, df):
        """Sum of the per-configuration sums, in configuration order (DiscreteAdaptator.hpp:315-320)."""
        res = 0.0
        slices = slice_indices(df, self.discrete, self.strides, self.num_factors) if self.discrete else [list(range(len(df)))]
        for f, rows in zip(self.factors, slices):
            if not rows or f is None:
                continue
            T, _ = _dense(df, rows, self.variables)
            if not T.shape[0]:
                continue
            if f[0] == "ckde":
                res += (oracle.ckde_logl(f[1], T, f[2])[1] if T.shape[1] > 1 else oracle.kde_logl(f[1], T, f[2])[1])
            else:
                res += oracle.lg_logl(T[:, 0], [T[:, j] for j in range(1, T.shape[1])], f[1], f[2])[1]
        return res


def discrete_factor_logprob(df, variable, evidence):
    """mle_DiscreteFactor.cpp:5-40: log P(variable | evidence) as a flat table, variable fastest."""
    names = [variable] + list(evidence)
    card, strides = cardinality_strides(df, names)
    counts = np.zeros(int(np.prod(card)), dtype=np.int64)
    codes = [df[v].cat.codes.to_numpy() for v in names]
    for r in range(len(df)):
        if any(c[r] < 0 for c in codes):
            continue
        counts[sum(int(c[r]) * s for c, s in zip(codes, strides))] += 1
    logprob = np.empty(counts.size)
    c0 = card[0]
    for k in range(counts.size // c0):
        tot = counts[k * c0:(k + 1) * c0].sum()
        for i in range(c0):
            with np.errstate(divide="ignore"):
                logprob[k * c0 + i] = np.log(1.0 / c0) if tot == 0 else np.log(float(counts[k * c0 + i])) - np.log(float(tot))
    return logprob, card, strides
