"""Factors on the hot path: the Factor / FactorType plugin interface and CKDE.

Mirrors factors/factors.hpp:28-198 (FactorType, Factor), factors/continuous/CKDE.{hpp,cpp}
and pybindings_factors.cpp:314-470, 574-640.  CKDE = joint KDE - marginal KDE; here both
are evaluated by ONE fused kernel pass (include/pbn_cuda.h: pbn_ckde_fit / pbn_kde_logl).
"""
import pickle

import numpy as np

from . import _lib
from .dataset import DataFrame
from .kde import KDE, BandwidthSelector, NormalReferenceRule, _ARROW_TYPE, _fit_handle, _run_logl
from ._lib import lib


def _random_seed(seed):
    """util::random_seed_arg (util/util_types.hpp:52-66): an unsigned 32-bit seed, std::random_device when absent."""
    if seed is None:
        import os
        return int.from_bytes(os.urandom(4), "little")
    return int(seed) & 0xFFFFFFFF


class FactorType:
    """factors/factors.hpp:28-101.  Singletons compared by identity/hash like the reference."""

    _instances = {}

    def __new__(cls, *args, **kwargs):
        inst = FactorType._instances.get(cls)
        if inst is None:
            inst = super().__new__(cls)
            FactorType._instances[cls] = inst
        return inst

    def new_factor(self, model, variable, evidence, *args, **kwargs):
        # pybind11's message for an un-overridden pure virtual of the trampoline (a RuntimeError in the reference;
        # NotImplementedError is one too)
        raise NotImplementedError('Tried to call pure virtual function "FactorType::new_factor"')

    def __eq__(self, other):
        return type(self) is type(other)

    def __ne__(self, other):
        return not self == other

    def __hash__(self):
        return hash(type(self).__name__)

    def __str__(self):
        return type(self).__name__

    __repr__ = __str__

    def __reduce__(self):
        return (type(self), ())


class Factor:
    """factors/factors.hpp:118-198."""

    def __init__(self, variable, evidence):
        self._variable = variable
        self._evidence = list(evidence)

    def variable(self):
        return self._variable

    def evidence(self):
        return list(self._evidence)

    def fitted(self):
        raise NotImplementedError

    def type(self):
        raise NotImplementedError

    def save(self, filename):
        if not filename.endswith(".pickle"):
            filename += ".pickle"
        with open(filename, "wb") as f:
            pickle.dump(self, f)


class UnknownFactorType(FactorType):
    """factors/factors.hpp:103-116: placeholder type of a node whose CPD type is not decided yet."""

    def new_factor(self, model, variable, evidence, *args, **kwargs):
        raise ValueError("UnknownFactorType cannot create a new Factor.")

    def __str__(self):
        return "UnknownFactorType"

    __repr__ = __str__


class LinearGaussianCPDType(FactorType):
    """factors/continuous/LinearGaussianCPD.hpp:18-58."""

    def new_factor(self, model, variable, evidence, *args, **kwargs):
        # LinearGaussianCPD.cpp:33-57: a discrete parent makes it a conditional linear Gaussian
        from . import hybrid
        if any(model.node_type(e) == hybrid.DiscreteFactorType() for e in evidence):
            return hybrid.CLinearGaussianCPD(variable, evidence, *args, **kwargs)
        return LinearGaussianCPD(variable, evidence, *args, **kwargs)

    def __str__(self):
        return "LinearGaussianFactor"

    __repr__ = __str__


class LinearGaussianCPD(Factor):
    """pybnesian.LinearGaussianCPD (factors/continuous/LinearGaussianCPD.{hpp,cpp},
    learning/parameters/mle_LinearGaussianCPD.hpp): y ~ N(beta0 + beta . evidence, variance)."""

    def __init__(self, variable, evidence, beta=None, variance=None):
        super().__init__(variable, evidence)
        self._variables = [variable] + list(evidence)
        self._fitted = False
        self._beta = np.empty(0)
        self._variance = -1.0
        if beta is not None:
            beta = np.asarray(beta, dtype=np.float64).ravel()
            if beta.size != len(self._evidence) + 1:
                raise ValueError("Wrong number of beta parameters. Beta vector size: %d. Expected beta vector size: %d."
                                 % (beta.size, len(self._evidence) + 1))
            if variance is None or variance <= 0:
                raise ValueError("Variance must be a positive value.")
            self._beta, self._variance, self._fitted = beta.copy(), float(variance), True

    def type(self):
        return LinearGaussianCPDType()

    def fitted(self):
        return self._fitted

    def data_type(self):
        import pyarrow as pa
        return pa.float64()

    @property
    def beta(self):
        return self._beta

    @beta.setter
    def beta(self, value):
        value = np.asarray(value, dtype=np.float64).ravel()
        if value.size != len(self._evidence) + 1:
            raise ValueError("Wrong number of beta parameters.")
        self._beta = value.copy()
        if self._variance > 0:
            self._fitted = True

    @property
    def variance(self):
        return self._variance

    @variance.setter
    def variance(self, value):
        if value <= 0:
            raise ValueError("Variance must be a positive value.")
        self._variance = float(value)
        if self._beta.size == len(self._evidence) + 1:
            self._fitted = True

    def _check_fitted(self):
        if not self._fitted:
            raise ValueError("LinearGaussianCPD factor not fitted.")

    def fit(self, df):
        import ctypes
        from ._lib import check, int_array
        frame = DataFrame.wrap(df)
        frame.dtype_code(self._variables, "fit LinearGaussianCPD")
        tbl, cols, _ = frame.device_table(self._variables)
        beta = np.empty(len(cols))
        var = ctypes.c_double()
        check(lib().pbn_lg_fit(tbl.ctx.handle, tbl.handle, int_array(cols), len(cols), tbl.rows(),
                               beta.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), ctypes.byref(var)))
        self._beta, self._variance, self._fitted = beta, var.value, True

    def _eval(self, df, want_logl):
        import ctypes
        from ._lib import check, int_array
        self._check_fitted()
        frame = DataFrame.wrap(df)
        frame.dtype_code(self._variables, "compute logl")
        tbl, cols, mask = frame.device_table(self._variables)
        out = np.empty(tbl.nrows) if want_logl else None
        s = ctypes.c_double(0.0)
        dp = ctypes.POINTER(ctypes.c_double)
        check(lib().pbn_lg_logl(tbl.ctx.handle, tbl.handle, int_array(cols), len(cols), tbl.rows(),
                                np.ascontiguousarray(self._beta).ctypes.data_as(dp), float(self._variance),
                                out.ctypes.data_as(dp) if want_logl else None, None if want_logl else ctypes.byref(s)))
        if want_logl and mask is not None:
            full = np.full(frame.num_rows, np.nan)
            full[mask] = out
            out = full
        return out, s.value

    def logl(self, df):
        return self._eval(df, True)[0]

    def slogl(self, df):
        return self._eval(df, False)[1]

    def cdf(self, df):
        """LinearGaussianCPD::cdf (LinearGaussianCPD.cpp:171-315): host arithmetic in the data's dtype,
        0.5 erfc((mean - x) / (sqrt(2) sigma)); NaN at rows with a null."""
        from scipy.special import erfc
        self._check_fitted()
        frame = DataFrame.wrap(df)
        code = frame.dtype_code(self._variables, "compute cdf")
        T = np.float64 if code == _lib.PBN_F64 else np.float32
        cols, mask = frame.dense_columns(self._variables)
        x = np.asarray(cols[0], dtype=T)
        means = np.full(x.shape[0], T(self._beta[0]), dtype=T)
        for j, c in enumerate(cols[1:], start=1):
            means += T(self._beta[j]) * np.asarray(c, dtype=T)
        inv_std = T(1.0 / np.sqrt(self._variance))
        t = (0.5 * erfc((means - x) * inv_std * T(0.70710678118654752440))).astype(np.float64)
        if mask is not None:
            full = np.full(frame.num_rows, np.nan)
            full[mask] = t
            t = full
        return t

    def sample(self, n, evidence_values=None, seed=None):
        """LinearGaussianCPD::sample (LinearGaussianCPD.cpp:317-372): always float64; the random stream is the
        reference's (include/pbn_cuda.h: pbn_lg_sample)."""
        import ctypes
        import pyarrow as pa
        from ._lib import check
        if n < 0:
            raise ValueError("n should be a non-negative number")
        self._check_fitted()
        seed = _random_seed(seed)
        cols, code = [], _lib.PBN_F64
        if self._evidence:
            frame = DataFrame.wrap(evidence_values) if evidence_values is not None else None
            if frame is None or not frame.has_columns(self._evidence):
                raise ValueError("Evidence values not present for sampling.")
            code = frame.dtype_code(self._evidence, "sample")
            cols = [np.ascontiguousarray(frame.column_numpy(e)[:n]) for e in self._evidence]
        out = np.empty(n)
        dp = ctypes.POINTER(ctypes.c_double)
        ptrs = (ctypes.c_void_p * max(1, len(cols)))(*[c.ctypes.data for c in cols])
        check(lib().pbn_lg_sample(np.ascontiguousarray(self._beta).ctypes.data_as(dp), float(self._variance), len(cols),
                                  ptrs, code, n, seed, out.ctypes.data_as(dp)))
        return pa.array(out)

    def __getstate__(self):
        return (self._variable, self._evidence, self._fitted, self._beta, self._variance)

    def __setstate__(self, t):
        self.__init__(t[0], t[1])
        if t[2]:
            self._beta, self._variance, self._fitted = np.asarray(t[3], dtype=np.float64), float(t[4]), True

    def __str__(self):
        if not self._fitted:
            return "[LinearGaussianCPD] P(" + self._variable + (" | " + ", ".join(self._evidence) if self._evidence else "") + ") not fitted."
        terms = "%.3f" % self._beta[0] + "".join(" + %.3f*%s" % (b, e) for b, e in zip(self._beta[1:], self._evidence))
        return "[LinearGaussianCPD] P(" + self._variable + (" | " + ", ".join(self._evidence) if self._evidence else "") + \
            ") = N(" + terms + ", %.3f)" % self._variance

    __repr__ = __str__


class LinearGaussianParams:
    """pybnesian.LinearGaussianParams (pybindings_parameters.cpp:43-62): beta (intercept first) and variance."""

    def __init__(self, beta, variance):
        self.beta = np.asarray(beta, dtype=np.float64).ravel().copy()
        self.variance = float(variance)


class MLELinearGaussianCPD:
    """MLE<LinearGaussianCPD> (learning/parameters/mle_LinearGaussianCPD.hpp:11-221) on the resident table
    (include/pbn_cuda.h: pbn_lg_fit).  Created with MLE(LinearGaussianCPDType())."""

    def estimate(self, df, variable, evidence):
        import ctypes
        from ._lib import check, int_array
        frame = DataFrame.wrap(df)
        variable = frame.column_name(variable)
        evidence = [frame.column_name(e) for e in evidence]
        variables = [variable] + evidence
        frame.dtype_code(variables, "fit LinearGaussianCPD")
        tbl, cols, _ = frame.device_table(variables)
        beta = np.empty(len(cols))
        var = ctypes.c_double()
        check(lib().pbn_lg_fit(tbl.ctx.handle, tbl.handle, int_array(cols), len(cols), tbl.rows(),
                               beta.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), ctypes.byref(var)))
        return LinearGaussianParams(beta, var.value)


def MLE(factor_type):
    """pybnesian.MLE (pybindings_parameters.cpp:31-40, pybindings_mle.cpp): only the factor types with a closed-form
    estimator have one; the CKDE is fitted, not estimated."""
    from . import hybrid
    if factor_type == LinearGaussianCPDType():
        return MLELinearGaussianCPD()
    if factor_type == hybrid.DiscreteFactorType():
        return hybrid.MLEDiscreteFactor()
    raise ValueError("MLE not available for factor type " + str(factor_type) + ".")


class CKDEType(FactorType):
    """factors/continuous/CKDE.hpp:17-60, CKDE.cpp:15-41 (a discrete parent makes it an HCKDE)."""

    def new_factor(self, model, variable, evidence, *args, **kwargs):
        from . import hybrid
        if any(model.node_type(e) == hybrid.DiscreteFactorType() for e in evidence):
            return hybrid.HCKDE(variable, evidence, *args, **kwargs)
        return CKDE(variable, evidence, *args, **kwargs)

    def __str__(self):
        return "CKDEFactor"

    __repr__ = __str__


class CKDE(Factor):
    """pybnesian.CKDE (factors/continuous/CKDE.hpp:62-287)."""

    def __init__(self, variable, evidence, bandwidth_selector=None):
        super().__init__(variable, evidence)
        if bandwidth_selector is None:
            bandwidth_selector = NormalReferenceRule()
        if not isinstance(bandwidth_selector, BandwidthSelector):
            raise RuntimeError("Bandwidth selector procedure must be non-null.")
        self._variables = [variable] + list(evidence)
        self._bselector = bandwidth_selector
        self._fitted = False
        self._handle = None
        self._train = None
        self._bandwidth = None
        self._N = 0
        self._dtype = _lib.PBN_F64
        self._joint = None
        self._marg = None

    def type(self):
        return CKDEType()

    def fitted(self):
        return self._fitted

    def _check_fitted(self):
        if not self._fitted:
            raise ValueError("CKDE factor not fitted.")

    def data_type(self):
        self._check_fitted()
        return _ARROW_TYPE[self._dtype]

    def num_instances(self):
        self._check_fitted()
        return self._N

    def bandwidth_type(self):
        return self._bselector

    def fit(self, df):
        frame = DataFrame.wrap(df)
        self._dtype = frame.dtype_code(self._variables, "fit KDE")
        H = np.asarray(self._bselector.bandwidth(frame, self._variables), dtype=np.float64)
        tbl, cols, _ = frame.device_table(self._variables)
        self._fit_table(tbl, cols, tbl.rows(), H)

    def _fit_table(self, tbl, cols, rows, H):
        self._handle = _fit_handle(tbl, cols, rows, H, ckde=True)
        self._train = (tbl, list(cols), rows)
        self._bandwidth = np.array(H)
        self._N = int(lib().pbn_kde_num_instances(self._handle.handle))
        self._dtype = tbl.dtype_code
        self._joint = self._marg = None
        self._fitted = True

    def _check_test(self, frame):
        self._check_fitted()
        if frame.same_type(self._variables) != _ARROW_TYPE[self._dtype]:
            raise ValueError("Data type of training and test datasets is different.")

    def logl(self, df):
        frame = DataFrame.wrap(df)
        self._check_test(frame)
        return _run_logl(self._handle, frame, self._variables, True, False)[0]

    def slogl(self, df):
        frame = DataFrame.wrap(df)
        self._check_test(frame)
        return _run_logl(self._handle, frame, self._variables, False, True)[1]

    def cdf(self, df):
        """CKDE::cdf (CKDE.cpp:123-140, CKDE.hpp:506-728): P(variable <= x | evidence) of every row, NaN at rows
        with a null; one fused launch over train x test tiles (include/pbn_cuda.h: pbn_ckde_cdf)."""
        import ctypes
        from ._lib import check, int_array
        frame = DataFrame.wrap(df)
        self._check_test(frame)
        from . import parallel
        tbl, cols, mask = frame.device_table(self._variables)
        m = tbl.nrows
        # several ranks: contiguous test-row shards against the replicated training set, zero-padded vector
        # all-reduce (each element written by exactly one rank), as _run_logl does
        b, e = parallel.shard_range(m) if parallel.active() else (0, m)
        out = np.zeros(m)
        with parallel.guard() as g:
            if e > b:
                shard = out[b:e]
                check(lib().pbn_ckde_cdf(tbl.ctx.handle, self._handle.handle, tbl.handle, int_array(cols), tbl.rows(b, e),
                                         shard.ctypes.data_as(ctypes.POINTER(ctypes.c_double))))
        if parallel.active():
            out = parallel.all_reduce_sum(out, tbl.ctx, error=g.error)
        if mask is not None:
            full = np.full(frame.num_rows, np.nan)
            full[mask] = out
            out = full
        return out

    def sample(self, n, evidence_values=None, seed=None, _return_indices=False):
        """CKDE::sample (CKDE.cpp:97-121, CKDE.hpp:289-504): a training instance is drawn with probability
        proportional to its marginal kernel weight at the evidence (uniformly without evidence), then the variable is
        drawn from the conditional Gaussian of that kernel.  Index selection runs on the device
        (pbn_ckde_sample_indices); the random streams are the reference's.  Returns a pyarrow array of the training
        dtype."""
        import ctypes
        import pyarrow as pa
        from ._lib import check, int_array
        from .dataset import DeviceTable, _NP_DTYPE
        if n < 0:
            raise ValueError("n should be a non-negative number")
        self._check_fitted()
        seed = _random_seed(seed)
        npdt = _NP_DTYPE[self._dtype]
        tbl, cols, rows = self._train
        ev_tbl, ev_cols, ev_host, ev_ptrs = None, None, [], None
        if self._evidence:
            frame = DataFrame.wrap(evidence_values) if evidence_values is not None else None
            if frame is None or not frame.has_columns(self._evidence):
                raise ValueError("Evidence values not present for sampling.")
            t = frame.same_type(self._evidence)
            if t != _ARROW_TYPE[self._dtype]:
                raise ValueError("Data type of evidence values (%s) is different from CKDE training data (%s)."
                                 % (t, _ARROW_TYPE[self._dtype]))
            if frame.num_rows < n:
                raise ValueError("Evidence values not present for sampling.")
            ev_host = [np.ascontiguousarray(frame.column_numpy(e)[:n], dtype=npdt) for e in self._evidence]
            if n > 0:
                ev_tbl = DeviceTable(tbl.ctx, ev_host, self._dtype)
                ev_cols = int_array(list(range(len(ev_host))))
            ev_ptrs = (ctypes.c_void_p * len(ev_host))(*[c.ctypes.data for c in ev_host])
        out = np.empty(n, dtype=npdt)
        idx = np.empty(n, dtype=np.int32)
        H = np.asfortranarray(self._bandwidth, dtype=np.float64)
        check(lib().pbn_ckde_sample(tbl.ctx.handle, self._handle.handle, H.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                    tbl.handle, int_array(cols), rows, ev_tbl.handle if ev_tbl is not None else None,
                                    ev_cols, ev_ptrs, n, seed, out.ctypes.data_as(ctypes.c_void_p),
                                    idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))))
        arr = pa.array(out)
        return (arr, idx) if _return_indices else arr

    def kde_joint(self):
        """The joint KDE over [variable] + evidence (CKDE.hpp:105-108)."""
        self._check_fitted()
        if self._joint is None:
            tbl, cols, rows = self._train
            k = KDE(self._variables, self._bselector)
            k._fit_table(tbl, cols, rows, self._bandwidth)
            self._joint = k
        return self._joint

    def kde_marg(self):
        """The marginal KDE over the evidence with H[1:,1:] (CKDE.hpp:109-112, 187-199)."""
        self._check_fitted()
        if self._marg is None and self._evidence:
            tbl, cols, rows = self._train
            k = KDE(self._evidence, self._bselector)
            k._fit_table(tbl, cols[1:], rows, self._bandwidth[1:, 1:])
            self._marg = k
        elif self._marg is None:
            self._marg = KDE.__new__(KDE)
            KDE.__init__(self._marg, ["_"], self._bselector)
            self._marg._variables = []
        return self._marg

    # pickle: CKDE::__getstate__ (CKDE.cpp:164-218) stores the joint KDE and rebuilds the marginal
    def __getstate__(self):
        joint = self.kde_joint().__getstate__() if self._fitted else None
        return (self._variable, self._evidence, self._fitted, self._bselector, joint)

    def __setstate__(self, t):
        self.__init__(t[0], t[1], t[3])
        if t[2]:
            k = KDE.__new__(KDE)
            k.__setstate__(t[4])
            tbl, cols, rows = k._train
            self._fit_table(tbl, cols, rows, k.bandwidth)

    def __str__(self):
        if self._evidence:
            return "[CKDE] P(" + self._variable + " | " + ", ".join(self._evidence) + ")" + (
                " with %d instances" % self._N if self._fitted else " not fitted")
        return "[CKDE] P(" + self._variable + ")" + (" with %d instances" % self._N if self._fitted else " not fitted")

    __repr__ = __str__
