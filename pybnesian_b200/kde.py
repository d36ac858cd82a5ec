"""KDE and the bandwidth selectors — Python surface of pybnesian's `kde` package.

Mirrors pybindings_kde.cpp:116-302 (BandwidthSelector, ScottsBandwidth,
NormalReferenceRule, KDE) with the same names, argument meaning and error behaviour;
all arithmetic runs in libpbn_cuda.so (C ABI, include/pbn_cuda.h).
"""
import ctypes
import pickle

import numpy as np
import pyarrow as pa

from . import _lib
from ._lib import Rows, check, lib, int_array
from .dataset import DataFrame

_ARROW_TYPE = {_lib.PBN_F64: pa.float64(), _lib.PBN_F32: pa.float32()}


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


class BandwidthSelector:
    """Subclassable in Python like the reference's trampoline (pybindings_kde.cpp:19-154)."""

    def diag_bandwidth(self, df, variables):
        raise NotImplementedError("BandwidthSelector::diag_bandwidth is a pure virtual function")

    def bandwidth(self, df, variables):
        raise NotImplementedError("BandwidthSelector::bandwidth is a pure virtual function")

    def __str__(self):
        return "BandwidthSelector"

    __repr__ = __str__

    def __getstate_extra__(self):
        return ()

    def __setstate_extra__(self, extra):
        pass


class _NativeSelector(BandwidthSelector):
    _rule = None
    _name = None

    def bandwidth(self, df, variables):
        variables = list(variables)
        if not variables:
            return np.empty((0, 0))
        frame = DataFrame.wrap(df)
        frame.dtype_code(variables, "fit bandwidth")
        tbl, cols, _ = frame.device_table(variables)
        return self._bandwidth_rows(tbl, cols, tbl.rows())

    def _bandwidth_rows(self, tbl, cols, rows):
        d = len(cols)
        H = np.empty((d, d), order="F")
        check(lib().pbn_bandwidth(tbl.ctx.handle, tbl.handle, int_array(cols), d, rows, self._rule, _dp(H)))
        return H

    def diag_bandwidth(self, df, variables):
        variables = list(variables)
        if not variables:
            return np.empty(0)
        frame = DataFrame.wrap(df)
        frame.dtype_code(variables, "fit bandwidth")
        tbl, cols, _ = frame.device_table(variables)
        d = len(cols)
        h = np.empty(d)
        check(lib().pbn_diag_bandwidth(tbl.ctx.handle, tbl.handle, int_array(cols), d, tbl.rows(), self._rule, _dp(h)))
        return h

    def __str__(self):
        return self._name

    __repr__ = __str__

    def __eq__(self, other):
        return type(other) is type(self)

    def __hash__(self):
        return hash(type(self))

    def __getstate__(self):
        return ()

    def __setstate__(self, state):
        pass


class NormalReferenceRule(_NativeSelector):
    """kde/NormalReferenceRule.hpp:10-134."""
    _rule = _lib.BW_NORMAL_REFERENCE
    _name = "NormalReferenceRule"


class ScottsBandwidth(_NativeSelector):
    """kde/ScottsBandwidth.hpp:8-117."""
    _rule = _lib.BW_SCOTT
    _name = "ScottsBandwidth"


class UCVScorer:
    """pybnesian.UCVScorer (kde/UCV.hpp:12-45, pybindings_kde.cpp:187-190): N * UCV(H) for a
    diagonal (`score_diagonal(h)`) or full (`score_unconstrained(H)`) bandwidth."""

    def __init__(self, df, variables):
        frame = DataFrame.wrap(df)
        self._variables = list(variables)
        frame.dtype_code(self._variables, "score UCV")
        self._tbl, self._cols, _ = frame.device_table(self._variables)
        self._handle = ctypes.c_void_p()
        check(lib().pbn_ucv_create(self._tbl.ctx.handle, self._tbl.handle, int_array(self._cols), len(self._cols),
                                   self._tbl.rows(), ctypes.byref(self._handle)))

    def score_diagonal(self, diagonal_bandwidth):
        h = np.ascontiguousarray(np.asarray(diagonal_bandwidth, dtype=np.float64).ravel())
        d = len(self._variables)
        if h.size != d:
            raise ValueError("Wrong dimension for bandwidth vector. it should be a %d vector." % d)
        return self._score(h, 1)

    def score_unconstrained(self, bandwidth):
        d = len(self._variables)
        H = np.asarray(bandwidth, dtype=np.float64)
        if H.shape != (d, d):
            raise ValueError("Wrong dimension for bandwidth matrix. it should be a %dx%d matrix." % (d, d))
        return self._score(np.asfortranarray(H), 0)

    def _score(self, H, is_diag):
        from . import parallel
        out = ctypes.c_double()
        if parallel.active():
            # each rank sums its slice of the pair-tile schedule; 2 doubles are all-reduced
            s2, s1 = ctypes.c_double(), ctypes.c_double()
            with parallel.guard() as g:
                check(lib().pbn_ucv_pair_sums(self._handle, _dp(H), is_diag, parallel.rank(), parallel.world_size(),
                                              ctypes.byref(s2), ctypes.byref(s1)))
            tot = parallel.all_reduce_sum(np.array([s2.value, s1.value]), self._tbl.ctx, error=g.error)
            check(lib().pbn_ucv_score_from_sums(self._handle, _dp(H), is_diag, float(tot[0]), float(tot[1]), ctypes.byref(out)))
        else:
            check(lib().pbn_ucv_score(self._handle, _dp(H), is_diag, ctypes.byref(out)))
        return out.value

    def pair_sums(self, bandwidth, part=0, nparts=1, diagonal=False):
        """(sum e^{-s/4}, sum e^{-s/2}) over one slice of the pair-tile schedule (multi-GPU split)."""
        H = np.asfortranarray(np.asarray(bandwidth, dtype=np.float64))
        s2, s1 = ctypes.c_double(), ctypes.c_double()
        check(lib().pbn_ucv_pair_sums(self._handle, _dp(H), 1 if diagonal else 0, int(part), int(nparts),
                                      ctypes.byref(s2), ctypes.byref(s1)))
        return s2.value, s1.value

    def num_pairs(self):
        return int(lib().pbn_ucv_pairs(self._handle))

    def __del__(self):
        try:
            if self._handle:
                lib().pbn_ucv_free(self._handle)
                self._handle = None
        except Exception:
            pass


class UCV(BandwidthSelector):
    """pybnesian.UCV (kde/UCV.hpp:47-56, UCV.cpp:452-525): unbiased cross-validation bandwidth,
    Nelder-Mead from the normal-reference start."""

    last_evaluations = 0

    def _run(self, df, variables, diagonal):
        variables = list(variables)
        frame = DataFrame.wrap(df)
        frame.dtype_code(variables, "fit bandwidth")
        tbl, cols, _ = frame.device_table(variables)
        d = len(cols)
        out = np.empty(d) if diagonal else np.empty((d, d), order="F")
        n = ctypes.c_int()
        check(lib().pbn_ucv_bandwidth(tbl.ctx.handle, tbl.handle, int_array(cols), d, tbl.rows(), 1 if diagonal else 0,
                                      _dp(out), ctypes.byref(n)))
        self.last_evaluations = n.value
        return out

    def bandwidth(self, df, variables):
        if not list(variables):
            return np.empty((0, 0))
        return self._run(df, variables, False)

    def diag_bandwidth(self, df, variables):
        if not list(variables):
            return np.empty(0)
        return self._run(df, variables, True)

    def __str__(self):
        return "UCV"

    __repr__ = __str__

    def __getstate__(self):
        return ()

    def __setstate__(self, state):
        pass


class _FittedHandle:
    """Owns a pbn_kde (whitened training rows resident on the GPU)."""

    def __init__(self, handle):
        self.handle = handle

    def __del__(self):
        try:
            if self.handle:
                lib().pbn_kde_free(self.handle)
                self.handle = None
        except Exception:
            pass


def _fit_handle(tbl, cols, rows, H, ckde=False, product=False):
    H = np.asfortranarray(np.asarray(H, dtype=np.float64))
    h = ctypes.c_void_p()
    fn = lib().pbn_product_kde_fit if product else (lib().pbn_ckde_fit if ckde else lib().pbn_kde_fit)
    check(fn(tbl.ctx.handle, tbl.handle, int_array(cols), len(cols), rows, _dp(H), ctypes.byref(h)))
    return _FittedHandle(h)


def _run_logl(fitted, frame, variables, want_logl, want_slogl):
    """Shared by KDE and CKDE: returns (logl over all rows with NaN at null rows, slogl).

    With several ranks (pybnesian_b200.parallel) every rank holds the same frame and evaluates a contiguous
    shard of its test rows against the replicated training set; slogl is a 1-double all-reduce, logl a
    zero-padded vector all-reduce (each element written by exactly one rank)."""
    from . import parallel
    tbl, cols, mask = frame.device_table(variables)
    m = tbl.nrows
    b, e = parallel.shard_range(m) if parallel.active() else (0, m)
    out = np.zeros(m) if want_logl else None
    s = ctypes.c_double(0.0)
    with parallel.guard() as g:  # a failing rank still enters the collective; every rank raises after it
        if e > b or not parallel.active():
            shard = out[b:e] if want_logl else None
            check(lib().pbn_kde_logl(tbl.ctx.handle, fitted.handle, tbl.handle, int_array(cols), tbl.rows(b, e),
                                     _dp(shard) if want_logl else None, ctypes.byref(s) if want_slogl else None))
    total = s.value
    if parallel.active():
        if want_slogl:
            total = float(parallel.all_reduce_sum(np.array([s.value]), tbl.ctx, error=g.error)[0])
        if want_logl:
            out = parallel.all_reduce_sum(out, tbl.ctx, error=g.error)
    if want_logl and mask is not None:
        full = np.full(frame.num_rows, np.nan)
        full[mask] = out
        out = full
    return out, total


class KDE:
    """pybnesian.KDE (kde/KDE.hpp:292-417, pybindings_kde.cpp:212-302)."""

    def __init__(self, variables, bandwidth_selector=None):
        variables = list(variables)
        if bandwidth_selector is None:
            bandwidth_selector = NormalReferenceRule()
        if not isinstance(bandwidth_selector, BandwidthSelector):
            raise RuntimeError("Bandwidth selector procedure must be non-null.")
        if not variables:
            raise ValueError("Cannot create a KDE model with 0 variables")
        self._variables = variables
        self._bselector = bandwidth_selector
        self._fitted = False
        self._bandwidth = np.empty((0, 0))
        self._handle = None
        self._train = None  # (DeviceTable, cols, rows)
        self._N = 0
        self._dtype = _lib.PBN_F64

    def variables(self):
        return list(self._variables)

    def num_variables(self):
        return len(self._variables)

    def fitted(self):
        return self._fitted

    def _check_fitted(self):
        if not self._fitted:
            raise ValueError("KDE factor not fitted.")

    def num_instances(self):
        self._check_fitted()
        return self._N

    def data_type(self):
        self._check_fitted()
        return _ARROW_TYPE[self._dtype]

    @property
    def bandwidth(self):
        return self._bandwidth

    @bandwidth.setter
    def bandwidth(self, new_bandwidth):
        H = np.asarray(new_bandwidth, dtype=np.float64)
        d = len(self._variables)
        if H.ndim != 2 or H.shape[0] != H.shape[1] or H.shape[0] != d:
            raise ValueError("The bandwidth matrix must be a square matrix with shape (%d, %d)" % (d, d))
        self._bandwidth = np.array(H)
        if self._fitted and d > 0:
            tbl, cols, rows = self._train
            self._handle = _fit_handle(tbl, cols, rows, self._bandwidth)

    def fit(self, df):
        frame = DataFrame.wrap(df)
        self._dtype = frame.dtype_code(self._variables, "fit KDE")
        H = np.asarray(self._bselector.bandwidth(frame, self._variables), dtype=np.float64)
        d = len(self._variables)
        if H.shape != (d, d):
            raise ValueError("BandwidthSelector::bandwidth matrix must return an square matrix with shape (%d, %d)" % (d, d))
        tbl, cols, _ = frame.device_table(self._variables)
        self._fit_table(tbl, cols, tbl.rows(), H)

    def _fit_table(self, tbl, cols, rows, H):
        self._handle = _fit_handle(tbl, cols, rows, H)
        self._train = (tbl, list(cols), rows)
        self._bandwidth = np.array(H)
        self._N = int(lib().pbn_kde_num_instances(self._handle.handle))
        self._dtype = tbl.dtype_code
        self._fitted = True

    def _check_test(self, frame):
        self._check_fitted()
        t = frame.same_type(self._variables)
        if t != _ARROW_TYPE[self._dtype]:
            raise ValueError("Data type of training and test datasets is different.")

    def logl(self, df):
        frame = DataFrame.wrap(df)
        self._check_test(frame)
        return _run_logl(self._handle, frame, self._variables, True, False)[0]

    def slogl(self, df):
        frame = DataFrame.wrap(df)
        self._check_test(frame)
        return _run_logl(self._handle, frame, self._variables, False, True)[1]

    def dataset(self):
        """Training data read back from the device (KDE::training_data, kde/KDE.hpp:419-449)."""
        self._check_fitted()
        tbl, cols, rows = self._train
        arrays = [pa.array(tbl.download(c, rows)) for c in cols]
        # a pyarrow.RecordBatch, what the reference's DataFrame type caster returns (dataset/dataset.hpp:2120-2143)
        return pa.RecordBatch.from_arrays(arrays, names=self._variables)

    def lognorm_const(self):
        self._check_fitted()
        return lib().pbn_kde_lognorm(self._handle.handle)

    def save(self, filename):
        if not filename.endswith(".pickle"):
            filename += ".pickle"
        with open(filename, "wb") as f:
            pickle.dump(self, f)

    # pickle layout follows KDE::__getstate__ (kde/KDE.hpp:642-666)
    def __getstate__(self):
        bw, training, lognorm, n_export, type_id = np.empty((0, 0)), np.empty(0), -1.0, -1, -1
        if self._fitted:
            tbl, cols, rows = self._train
            training = np.concatenate([tbl.download(c, rows) for c in cols])
            lognorm = self.lognorm_const()
            n_export = self._N
            type_id = self._dtype
            bw = self._bandwidth
        return (self._variables, self._fitted, self._bselector, bw, training, lognorm, n_export, type_id)

    def __setstate__(self, t):
        self.__init__(t[0], t[2])
        if t[1]:
            n, code = int(t[6]), int(t[7])
            data = np.asarray(t[4]).reshape(len(t[0]), n)
            from .dataset import DeviceTable
            tbl = DeviceTable(_lib.default_context(), [data[i] for i in range(len(t[0]))], code)
            self._fit_table(tbl, list(range(len(t[0]))), tbl.rows(), np.asarray(t[3]))

    def __str__(self):
        return "KDE(" + ", ".join(self._variables) + ")"

    __repr__ = __str__


class ProductKDE:
    """pybnesian.ProductKDE (kde/ProductKDE.{hpp,cpp}, pybindings_kde.cpp:304-393): product of univariate Gaussian
    kernels, i.e. a KDE with a diagonal bandwidth `h` (variances) from `BandwidthSelector.diag_bandwidth`.  Evaluated
    by the same fused pair kernel as KDE (whitening with diag(1/sqrt(h)), include/pbn_cuda.h: pbn_product_kde_fit)."""

    def __init__(self, variables, bandwidth_selector=None):
        variables = list(variables)
        if bandwidth_selector is None:
            bandwidth_selector = NormalReferenceRule()
        if not isinstance(bandwidth_selector, BandwidthSelector):
            raise RuntimeError("Bandwidth selector procedure must be non-null.")
        if not variables:
            raise ValueError("Cannot create a ProductKDE model with 0 variables")
        self._variables = variables
        self._bselector = bandwidth_selector
        self._fitted = False
        self._bandwidth = np.empty(0)
        self._handle = None
        self._train = None
        self._N = 0
        self._dtype = _lib.PBN_F64

    def variables(self):
        return list(self._variables)

    def num_variables(self):
        return len(self._variables)

    def fitted(self):
        return self._fitted

    def _check_fitted(self):
        if not self._fitted:
            raise ValueError("ProductKDE factor not fitted.")

    def num_instances(self):
        self._check_fitted()
        return self._N

    def data_type(self):
        self._check_fitted()
        return _ARROW_TYPE[self._dtype]

    def bandwidth_type(self):
        return self._bselector

    @property
    def bandwidth(self):
        return self._bandwidth

    @bandwidth.setter
    def bandwidth(self, new_bandwidth):
        h = np.asarray(new_bandwidth, dtype=np.float64)
        d = len(self._variables)
        if h.ndim != 1 or h.shape[0] != d:
            raise ValueError("The bandwidth matrix must be a vector with shape (%d)" % d)
        self._bandwidth = np.array(h)
        if self._fitted and d > 0:
            tbl, cols, rows = self._train
            self._handle = _fit_handle(tbl, cols, rows, self._bandwidth, product=True)

    def fit(self, df):
        frame = DataFrame.wrap(df)
        self._dtype = frame.dtype_code(self._variables, "fit ProductKDE")
        h = np.asarray(self._bselector.diag_bandwidth(frame, self._variables), dtype=np.float64).ravel()
        d = len(self._variables)
        if h.shape != (d,):
            raise ValueError("BandwidthSelector::diag_bandwidth must return a vector with shape (%d)" % d)
        tbl, cols, _ = frame.device_table(self._variables)
        self._fit_table(tbl, cols, tbl.rows(), h)

    def _fit_table(self, tbl, cols, rows, h):
        self._handle = _fit_handle(tbl, cols, rows, h, product=True)
        self._train = (tbl, list(cols), rows)
        self._bandwidth = np.array(h)
        self._N = int(lib().pbn_kde_num_instances(self._handle.handle))
        self._dtype = tbl.dtype_code
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

    def dataset(self):
        """ProductKDE::training_data (kde/ProductKDE.hpp:122-151): read back from the device."""
        self._check_fitted()
        tbl, cols, rows = self._train
        arrays = [pa.array(tbl.download(c, rows)) for c in cols]
        # a pyarrow.RecordBatch, what the reference's DataFrame type caster returns (dataset/dataset.hpp:2120-2143)
        return pa.RecordBatch.from_arrays(arrays, names=self._variables)

    def lognorm_const(self):
        self._check_fitted()
        return lib().pbn_kde_lognorm(self._handle.handle)

    def save(self, filename):
        if not filename.endswith(".pickle"):
            filename += ".pickle"
        with open(filename, "wb") as f:
            pickle.dump(self, f)

    # pickle layout of ProductKDE::__getstate__ (kde/ProductKDE.hpp:311-337): one training vector per variable
    def __getstate__(self):
        bw, training, lognorm, n_export, type_id = np.empty(0), [], -1.0, -1, -1
        if self._fitted:
            tbl, cols, rows = self._train
            training = [tbl.download(c, rows) for c in cols]
            lognorm, n_export, type_id, bw = self.lognorm_const(), self._N, self._dtype, self._bandwidth
        return (self._variables, self._fitted, self._bselector, bw, training, lognorm, n_export, type_id)

    def __setstate__(self, t):
        if len(t) != 8:
            raise RuntimeError("Not valid ProductKDE.")
        self.__init__(t[0], t[2])
        if t[1]:
            from .dataset import DeviceTable
            tbl = DeviceTable(_lib.default_context(), [np.asarray(c) for c in t[4]], int(t[7]))
            self._fit_table(tbl, list(range(len(t[0]))), tbl.rows(), np.asarray(t[3], dtype=np.float64))

    def __str__(self):
        return "ProductKDE(" + ", ".join(self._variables) + ")"

    __repr__ = __str__
