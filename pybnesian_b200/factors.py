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
        raise NotImplementedError

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


class CKDEType(FactorType):
    """factors/continuous/CKDE.hpp:17-60, CKDE.cpp:15-41 (discrete parents -> HCKDE is out of scope)."""

    def new_factor(self, model, variable, evidence, *args, **kwargs):
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
