"""Hybrid factors: continuous CPDs with discrete parents.

Mirrors factors/discrete/DiscreteAdaptator.hpp:88-568 (``HCKDE`` = DiscreteAdaptator<CKDE, CKDEFitter>,
CKDE.hpp:747-770; ``CLinearGaussianCPD`` = DiscreteAdaptator<LinearGaussianCPD, LinearGaussianFitter>,
LinearGaussianCPD.hpp:122-140), factors/discrete/discrete_indices.{hpp,cpp}, factors/assignment.hpp and the
sliver of factors/discrete/DiscreteFactor.{hpp,cpp} + learning/parameters/mle_DiscreteFactor.cpp a hybrid
network needs for its discrete nodes.

The reference takes one Arrow ``Take`` of the frame per discrete configuration per call and runs the base
factor on each.  Here a frame is gathered ONCE on the device into configuration-major order
(``pbn_discrete_slices`` + ``pbn_table_take``): each configuration is a contiguous row range of the grouped
table; base factors are fitted on row ranges, and all per-configuration CKDEs are evaluated by one
multi-job pair-kernel launch (``pbn_kde_logl_multi``).
"""
import ctypes
import pickle

import numpy as np
import pandas as pd
import pyarrow as pa

from . import _lib
from ._lib import Rows, SingularCovarianceData, check, int_array, lib
from .dataset import DataFrame, DeviceTable, _DTYPE_CODE
from .factors import CKDE, CKDEType, Factor, FactorType, LinearGaussianCPD, LinearGaussianCPDType
from .kde import BandwidthSelector, NormalReferenceRule, _NativeSelector

MACHINE_TOL = 1.4901161193847656e-08  # util/math_constants.hpp:30


# ----------------------------------------------------------------------------------------------
# Assignment (factors/assignment.hpp:150-270, pybindings_factors.cpp:663-724)
# ----------------------------------------------------------------------------------------------
def _assignment_value(v):
    if isinstance(v, str):
        return v
    if isinstance(v, (int, float, np.integer, np.floating)) and not isinstance(v, bool):
        return float(v)
    raise TypeError("an assignment value must be a str or a number")


class Assignment:
    def __init__(self, assignments=None):
        self._map = {str(k): _assignment_value(v) for k, v in dict(assignments or {}).items()}

    def value(self, variable):
        if variable not in self._map:
            raise ValueError("Variable " + str(variable) + " not found in the assignment.")
        return self._map[variable]

    def has_variables(self, variables):
        return all(v in self._map for v in variables)

    def empty(self):
        return not self._map

    def size(self):
        return len(self._map)

    def insert(self, variable, value):
        self._map.setdefault(str(variable), _assignment_value(value))

    def remove(self, variable):
        self._map.pop(variable, None)

    def index(self, variables, variable_values, strides):
        """Assignment::index (assignment.hpp:198-216)."""
        idx = 0
        for i, v in enumerate(variables):
            val = self.value(v)
            if not isinstance(val, str):
                raise RuntimeError("Assignment value is not string.")
            if val not in variable_values[i]:
                raise ValueError("Category \"" + val + "\" is not valid for variable " + v)
            idx += variable_values[i].index(val) * int(strides[i])
        return idx

    @staticmethod
    def from_index(index, variables, variable_values, cardinality, strides):
        """Assignment::from_index (assignment.hpp:218-231)."""
        return Assignment({v: variable_values[i][(index // int(strides[i])) % int(cardinality[i])]
                           for i, v in enumerate(variables)})

    def __iter__(self):
        return iter(self._map.items())

    def __eq__(self, other):
        return isinstance(other, Assignment) and self._map == other._map

    def __ne__(self, other):
        return not self == other

    def __hash__(self):
        return hash(frozenset(self._map.items()))

    def __str__(self):
        def show(v):
            return v if isinstance(v, str) else "%f" % v
        return "[" + ", ".join("%s = %s" % (k, show(v)) for k, v in self._map.items()) + "]"

    __repr__ = __str__

    def __getstate__(self):
        return dict(self._map)

    def __setstate__(self, state):
        self._map = dict(state)


# ----------------------------------------------------------------------------------------------
# categorical columns (factors/discrete/discrete_indices.{hpp,cpp})
# ----------------------------------------------------------------------------------------------
def _dictionary_column(frame, name):
    col = frame._col(name)
    if not pa.types.is_dictionary(col.type):
        raise ValueError("Variable " + name + " is not categorical.")
    # check_is_string_dictionary (discrete_indices.cpp:5-11)
    if not (pa.types.is_string(col.type.value_type) or pa.types.is_large_string(col.type.value_type)):
        raise ValueError("The categories of the data must be of type string. The categories of the variable " + name +
                         " are of type " + str(col.type.value_type) + ".")
    return col


def _categories(frame, name):
    return [str(s) for s in _dictionary_column(frame, name).dictionary.to_pylist()]


def _codes(frame, name):
    """Dictionary indices widened to int32 (null slots hold 0 and are masked by the caller)."""
    col = _dictionary_column(frame, name)
    idx = col.indices
    if idx.null_count:
        idx = idx.fill_null(0)
    return np.ascontiguousarray(idx.to_numpy(zero_copy_only=False), dtype=np.int32)


def check_domain_variable(frame, name, values):
    """discrete_indices.cpp:203-224."""
    if name not in frame._index:
        raise IndexError("Column index " + str(name) + " do not exist in DataFrame.")
    cats = _categories(frame, name)
    if len(cats) != len(values):
        raise ValueError("Variable " + name + " does not contain the same categories.")
    for j, (a, b) in enumerate(zip(values, cats)):
        if a != b:
            raise ValueError("Category at index " + str(j) + " is different for variable " + name)


def create_cardinality_strides(frame, variables):
    """discrete_indices.cpp:118-140."""
    card = np.array([len(_categories(frame, v)) for v in variables], dtype=np.int32)
    strides = np.ones(len(variables), dtype=np.int32)
    for i in range(1, len(variables)):
        strides[i] = strides[i - 1] * card[i - 1]
    return card, strides


def discrete_slices(frame, discrete_vars, strides, num_factors, extra_valid=None):
    """discrete_slice_indices (discrete_indices.cpp:166-201) through the C ABI: (order, offsets) where
    order[offsets[c]:offsets[c+1]] are the row ids of configuration c, ascending.  ``extra_valid`` further
    restricts the participating rows (the rows the base factor would drop anyway: nulls in its own columns)."""
    n = frame.num_rows
    codes = [_codes(frame, v) for v in discrete_vars]
    valid = frame.combined_valid(discrete_vars)
    if extra_valid is not None:
        valid = extra_valid if valid is None else (valid & extra_valid)
    i32p = ctypes.POINTER(ctypes.c_int32)
    ptrs = (i32p * max(len(codes), 1))(*[c.ctypes.data_as(i32p) for c in codes])
    order = np.empty(n, dtype=np.int32)
    offsets = np.zeros(num_factors + 1, dtype=np.int64)
    vmask = None if valid is None else np.ascontiguousarray(valid, dtype=np.uint8)
    strides = np.ascontiguousarray(strides, dtype=np.int32)
    check(lib().pbn_discrete_slices(ptrs, strides.ctypes.data_as(i32p), len(codes), n,
                                    vmask.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)) if vmask is not None else None,
                                    int(num_factors), order.ctypes.data_as(i32p),
                                    offsets.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))))
    return order[:int(offsets[-1])], offsets


def _take_table(tbl, order):
    """pbn_table_take: a new resident table holding rows `order` of `tbl`."""
    out = DeviceTable.__new__(DeviceTable)
    out.ctx, out.ncols, out.nrows, out.dtype_code = tbl.ctx, tbl.ncols, int(order.size), tbl.dtype_code
    out.handle = ctypes.c_void_p()
    order = np.ascontiguousarray(order, dtype=np.int32)
    check(lib().pbn_table_take(tbl.ctx.handle, tbl.handle, order.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                               order.size, ctypes.byref(out.handle)))
    return out


# ----------------------------------------------------------------------------------------------
# DiscreteFactor (factors/discrete/DiscreteFactor.{hpp,cpp}; mle_DiscreteFactor.cpp) — host integer work
# ----------------------------------------------------------------------------------------------
class DiscreteFactorType(FactorType):
    def new_factor(self, model, variable, evidence, *args, **kwargs):
        return DiscreteFactor(variable, evidence, *args, **kwargs)

    def __str__(self):
        return "DiscreteFactor"

    __repr__ = __str__


class DiscreteFactor(Factor):
    """Categorical CPD P(variable | discrete evidence) as a table of log-probabilities."""

    def __init__(self, variable, evidence):
        super().__init__(variable, evidence)
        self._variables = [variable] + list(evidence)
        self._fitted = False
        self._logprob = None
        self._cardinality = None
        self._strides = None
        self._values = None  # categories per variable (variable first)
        self._arrow_type = None

    def type(self):
        return DiscreteFactorType()

    def fitted(self):
        return self._fitted

    def _check_fitted(self):
        if not self._fitted:
            raise ValueError("DiscreteFactor factor not fitted.")

    def data_type(self):
        self._check_fitted()
        return self._arrow_type

    def fit(self, df):
        frame = DataFrame.wrap(df)
        t = frame.same_type(self._variables)
        if not pa.types.is_dictionary(t):
            raise ValueError("Wrong data type to fit DiscreteFactor. Categorical data is expected.")
        card, strides = create_cardinality_strides(frame, self._variables)
        idx, _ = self._indices(frame, strides)
        counts = np.bincount(idx, minlength=int(card.prod())).astype(np.int64)
        # mle_DiscreteFactor.cpp:14-36: normalise per parent configuration; an unseen one is uniform
        c0 = int(card[0])
        counts2 = counts.reshape(-1, c0)
        tot = counts2.sum(axis=1, keepdims=True)
        with np.errstate(divide="ignore"):
            logprob = np.where(tot == 0, np.log(1.0 / c0), np.log(counts2.astype(np.float64)) - np.log(np.maximum(tot, 1).astype(np.float64)))
        self._logprob = logprob.reshape(-1)
        self._cardinality, self._strides = card, strides
        self._values = [_categories(frame, v) for v in self._variables]
        self._arrow_type = t
        self._fitted = True

    def _indices(self, frame, strides):
        valid = frame.combined_valid(self._variables)
        idx = np.zeros(frame.num_rows, dtype=np.int64)
        for v, s in zip(self._variables, strides):
            idx += _codes(frame, v).astype(np.int64) * int(s)
        return (idx if valid is None else idx[valid]), valid

    def _check_domain(self, frame):
        self._check_fitted()
        for v, vals in zip(self._variables, self._values):
            check_domain_variable(frame, v, vals)

    def logl(self, df):
        frame = DataFrame.wrap(df)
        self._check_domain(frame)
        idx, valid = self._indices(frame, self._strides)
        if valid is None:
            return self._logprob[idx]
        out = np.full(frame.num_rows, np.nan)
        out[valid] = self._logprob[idx]
        return out

    def slogl(self, df):
        frame = DataFrame.wrap(df)
        self._check_domain(frame)
        idx, _ = self._indices(frame, self._strides)
        if idx.size == 0:
            return 0.0
        # sequential double accumulation in row order (DiscreteFactor.cpp:150-159); cumsum adds left to right
        return float(np.cumsum(self._logprob[idx])[-1])

    def sample(self, n, evidence_values=None, seed=None):
        """DiscreteFactor::sample / sample_indices (DiscreteFactor.cpp:173-208, DiscreteFactor.hpp:144-207): one
        std::uniform_real_distribution<double> draw per instance against the running sum of the row of the
        probability table selected by the discrete evidence; the last category takes the remainder."""
        from .factors import _random_seed
        if n < 0:
            raise ValueError("n should be a non-negative number")
        self._check_fitted()
        seed = _random_seed(seed)
        c0 = int(self._cardinality[0])
        # running sums over the first c0 - 1 categories of every parent configuration, left to right
        accum = np.cumsum(np.exp(self._logprob.reshape(-1, c0))[:, :max(c0 - 1, 0)], axis=1)
        parent = np.zeros(n, dtype=np.int64)
        if self._evidence:
            frame = DataFrame.wrap(evidence_values) if evidence_values is not None else None
            if frame is None or not frame.has_columns(self._evidence):
                raise ValueError("Evidence values not present for sampling.")
            for e, vals in zip(self._evidence, self._values[1:]):
                check_domain_variable(frame, e, vals)
            if frame.num_rows != n:
                raise ValueError("Evidence values do not have " + str(n) + " rows to sample.")
            if frame.null_count(self._evidence) > 0:
                raise ValueError("Evidence values contain null rows in the evidence variables.")
            for e, st in zip(self._evidence, self._strides[1:]):
                parent += _codes(frame, e).astype(np.int64) * (int(st) // c0)
        u = np.empty(n)
        check(lib().pbn_uniform_real(n, seed, _lib.PBN_F64, u.ctypes.data_as(ctypes.c_void_p)))
        # first j with u < accum[parent, j], else the last category
        idx = (u[:, None] >= accum[parent]).sum(axis=1) if c0 > 1 else np.zeros(n, dtype=np.int64)
        index_type = self._arrow_type.index_type
        indices = pa.array(idx.astype(index_type.to_pandas_dtype()), type=index_type)
        return pa.DictionaryArray.from_arrays(indices, pa.array(self._values[0], type=self._arrow_type.value_type))

    def __getstate__(self):
        return (self._variable, self._evidence, self._fitted, self._logprob, self._cardinality, self._strides,
                self._values, self._arrow_type)

    def __setstate__(self, t):
        self.__init__(t[0], t[1])
        if t[2]:
            self._logprob, self._cardinality, self._strides, self._values, self._arrow_type = t[3:8]
            self._fitted = True

    def __str__(self):
        ev = (" | " + ", ".join(self._evidence)) if self._evidence else ""
        return "[DiscreteFactor] P(" + self._variable + ev + ")" + ("" if self._fitted else " not fitted")

    __repr__ = __str__


class DiscreteFactorParams:
    """pybnesian.DiscreteFactorParams (pybindings_parameters.cpp:93-130): the log-probability table, one axis per
    variable (variable first)."""

    def __init__(self, logprob):
        self.logprob = np.asarray(logprob, dtype=np.float64)


class MLEDiscreteFactor:
    """MLE<DiscreteFactor> (learning/parameters/mle_DiscreteFactor.cpp:14-36)."""

    def estimate(self, df, variable, evidence):
        frame = DataFrame.wrap(df)
        f = DiscreteFactor(frame.column_name(variable), [frame.column_name(e) for e in evidence])
        f.fit(frame)
        shape = [int(c) for c in f._cardinality]
        return DiscreteFactorParams(f._logprob.reshape(shape, order="F"))


# ----------------------------------------------------------------------------------------------
# DiscreteAdaptator (factors/discrete/DiscreteAdaptator.hpp:88-568)
# ----------------------------------------------------------------------------------------------
class _GroupedFrame:
    """A frame gathered on the device into configuration-major order: `tbl` holds the continuous columns
    `variables` of the rows `order`; configuration c is the row range [offsets[c], offsets[c+1])."""

    @staticmethod
    def of(frame, variables, discrete_vars, strides, num_factors):
        """Grouped tables are cached on the frame wrapper (frames are immutable views)."""
        key = ("grouped", tuple(variables), tuple(discrete_vars), tuple(int(s) for s in strides), int(num_factors),
               id(_lib.default_context()))
        g = frame._tables.get(key)
        if g is None:
            g = _GroupedFrame(frame, variables, discrete_vars, strides, num_factors)
            frame._tables[key] = g
        return g

    def __init__(self, frame, variables, discrete_vars, strides, num_factors):
        self.frame = frame
        code = frame.dtype_code(variables)
        cont_valid = frame.combined_valid(variables)
        self.order, self.offsets = discrete_slices(frame, discrete_vars, strides, num_factors, cont_valid)
        # configurations the reference would see as non-empty (its slices ignore nulls in continuous columns)
        if cont_valid is None:
            self.present = np.diff(self.offsets) > 0
        else:
            _, off_all = discrete_slices(frame, discrete_vars, strides, num_factors)
            self.present = np.diff(off_all) > 0
        base, self.cols, mask = frame.device_table(variables)
        if mask is not None:
            # device_table compacted the rows with nulls away: translate frame row ids to compacted ids
            remap = np.cumsum(mask, dtype=np.int64) - 1
            take = remap[self.order].astype(np.int32)
        else:
            take = self.order
        self.tbl = _take_table(base, take)
        self.dtype_code = code

    def rows(self, c):
        return Rows.single(int(self.offsets[c]), int(self.offsets[c + 1]))

    def count(self, c):
        return int(self.offsets[c + 1] - self.offsets[c])


class DiscreteAdaptator(Factor):
    """One base factor per configuration of the discrete evidence (DiscreteAdaptator.hpp:88-161)."""

    _name = None
    _base_type = None

    def __init__(self, variable, evidence, *args):
        super().__init__(variable, evidence)
        # BaseFactorParametersImpl / SpecificBaseFactorParameters (DiscreteAdaptator.hpp:22-86)
        if len(args) == 1 and isinstance(args[0], dict) and all(isinstance(k, Assignment) for k in args[0]):
            self._specific = True
            self._args = {k: (v if isinstance(v, tuple) else (v,)) for k, v in args[0].items()}
        else:
            self._specific = False
            self._args = tuple(args)
        self._check_args()
        self._fitted = False
        self._discrete_evidence = []
        self._discrete_values = []
        self._continuous_evidence = []
        self._cardinality = np.empty(0, dtype=np.int32)
        self._strides = np.empty(0, dtype=np.int32)
        self._factors = []

    def _check_args(self):
        pass

    def _new_base(self, variable, evidence, args):
        raise NotImplementedError

    def _base_fit(self, factor, grouped, c, frame, rows_of_c):
        """BaseFitter::fit: returns False when the configuration must be left without a factor."""
        raise NotImplementedError

    def _initialize(self, assignment):
        if self._specific:
            args = self._args.get(assignment, ())
        else:
            args = self._args
        return self._new_base(self._variable, self._continuous_evidence, args)

    def type(self):
        return self._base_type()

    def fitted(self):
        return self._fitted

    def _check_fitted(self):
        if not self._fitted:
            raise ValueError("Factor " + str(self) + " not fitted.")

    def data_type(self):
        self._check_fitted()
        for f in self._factors:
            if f is not None:
                return f.data_type()
        raise ValueError("Factor " + str(self) + " has no fitted configuration.")

    # -- checks (DiscreteAdaptator.hpp:168-199) --------------------------------------------------
    @staticmethod
    def _raise_continuous(frame, name):
        if name not in frame._index:
            raise IndexError("Column index " + str(name) + " do not exist in DataFrame.")
        if frame._col(name).type not in _DTYPE_CODE:
            raise ValueError("Variable " + name + " must have \"double\" or \"float\" data type.")

    def _run_checks(self, frame, check_variable):
        self._check_fitted()
        if check_variable:
            self._raise_continuous(frame, self._variable)
        for e in self._evidence:
            if e not in frame._index:
                raise IndexError("Column index " + str(e) + " do not exist in DataFrame.")
        for e in self._continuous_evidence:
            self._raise_continuous(frame, e)
        for e, vals in zip(self._discrete_evidence, self._discrete_values):
            check_domain_variable(frame, e, vals)

    # -- fit (DiscreteAdaptator.hpp:201-257) -------------------------------------------------------
    def fit(self, df):
        frame = DataFrame.wrap(df)
        discrete, continuous = [], []
        for e in self._evidence:
            t = frame._col(e).type
            if pa.types.is_dictionary(t):
                discrete.append(e)
            elif t in _DTYPE_CODE:
                continuous.append(e)
            else:
                raise ValueError("Non valid data type for variable " + e + ". Only \"dictionary\", \"double\" and "
                                 "\"float\" data types are allowed.")
        self._discrete_evidence, self._continuous_evidence = discrete, continuous
        self._discrete_values = []
        self._factors = []
        if not discrete:
            f = self._initialize(Assignment())
            f.fit(frame)
            self._factors = [f]
            self._fitted = True
            return
        self._cardinality, self._strides = create_cardinality_strides(frame, discrete)
        self._discrete_values = [_categories(frame, e) for e in discrete]
        num_factors = int(self._cardinality.prod())
        variables = [self._variable] + continuous
        grouped = _GroupedFrame.of(frame, variables, discrete, self._strides, num_factors)
        for c in range(num_factors):
            if not grouped.present[c]:
                self._factors.append(None)
                continue
            assignment = Assignment.from_index(c, discrete, self._discrete_values, self._cardinality, self._strides)
            f = self._initialize(assignment)
            if not f.fitted():
                if not self._base_fit(f, grouped, c):
                    f = None
            self._factors.append(f)
        self._fitted = True

    def conditional_factor(self, assignment):
        self._check_fitted()
        return self._factors[assignment.index(self._discrete_evidence, self._discrete_values, self._strides)]

    # -- logl / slogl (DiscreteAdaptator.hpp:259-325) ------------------------------------------------
    def _grouped_test(self, frame):
        return _GroupedFrame.of(frame, [self._variable] + self._continuous_evidence, self._discrete_evidence,
                                self._strides, len(self._factors))

    def _eval_grouped(self, grouped, want_logl):
        """(logl in grouped order or None, per-configuration sums)."""
        raise NotImplementedError

    def logl(self, df):
        frame = DataFrame.wrap(df)
        self._run_checks(frame, True)
        if not self._discrete_evidence:
            return self._factors[0].logl(frame)
        grouped = self._grouped_test(frame)
        vals, _ = self._eval_grouped(grouped, True)
        out = np.full(frame.num_rows, np.nan)
        out[grouped.order] = vals
        return out

    def slogl(self, df):
        frame = DataFrame.wrap(df)
        self._run_checks(frame, True)
        if not self._discrete_evidence:
            return self._factors[0].slogl(frame)
        grouped = self._grouped_test(frame)
        _, sums = self._eval_grouped(grouped, False)
        res = 0.0
        for c, f in enumerate(self._factors):  # configuration order (DiscreteAdaptator.hpp:315-320)
            if f is not None and grouped.count(c) > 0:
                res += float(sums[c])
        return res

    # -- sample (DiscreteAdaptator.hpp:426-520) --------------------------------------------------------
    def sample(self, n, evidence_values=None, seed=None):
        """Configuration i of the discrete evidence is sampled by its base factor with seed + i on the rows of that
        configuration; rows of a configuration without a factor are NaN.  As in the reference every base factor is asked
        for n instances (its random streams are laid out for n draws) and the first len(rows) are kept."""
        from .factors import _random_seed
        from .dataset import _NP_DTYPE
        if n < 0:
            raise ValueError("n should be a non-negative number")
        frame = DataFrame.wrap(evidence_values) if evidence_values is not None else None
        if self._evidence:
            if frame is None:
                raise ValueError("Evidence values not present for sampling.")
            self._run_checks(frame, False)
            if frame.num_rows != n:
                raise ValueError("Evidence values do not have " + str(n) + " rows to sample.")
            if frame.null_count(self._evidence) > 0:
                raise ValueError("Evidence values contain null rows in the evidence variables.")
        else:
            self._check_fitted()
        seed = _random_seed(seed)
        if not self._discrete_evidence:
            return self._factors[0].sample(n, frame, seed)
        order, offsets = discrete_slices(frame, self._discrete_evidence, self._strides, len(self._factors))
        npdt = _NP_DTYPE[_DTYPE_CODE[self.data_type()]]
        res = np.full(n, np.nan, dtype=npdt)
        cont = {e: frame.column_numpy(e) for e in self._continuous_evidence}
        for i, f in enumerate(self._factors):
            rows = order[int(offsets[i]):int(offsets[i + 1])]
            if rows.size == 0 or f is None:
                continue
            ev = None
            if cont:
                # the configuration's evidence rows, padded to n rows with its first row (the draws beyond len(rows)
                # are discarded; the reference reads past the end of the filtered columns there)
                pad = np.concatenate([rows, np.full(n - rows.size, rows[0], dtype=rows.dtype)])
                ev = pd.DataFrame({e: np.ascontiguousarray(c[pad]) for e, c in cont.items()})
            smp = f.sample(n, ev, (seed + i) & 0xFFFFFFFF)
            res[rows] = smp.to_numpy(zero_copy_only=False)[:rows.size]
        return pa.array(res)

    # -- text / pickle -------------------------------------------------------------------------------
    def __str__(self):
        ev = (" | " + ", ".join(self._evidence)) if self._evidence else ""
        s = "[" + self._name + "] P(" + self._variable + ev + ")"
        if not self._fitted:
            return s + " not fitted."
        if not self._discrete_evidence:
            return s + " = " + str(self._factors[0])
        lines = [s]
        for c, f in enumerate(self._factors):
            a = Assignment.from_index(c, self._discrete_evidence, self._discrete_values, self._cardinality, self._strides)
            lines.append("  " + str(a) + ": " + (str(f) if f is not None else "not fitted"))
        return "\n".join(lines)

    __repr__ = __str__

    def __getstate__(self):
        return (self._variable, self._evidence, (self._specific, pickle.dumps(self._args)), self._fitted,
                self._discrete_evidence, self._discrete_values, self._continuous_evidence, self._cardinality,
                self._strides, self._factors)

    def __setstate__(self, t):
        specific, blob = t[2]
        args = pickle.loads(blob)
        if specific:
            type(self).__init__(self, t[0], t[1], args)
        else:
            type(self).__init__(self, t[0], t[1], *args)
        if t[3]:
            (self._discrete_evidence, self._discrete_values, self._continuous_evidence, self._cardinality,
             self._strides, self._factors) = t[4:10]
            self._fitted = True


class HCKDE(DiscreteAdaptator):
    """pybnesian.HCKDE (CKDE.hpp:747-770, pybindings_factors.cpp:788-858): a CKDE per discrete configuration."""

    _name = "HCKDE"
    _base_type = CKDEType

    def _check_args(self):
        sels = [a for v in self._args.values() for a in v] if self._specific else list(self._args)
        if not self._specific and len(sels) > 1:
            raise TypeError("HCKDE(variable, evidence[, bandwidth_selector])")
        for s in sels:
            if not isinstance(s, BandwidthSelector):
                raise RuntimeError("Bandwidth selector procedure must be non-null.")

    def _new_base(self, variable, evidence, args):
        return CKDE(variable, evidence, *args)

    def _base_fit(self, f, grouped, c):
        """CKDEFitter::fit (CKDE.hpp:752-768): a singular covariance leaves the configuration unfitted."""
        rows = grouped.rows(c)
        try:
            sel = f.bandwidth_type()
            if isinstance(sel, _NativeSelector):
                H = sel._bandwidth_rows(grouped.tbl, grouped.cols, rows)
            else:  # Python-derived selector: called back with the configuration's own frame, as the reference does
                sub = grouped.frame.take(grouped.order[int(grouped.offsets[c]):int(grouped.offsets[c + 1])])
                H = np.asarray(sel.bandwidth(sub, [self._variable] + self._continuous_evidence), dtype=np.float64)
            f._fit_table(grouped.tbl, grouped.cols, rows, H)
            return True
        except SingularCovarianceData:
            return False

    def _eval_grouped(self, grouped, want_logl):
        from . import parallel
        F = len(self._factors)
        dt = self._factors_dtype()
        if dt is not None and grouped.dtype_code != dt:
            raise ValueError("Data type of training and test datasets is different.")
        handles = (ctypes.c_void_p * F)(*[f._handle.handle if f is not None else None for f in self._factors])
        b, e = parallel.shard_range(grouped.tbl.nrows) if parallel.active() else (0, grouped.tbl.nrows)
        rows = (Rows * F)()
        counts = np.zeros(F, dtype=np.int64)
        for c in range(F):
            lo, hi = max(int(grouped.offsets[c]), b), min(int(grouped.offsets[c + 1]), e)
            hi = max(hi, lo)
            rows[c] = Rows.single(lo, hi)
            counts[c] = hi - lo
        total = int(counts.sum())
        dp = ctypes.POINTER(ctypes.c_double)
        vals = np.empty(total) if want_logl else None
        sums = np.zeros(F)
        with parallel.guard() as g:
            check(lib().pbn_kde_logl_multi(grouped.tbl.ctx.handle, handles, F, grouped.tbl.handle, int_array(grouped.cols), rows,
                                           vals.ctypes.data_as(dp) if want_logl else None,
                                           None if want_logl else sums.ctypes.data_as(dp)))
        if parallel.active():
            if g.error is not None:
                parallel.all_reduce_sum(np.zeros(1), grouped.tbl.ctx, error=g.error)
            if want_logl:
                full = np.zeros(grouped.tbl.nrows)
                pos = 0
                for c in range(F):
                    full[rows[c].b0:rows[c].e0] = vals[pos:pos + counts[c]]
                    pos += counts[c]
                # NaN rows (configurations without a factor) do not survive a sum: mark and restore them
                nan = np.isnan(full)
                full[nan] = 0.0
                full = parallel.all_reduce_sum(full, grouped.tbl.ctx)
                nanc = parallel.all_reduce_sum(nan.astype(np.float64), grouped.tbl.ctx)
                full[nanc > 0] = np.nan
                vals = full
            else:
                sums = parallel.all_reduce_sum(sums, grouped.tbl.ctx)
        return vals, sums

    def _factors_dtype(self):
        for f in self._factors:
            if f is not None:
                return f._dtype
        return None


class CLinearGaussianCPD(DiscreteAdaptator):
    """pybnesian.CLinearGaussianCPD (LinearGaussianCPD.hpp:122-140, pybindings_factors.cpp:726-786)."""

    _name = "CLinearGaussianCPD"
    _base_type = LinearGaussianCPDType

    def _check_args(self):
        if not self._specific and len(self._args) not in (0, 2):
            raise TypeError("CLinearGaussianCPD(variable, evidence[, beta, variance])")

    def _new_base(self, variable, evidence, args):
        return LinearGaussianCPD(variable, evidence, *args)

    def _base_fit(self, f, grouped, c):
        """LinearGaussianFitter::fit (LinearGaussianCPD.hpp:127-138)."""
        cols = grouped.cols
        beta = np.empty(len(cols))
        var = ctypes.c_double()
        dp = ctypes.POINTER(ctypes.c_double)
        check(lib().pbn_lg_fit(grouped.tbl.ctx.handle, grouped.tbl.handle, int_array(cols), len(cols), grouped.rows(c),
                               beta.ctypes.data_as(dp), ctypes.byref(var)))
        f._beta, f._variance, f._fitted = beta, var.value, True
        return not (var.value < MACHINE_TOL or np.isinf(var.value))

    def _eval_grouped(self, grouped, want_logl):
        F = len(self._factors)
        dp = ctypes.POINTER(ctypes.c_double)
        vals = np.full(grouped.tbl.nrows, np.nan) if want_logl else None
        sums = np.zeros(F)
        for c, f in enumerate(self._factors):
            n = grouped.count(c)
            if f is None or n == 0:
                continue
            lo = int(grouped.offsets[c])
            s = ctypes.c_double(0.0)
            seg = vals[lo:lo + n] if want_logl else None
            check(lib().pbn_lg_logl(grouped.tbl.ctx.handle, grouped.tbl.handle, int_array(grouped.cols), len(grouped.cols),
                                    grouped.rows(c), np.ascontiguousarray(f._beta).ctypes.data_as(dp), float(f._variance),
                                    seg.ctypes.data_as(dp) if want_logl else None, None if want_logl else ctypes.byref(s)))
            sums[c] = s.value
        return vals, sums
