"""Data ingestion: pandas / pyarrow -> dense column buffers -> resident device tables.

Mirrors the sliver of dataset::DataFrame that is on the KDE hot path
(/root/reference/pybnesian/dataset/dataset.{hpp,cpp}): `same_type` (dataset.cpp:253-271),
`combined_bitmap` / `valid_rows` (dataset.cpp:208-251) and `to_eigen` (dataset.hpp:236-338:
contiguous columns with the rows that hold a null in any selected column removed).
pandas input is converted with ``pyarrow.RecordBatch.from_pandas(df, None, False)`` exactly
as the reference does (dataset.cpp:53-56), so a pandas NaN arrives as an Arrow null.
"""
import ctypes

import numpy as np
import pyarrow as pa

from . import _lib
from ._lib import Rows, check, lib


def _to_record_batch(obj):
    if isinstance(obj, pa.RecordBatch):
        return obj
    if isinstance(obj, pa.Table):
        batches = obj.combine_chunks().to_batches()
        if len(batches) == 1:
            return batches[0]
        return pa.RecordBatch.from_arrays([obj.column(i).combine_chunks() for i in range(obj.num_columns)],
                                          names=obj.column_names)
    try:
        import pandas as pd
        if isinstance(obj, pd.DataFrame):
            return pa.RecordBatch.from_pandas(obj, None, False)
    except ImportError:  # pragma: no cover
        pass
    # Arrow C Data Interface, the boundary the reference's type caster uses (dataset/dataset.hpp:2088-2143:
    # extract_pycapsule_array + arrow::ImportRecordBatch): a (schema, array) pair of PyCapsules, or any object that
    # exports them through the PyCapsule protocol (__arrow_c_array__ for one struct array, __arrow_c_stream__ for a
    # stream of batches - polars / nanoarrow / duckdb frames)
    if isinstance(obj, tuple) and len(obj) == 2 and all(type(c).__name__ == "PyCapsule" for c in obj):
        return pa.RecordBatch._import_from_c_capsule(*obj)
    if hasattr(obj, "__arrow_c_array__"):
        return pa.record_batch(obj)
    if hasattr(obj, "__arrow_c_stream__"):
        return _to_record_batch(pa.table(obj))
    raise TypeError("expected a pandas.DataFrame, pyarrow.RecordBatch / Table, or an Arrow C-Data capsule exporter")


_DTYPE_CODE = {pa.float64(): _lib.PBN_F64, pa.float32(): _lib.PBN_F32}
_NP_DTYPE = {_lib.PBN_F64: np.float64, _lib.PBN_F32: np.float32}


class DeviceTable:
    """A resident, null-free column store on one GPU (pbn_table)."""

    def __init__(self, ctx, columns, dtype_code):
        self.ctx = ctx
        self.ncols = len(columns)
        self.nrows = int(columns[0].shape[0]) if columns else 0
        self.dtype_code = dtype_code
        cols = [np.ascontiguousarray(c, dtype=_NP_DTYPE[dtype_code]) for c in columns]
        ptrs = (ctypes.c_void_p * len(cols))(*[c.ctypes.data for c in cols])
        self.handle = ctypes.c_void_p()
        check(lib().pbn_table_upload(ctx.handle, ptrs, len(cols), self.nrows, dtype_code, ctypes.byref(self.handle)))

    def rows(self, begin=0, end=None):
        return Rows.single(begin, self.nrows if end is None else end)

    def download(self, col, rows=None):
        rows = rows or self.rows()
        n = (rows.e0 - rows.b0) + (rows.e1 - rows.b1)
        out = np.empty(n, dtype=_NP_DTYPE[self.dtype_code])
        check(lib().pbn_table_download(self.ctx.handle, self.handle, int(col), rows, out.ctypes.data_as(ctypes.c_void_p)))
        return out

    def __del__(self):
        try:
            if self.handle:
                lib().pbn_table_free(self.handle)
                self.handle = None
        except Exception:
            pass


class DataFrame:
    """Thin view of one Arrow RecordBatch (what dataset::DataFrame wraps).

    Wrap your data once (``DataFrame(df)``) and pass the wrapper to ``fit`` / ``logl`` /
    ``slogl`` to keep its columns resident on the GPU between calls; passing a raw pandas
    frame uploads per call, which is what the reference does.
    """

    def __init__(self, data):
        if isinstance(data, DataFrame):
            self.__dict__ = data.__dict__
            return
        self.rb = _to_record_batch(data)
        self.names = list(self.rb.schema.names)
        self._index = {n: i for i, n in enumerate(self.names)}
        self._tables = {}

    @staticmethod
    def wrap(data):
        return data if isinstance(data, DataFrame) else DataFrame(data)

    # -- schema ---------------------------------------------------------------------
    @property
    def num_rows(self):
        return self.rb.num_rows

    @property
    def columns(self):
        return list(self.names)

    def has_columns(self, cols):
        if isinstance(cols, str):
            cols = [cols]
        return all(c in self._index for c in cols)

    def _col(self, name):
        if name not in self._index:
            raise IndexError("Column index " + str(name) + " do not exist in DataFrame.")
        return self.rb.column(self._index[name])

    def same_type(self, variables):
        """dataset.cpp:253-271; returns the pbn dtype code."""
        variables = list(variables)
        if not variables:
            raise ValueError("Cannot check the data type of no columns")
        t0 = self._col(variables[0]).type
        for i, v in enumerate(variables[1:], start=1):
            t = self._col(v).type
            if t != t0:
                raise ValueError("Column 0 [%s] and column %d [%s] have different data types" % (t0, i, t))
        return t0

    def dtype_code(self, variables, what="fit KDE"):
        t = self.same_type(variables)
        if t not in _DTYPE_CODE:
            raise ValueError("Wrong data type to %s. [double] or [float] data is expected." % what)
        return _DTYPE_CODE[t]

    def arrow_type(self, variables):
        return self.same_type(variables)

    # -- nulls ------------------------------------------------------------------------
    def null_count(self, variables=None):
        variables = self.names if variables is None else list(variables)
        return sum(self._col(v).null_count for v in variables)

    def combined_valid(self, variables=None):
        """Boolean mask of rows with no null in any selected column, or None (dataset.cpp:208-235)."""
        variables = self.names if variables is None else list(variables)
        mask = None
        for v in variables:
            col = self._col(v)
            if col.null_count:
                valid = np.asarray(col.is_valid())
                mask = valid if mask is None else (mask & valid)
        return mask

    def valid_rows(self, variables):
        mask = self.combined_valid(variables)
        return self.num_rows if mask is None else int(mask.sum())

    # -- columns ----------------------------------------------------------------------
    def column_numpy(self, name):
        col = self._col(name)
        if col.null_count:
            return col.to_numpy(zero_copy_only=False)
        return col.to_numpy(zero_copy_only=True)

    def dense_columns(self, variables, mask=None):
        """to_eigen (dataset.hpp:236-338): contiguous columns, null rows removed."""
        if mask is None:
            mask = self.combined_valid(variables)
        cols = []
        for v in variables:
            a = self.column_numpy(v)
            cols.append(a if mask is None else a[mask])
        return cols, mask

    def to_numpy(self, variables, mask=None):
        cols, mask = self.dense_columns(variables, mask)
        return np.asfortranarray(np.column_stack(cols)) if cols else np.empty((0, 0))

    # -- device residency -----------------------------------------------------------------
    def device_table(self, variables, ctx=None, mask=None, cache=True):
        """(DeviceTable, column indices, valid-row mask or None) for the selected variables.

        Without nulls all floating columns of the frame are uploaded once and shared by
        every later call on this wrapper; with nulls a compacted table for exactly this
        variable set (and null pattern) is uploaded and cached.
        """
        ctx = ctx or _lib.default_context()
        variables = list(variables)
        code = self.dtype_code(variables)
        own_mask = mask is None
        if mask is None:
            mask = self.combined_valid(variables)
        if mask is None:
            key = ("all", code, id(ctx))
            entry = self._tables.get(key)
            if entry is None:
                names = [n for n in self.names if _DTYPE_CODE.get(self._col(n).type) == code and self._col(n).null_count == 0]
                cols = [self.column_numpy(n) for n in names]
                entry = (DeviceTable(ctx, cols, code), {n: i for i, n in enumerate(names)})
                if cache:
                    self._tables[key] = entry
            tbl, index = entry
            return tbl, [index[v] for v in variables], None
        key = ("sel", tuple(variables), code, id(ctx)) if own_mask else None
        entry = self._tables.get(key) if key else None
        if entry is None:
            cols, _ = self.dense_columns(variables, mask)
            entry = DeviceTable(ctx, cols, code)
            if cache and key:
                self._tables[key] = entry
        return entry, list(range(len(variables))), mask

    def device_table_all(self, code, ctx=None):
        """(DeviceTable, {column name: index}) holding EVERY column of dtype `code`, null slots filled with
        NaN.  Used by the cross-validation scores, which address rows through an explicit index list
        that already excludes the rows with nulls (crossvalidation_adaptator.hpp:24-37)."""
        ctx = ctx or _lib.default_context()
        if all(self._col(n).null_count == 0 for n in self.names if _DTYPE_CODE.get(self._col(n).type) == code):
            tbl, _, _ = self.device_table([n for n in self.names if _DTYPE_CODE.get(self._col(n).type) == code][:1], ctx)
            return tbl, dict(self._tables[("all", code, id(ctx))][1])
        key = ("allnull", code, id(ctx))
        entry = self._tables.get(key)
        if entry is None:
            names = [n for n in self.names if _DTYPE_CODE.get(self._col(n).type) == code]
            cols = [np.asarray(self.column_numpy(n), dtype=_NP_DTYPE[code]) for n in names]
            entry = (DeviceTable(ctx, cols, code), {n: i for i, n in enumerate(names)})
            self._tables[key] = entry
        return entry[0], dict(entry[1])

    @property
    def num_columns(self):
        return len(self.names)

    def column_name(self, col):
        if isinstance(col, (int, np.integer)):
            if col < 0 or col >= len(self.names):
                raise IndexError("Column index " + str(col) + " do not exist in DataFrame.")
            return self.names[int(col)]
        return col

    def valid_row_indices(self, include_null=False):
        """Row ids with no null in ANY column (all rows if include_null), as int32."""
        mask = None if include_null else self.combined_valid()
        if mask is None:
            return np.arange(self.num_rows, dtype=np.int32)
        return np.flatnonzero(mask).astype(np.int32)

    def take(self, indices):
        indices = pa.array(np.asarray(indices, dtype=np.int32))
        return DataFrame(self.rb.take(indices))

    def loc(self, variables):
        if isinstance(variables, (str, int, np.integer)):
            variables = [variables]
        variables = [self.column_name(v) for v in variables]
        return DataFrame(pa.RecordBatch.from_arrays([self._col(v) for v in variables], names=list(variables)))

    def to_pandas(self):
        return self.rb.to_pandas()


def _random_seed(seed):
    """util::random_seed_arg: std::random_device when no seed is given."""
    if seed is None:
        import secrets
        return secrets.randbits(32)
    seed = int(seed)
    if seed < 0 or seed > 0xFFFFFFFF:
        raise TypeError("seed must be an unsigned 32-bit integer")
    return seed


def _i32p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))


class _CVProperties:
    """dataset::CrossValidationProperties (crossvalidation_adaptator.hpp:15-67): shuffled valid-row
    indices + fold limits, computed by libstdc++'s std::shuffle / std::mt19937 (pbn_cv_split)."""

    def __init__(self, frame, k, seed, include_null):
        k = int(k)
        self.k = k
        self.seed = seed
        if k <= 1 or k > frame.num_rows:
            raise ValueError("Cannot split %d instances into %d folds." % (frame.num_rows, k))
        self.indices = np.ascontiguousarray(frame.valid_row_indices(include_null))
        self.limits = np.empty(k + 1, dtype=np.int32)
        check(lib().pbn_cv_split(_i32p(self.indices), self.indices.size, k, ctypes.c_uint32(seed), _i32p(self.limits)))


class CrossValidation:
    """pybnesian.CrossValidation (dataset/crossvalidation_adaptator.{hpp,cpp}, pybindings_dataset.cpp:13-113)."""

    def __init__(self, df, k=10, seed=None, include_null=False, _prop=None):
        self._frame = DataFrame.wrap(df)
        self._prop = _prop if _prop is not None else _CVProperties(self._frame, k, _random_seed(seed), bool(include_null))

    @property
    def k(self):
        return self._prop.k

    def data(self):
        return self._frame

    def _fold_indices(self, fold):
        p = self._prop
        if fold < 0 or fold >= p.k:
            raise IndexError("fold index out of range")
        a, b = int(p.limits[fold]), int(p.limits[fold + 1])
        train = np.concatenate([p.indices[:a], p.indices[b:int(p.limits[-1])]])
        return train, p.indices[a:b]

    def fold(self, index):
        train, test = self._fold_indices(int(index))
        return self._frame.take(train).rb, self._frame.take(test).rb

    def __iter__(self):
        for f in range(self._prop.k):
            yield self.fold(f)

    def indices(self):
        for f in range(self._prop.k):
            train, test = self._fold_indices(f)
            yield train.tolist(), test.tolist()

    def loc(self, columns):
        return CrossValidation(self._frame.loc(columns), _prop=self._prop)


class HoldOut:
    """pybnesian.HoldOut (dataset/holdout_adaptator.{hpp,cpp}, pybindings_dataset.cpp:116-146)."""

    def __init__(self, df, test_ratio=0.2, seed=None, include_null=False):
        self._frame = DataFrame.wrap(df)
        self.seed = _random_seed(seed)
        test_ratio = float(test_ratio)
        if test_ratio <= 0 or test_ratio >= 1.0:
            raise ValueError("test_ratio must be a number between 0 and 1.")
        idx = np.ascontiguousarray(self._frame.valid_row_indices(bool(include_null)))
        ntr = ctypes.c_int32()
        check(lib().pbn_holdout_split(_i32p(idx), idx.size, test_ratio, ctypes.c_uint32(self.seed), ctypes.byref(ntr)))
        self.train_indices = idx[:ntr.value]
        self.test_indices = idx[ntr.value:]
        self._train = self._frame.take(self.train_indices)
        self._test = self._frame.take(self.test_indices)

    def training_data(self):
        return self._train.rb

    def test_data(self):
        return self._test.rb

    def training_frame(self):
        return self._train

    def test_frame(self):
        return self._test
