"""Likelihood scores on the hot path: CVLikelihood, HoldoutLikelihood, ValidatedLikelihood.

Mirrors learning/scores/{scores,cv_likelihood,holdout_likelihood,validated_likelihood}.{hpp,cpp} and
pybindings_scores.cpp:27-275, 555-658 (same constructor arguments, methods and error messages).
Where the reference evaluates one (candidate, fold) at a time - Arrow Take, upload, 4 launches per
test row - every score class here owns a `pbn_cv` (include/pbn_cuda.h): the data set is resident on
the GPU in shuffled order and `local_score_batch` scores any number of candidate CPDs x all folds
with one whitening launch and one pair-kernel launch per distinct family size.  `local_score` is
the batch of one.  With several GPUs (pybnesian_b200.parallel) the candidates of a batch are dealt
over the ranks and the per-candidate scores are summed with one all-reduce.
"""
import ctypes

import numpy as np
import pyarrow as pa

from . import _lib, parallel
from ._lib import CVItem, check, lib
from .dataset import CrossValidation, DataFrame, HoldOut, _i32p
from .factors import CKDEType, FactorType, LinearGaussianCPDType
from .kde import NormalReferenceRule, ScottsBandwidth


class Args:
    """factors/arguments.hpp: positional construction arguments."""

    def __init__(self, *args):
        self.args = tuple(args)


class Kwargs:
    def __init__(self, **kwargs):
        self.kwargs = dict(kwargs)


class Arguments:
    """factors::Arguments (factors/arguments.hpp:36-140): constructor arguments of the factors a score
    creates, keyed by node name, FactorType or (name, FactorType)."""

    def __init__(self, dict_arguments=None):
        self._name, self._type, self._name_type = {}, {}, {}
        for key, val in (dict_arguments or {}).items():
            parsed = self._process(val)
            if isinstance(key, str):
                self._name[key] = parsed
            elif isinstance(key, FactorType):
                self._type[key] = parsed
            elif isinstance(key, tuple) and len(key) == 2 and isinstance(key[0], str) and isinstance(key[1], FactorType):
                self._name_type[key] = parsed
            else:
                raise ValueError("Key value is not of type str, FactorType or 2-tuple (str, FactorType).")

    @staticmethod
    def _process(o):
        if isinstance(o, tuple):
            if len(o) == 2 and isinstance(o[0], Args) and isinstance(o[1], Kwargs):
                return o[0].args, o[1].kwargs
            return tuple(o), {}
        if isinstance(o, dict):
            return (), dict(o)
        if isinstance(o, Args):
            return o.args, {}
        if isinstance(o, Kwargs):
            return (), o.kwargs
        raise ValueError("The provided arguments must be a 2-tuple (Args(...), Kwargs(...)), an Args(...) (or tuple) "
                         "or a Kwargs(...) (or dict).")

    def args(self, name, factor_type):
        for table, key in ((self._name_type, (name, factor_type)), (self._name, name), (self._type, factor_type)):
            if key in table:
                return table[key]
        return (), {}


class Score:
    """learning/scores/scores.hpp:14-48."""

    def score(self, model):
        return sum(self.local_score(model, node) for node in model.nodes())

    def local_score(self, model, variable, evidence=None):
        raise NotImplementedError

    def local_score_node_type(self, model, variable_type, variable, evidence):
        raise NotImplementedError

    def local_score_batch(self, model, requests):
        """requests: [(FactorType or None, variable, evidence list)] -> list of local scores.  Scores that can
        batch (the likelihood scores below) override this; the default is the reference's serial loop."""
        out = []
        for t, v, e in requests:
            out.append(self.local_score(model, v, e) if t is None else self.local_score_node_type(model, t, v, e))
        return out

    def data(self):
        return DataFrame.wrap(__import__("pyarrow").RecordBatch.from_arrays([], names=[]))

    def has_variables(self, cols):
        return self.data().has_columns(cols)

    def compatible_bn(self, model):
        return self.has_variables(model.nodes())

    def __str__(self):
        return type(self).__name__

    __repr__ = __str__


class ValidatedScore(Score):
    """learning/scores/scores.hpp:50-78."""

    def vscore(self, model):
        return sum(self.vlocal_score(model, node) for node in model.nodes())

    def vlocal_score(self, model, variable, evidence=None):
        raise NotImplementedError

    def vlocal_score_node_type(self, model, variable_type, variable, evidence):
        raise NotImplementedError

    def vlocal_score_batch(self, model, requests):
        out = []
        for t, v, e in requests:
            out.append(self.vlocal_score(model, v, e) if t is None else self.vlocal_score_node_type(model, t, v, e))
        return out


_NATIVE_RULE = {NormalReferenceRule: _lib.BW_NORMAL_REFERENCE, ScottsBandwidth: _lib.BW_SCOTT}


class _FoldScorer:
    """Device side shared by the three scores: a pbn_cv over `frame` with the given shuffled row ids and
    fold limits, scoring folds [fold_begin, fold_end)."""

    def __init__(self, frame, indices, limits, fold_begin, fold_end, arguments):
        self.frame = frame
        self.indices = np.ascontiguousarray(indices, dtype=np.int32)
        self.limits = np.ascontiguousarray(limits, dtype=np.int32)
        self.fold_begin, self.fold_end = int(fold_begin), int(fold_end)
        self.arguments = arguments
        self._dev = {}     # dtype code -> (handle owner, column index)
        self._memo = {}    # (type, variable, ordered evidence) -> score
        self.stats = {"requests": 0, "memo_hits": 0, "device_items": 0, "host_items": 0, "batches": 0}

    class _Handle:
        def __init__(self, h, tbl):
            self.h, self.tbl = h, tbl

        def __del__(self):
            try:
                if self.h:
                    lib().pbn_cv_free(self.h)
                    self.h = None
            except Exception:
                pass

    def _device(self, code):
        entry = self._dev.get(code)
        if entry is None:
            tbl, index = self.frame.device_table_all(code)
            h = ctypes.c_void_p()
            check(lib().pbn_cv_create(tbl.ctx.handle, tbl.handle, _i32p(self.indices), self.indices.size,
                                      _i32p(self.limits), self.limits.size - 1, ctypes.byref(h)))
            entry = (self._Handle(h, tbl), index)
            self._dev[code] = entry
        return entry

    # folds as DataFrames, for factors that only exist in Python (the reference's generic loop)
    def _fold_frames(self):
        for f in range(self.fold_begin, self.fold_end):
            a, b = int(self.limits[f]), int(self.limits[f + 1])
            train = np.concatenate([self.indices[:a], self.indices[b:int(self.limits[-1])]])
            yield self.frame.take(train), self.frame.take(self.indices[a:b])

    def _native_item(self, node_type, variable, evidence, args, kwargs):
        """(dtype code, CVItem fields) if the request can run inside pbn_cv_scores, else None."""
        variables = [variable] + list(evidence)
        if len(variables) > _lib.PBN_MAX_DIM:
            return None
        # categorical columns (discrete nodes, HCKDE / CLinearGaussianCPD families): generic loop
        if any(pa.types.is_dictionary(self.frame._col(v).type) for v in variables if v in self.frame._index):
            return None
        if node_type == LinearGaussianCPDType():
            if args or kwargs:
                return None
            rule, factor = 0, _lib.FACTOR_LINEAR_GAUSSIAN
        elif node_type == CKDEType():
            selector = args[0] if args else kwargs.get("bandwidth_selector")
            if len(args) > 1 or (set(kwargs) - {"bandwidth_selector"}):
                return None
            if selector is None:
                rule = _lib.BW_NORMAL_REFERENCE
            elif type(selector) in _NATIVE_RULE:
                rule = _NATIVE_RULE[type(selector)]
            else:
                return None  # UCV or a Python BandwidthSelector: generic loop
            factor = _lib.FACTOR_CKDE
        else:
            return None
        what = "fit KDE" if factor == _lib.FACTOR_CKDE else "fit LinearGaussianCPD"
        code = self.frame.dtype_code(variables, what)
        return code, factor, rule, variables

    def score_batch(self, model, requests):
        n = len(requests)
        out = [None] * n
        pending = {}   # key -> list of positions
        order = []
        self.stats["requests"] += n
        for pos, (node_type, variable, evidence) in enumerate(requests):
            evidence = list(evidence)
            key = (node_type, variable, tuple(evidence))
            if key in self._memo:
                out[pos] = self._memo[key]
                self.stats["memo_hits"] += 1
            elif key in pending:
                pending[key].append(pos)
            else:
                pending[key] = [pos]
                order.append(key)
        if not order:
            return out
        native = {}  # code -> list of (key, factor, rule, variables)
        generic = []
        for key in order:
            node_type, variable, evidence = key
            args, kwargs = self.arguments.args(variable, node_type)
            item = self._native_item(node_type, variable, evidence, args, kwargs)
            if item is None:
                generic.append((key, args, kwargs))
            else:
                native.setdefault(item[0], []).append((key,) + item[1:])
        results = {}
        for code, items in native.items():
            results.update(self._score_native(code, items))
        for key, args, kwargs in generic:
            node_type, variable, evidence = key
            cpd = node_type.new_factor(model, variable, list(evidence), *args, **kwargs)
            loglik = 0.0
            for train, test in self._fold_frames():
                cpd.fit(train.loc([variable] + list(evidence)))
                loglik += cpd.slogl(test.loc([variable] + list(evidence)))
            results[key] = loglik
            self.stats["host_items"] += 1
        for key in order:
            self._memo[key] = results[key]
            for pos in pending[key]:
                out[pos] = results[key]
        return out

    # several ranks: deal (item, fold) jobs instead of whole items (a hill-climbing step has only 20-40 candidates of
    # very different cost).  Subclasses whose _run_items cannot score a fold range set this to False.
    _fold_parallel = True

    def _score_native(self, code, items):
        rank, world = parallel.rank(), parallel.world_size()
        nfolds = self.fold_end - self.fold_begin
        if world > 1 and self._fold_parallel and nfolds > 1:
            return self._score_native_by_fold(code, items, rank, world)
        # deal the items over the ranks, most expensive first (CKDE cost grows with the family size; LG is free)
        cost = [len(v) if f == _lib.FACTOR_CKDE else 0 for _, f, _, v in items]
        mine = parallel.deal(cost, rank, world) if world > 1 else list(range(len(items)))
        scores = np.zeros(len(items))
        with parallel.guard() as g:  # e.g. SingularCovarianceData of an item only this rank was dealt
            if mine:
                scores[mine] = self._run_items(code, [items[i] for i in mine])
                self.stats["device_items"] += len(mine)
                self.stats["batches"] += 1
        if world > 1:
            scores = parallel.all_reduce_sum(scores, self._ctx(code), error=g.error)
        return {items[i][0]: float(scores[i]) for i in range(len(items))}

    def _score_native_by_fold(self, code, items, rank, world):
        """(item, fold) jobs dealt over the ranks, most expensive first; every rank scores ITS jobs in one
        pbn_cv_score_jobs call (one batched launch per family size, however the jobs spread over folds), the
        [items x folds] matrix is summed over ranks (each entry written by exactly one rank) and the folds of an item
        are added in fold order - the same additions as the single-GPU call."""
        nfolds = self.fold_end - self.fold_begin
        cost = []
        for _, f, _, v in items:
            cost.extend([len(v) if f == _lib.FACTOR_CKDE else 0] * nfolds)
        mine = parallel.deal(cost, rank, world)
        mat = np.zeros((len(items), nfolds))
        with parallel.guard() as g:
            if mine:
                mine = sorted(mine)
                ji = [job // nfolds for job in mine]
                jq = [job % nfolds for job in mine]
                mat[ji, jq] = self._run_jobs(code, items, ji, [self.fold_begin + q for q in jq])
                self.stats["device_items"] += len(mine)
                self.stats["batches"] += 1
        mat = parallel.all_reduce_sum(mat.ravel(), self._ctx(code), error=g.error).reshape(len(items), nfolds)
        out = {}
        for i in range(len(items)):
            total = 0.0
            for q in range(nfolds):
                total += float(mat[i, q])
            out[items[i][0]] = total
        return out

    def _run_jobs(self, code, items, job_item, job_fold):
        """One pbn_cv_score_jobs call: job j = fold job_fold[j] of items[job_item[j]]; returns the per-job slogl."""
        handle, index = self._device(code)
        arr = self._item_array(items, index)
        ji = np.ascontiguousarray(job_item, dtype=np.int32)
        jf = np.ascontiguousarray(job_fold, dtype=np.int32)
        local = np.zeros(len(ji))
        status = (ctypes.c_int * len(items))()
        check(lib().pbn_cv_score_jobs(handle.tbl.ctx.handle, handle.h, arr, len(items), _i32p(ji), _i32p(jf), len(ji),
                                      local.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), status))
        used = set(int(i) for i in ji)
        bad = [status[i] for i in sorted(used) if status[i] != _lib.PBN_OK]
        if bad:
            _lib.raise_for(bad[0], lib().pbn_last_error().decode("utf-8", "replace"))
        return local

    @staticmethod
    def _item_array(items, index):
        arr = (CVItem * len(items))()
        for slot, (_, factor, rule, variables) in enumerate(items):
            arr[slot].factor, arr[slot].rule, arr[slot].n_vars = factor, rule, len(variables)
            for q, v in enumerate(variables):
                arr[slot].vars[q] = index[v]
        return arr

    def _ctx(self, code):
        return self._device(code)[0].tbl.ctx

    def _run_items(self, code, items, fold_begin=None, fold_end=None):
        """One pbn_cv_scores call for `items` = [(key, factor, rule, variables)] over folds [fold_begin, fold_end)
        (default: all folds of this scorer); returns their scores."""
        fold_begin = self.fold_begin if fold_begin is None else fold_begin
        fold_end = self.fold_end if fold_end is None else fold_end
        handle, index = self._device(code)
        arr = self._item_array(items, index)
        local = np.zeros(len(items))
        status = (ctypes.c_int * len(items))()
        check(lib().pbn_cv_scores(handle.tbl.ctx.handle, handle.h, arr, len(items), fold_begin, fold_end,
                                  local.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), status))
        bad = [s for s in status if s != _lib.PBN_OK]
        if bad:
            _lib.raise_for(bad[0], lib().pbn_last_error().decode("utf-8", "replace"))
        return local


class CVLikelihood(Score):
    """pybnesian.CVLikelihood (learning/scores/cv_likelihood.{hpp,cpp}; pybindings_scores.cpp)."""

    def __init__(self, df, k=10, seed=None, construction_args=None):
        self._cv = CrossValidation(df, k, seed)
        self._arguments = construction_args if construction_args is not None else Arguments()
        p = self._cv._prop
        self._scorer = _FoldScorer(self._cv.data(), p.indices, p.limits, 0, p.k, self._arguments)

    @property
    def cv(self):
        return self._cv

    def data(self):
        return self._cv.data()

    def local_score(self, model, variable, evidence=None):
        if evidence is None:
            evidence = model.parents(variable)
        return self.local_score_node_type(model, model.underlying_node_type(self.data(), variable), variable, evidence)

    def local_score_node_type(self, model, variable_type, variable, evidence):
        return self._scorer.score_batch(model, [(variable_type, variable, list(evidence))])[0]

    def local_score_batch(self, model, requests):
        reqs = [(t if t is not None else model.underlying_node_type(self.data(), v), v, list(e)) for t, v, e in requests]
        return self._scorer.score_batch(model, reqs)

    def __str__(self):
        return "CVLikelihood"

    __repr__ = __str__


class HoldoutLikelihood(Score):
    """pybnesian.HoldoutLikelihood (learning/scores/holdout_likelihood.{hpp,cpp})."""

    def __init__(self, df, test_ratio=0.2, seed=None, construction_args=None):
        self._holdout = HoldOut(df, test_ratio, seed)
        self._arguments = construction_args if construction_args is not None else Arguments()
        h = self._holdout
        indices = np.concatenate([h.train_indices, h.test_indices])
        limits = np.array([0, h.train_indices.size, indices.size], dtype=np.int32)
        # a holdout is the 2-fold structure scored on fold 1 only: train = fold 0, test = fold 1
        self._scorer = _FoldScorer(h._frame, indices, limits, 1, 2, self._arguments)

    @property
    def holdout(self):
        return self._holdout

    def training_data(self):
        return self._holdout.training_data()

    def test_data(self):
        return self._holdout.test_data()

    def data(self):
        return self._holdout.training_frame()

    def local_score(self, model, variable, evidence=None):
        if evidence is None:
            evidence = model.parents(variable)
        return self.local_score_node_type(model, model.underlying_node_type(self.data(), variable), variable, evidence)

    def local_score_node_type(self, model, variable_type, variable, evidence):
        return self._scorer.score_batch(model, [(variable_type, variable, list(evidence))])[0]

    def local_score_batch(self, model, requests):
        reqs = [(t if t is not None else model.underlying_node_type(self.data(), v), v, list(e)) for t, v, e in requests]
        return self._scorer.score_batch(model, reqs)

    def __str__(self):
        return "HoldoutLikelihood"

    __repr__ = __str__


class ValidatedLikelihood(ValidatedScore):
    """pybnesian.ValidatedLikelihood (learning/scores/validated_likelihood.hpp:12-75): a holdout split, then
    k-fold CV on the holdout's training part with the same seed."""

    def __init__(self, df, test_ratio=0.2, k=10, seed=None, construction_args=None):
        from .dataset import _random_seed
        seed = _random_seed(seed)
        self._holdout = HoldoutLikelihood(df, test_ratio, seed, construction_args)
        self._cv = CVLikelihood(self._holdout.holdout.training_frame(), k, seed, construction_args)

    @property
    def holdout_lik(self):
        return self._holdout

    @property
    def cv_lik(self):
        return self._cv

    def training_data(self):
        return self._holdout.training_data()

    def validation_data(self):
        return self._holdout.test_data()

    def data(self):
        return self._cv.data()

    def local_score(self, model, variable, evidence=None):
        return self._cv.local_score(model, variable, evidence)

    def local_score_node_type(self, model, variable_type, variable, evidence):
        return self._cv.local_score_node_type(model, variable_type, variable, evidence)

    def local_score_batch(self, model, requests):
        return self._cv.local_score_batch(model, requests)

    def vlocal_score(self, model, variable, evidence=None):
        return self._holdout.local_score(model, variable, evidence)

    def vlocal_score_node_type(self, model, variable_type, variable, evidence):
        return self._holdout.local_score_node_type(model, variable_type, variable, evidence)

    def vlocal_score_batch(self, model, requests):
        return self._holdout.local_score_batch(model, requests)

    def __str__(self):
        return "ValidatedLikelihood"

    __repr__ = __str__


class BIC(Score):
    """pybnesian.BIC (learning/scores/bic.{hpp,cpp}): closed-form Bayesian information criterion of linear Gaussian,
    conditional linear Gaussian and discrete nodes.  The linear-Gaussian fits run on the resident table
    (pbn_lg_fit through MLELinearGaussianCPD); everything else is host integer / scalar work as in the reference.
    The default score of GaussianNetwork hill climbing (util/validate_options.cpp:36-44)."""

    _MACHINE_TOL = float(np.sqrt(np.finfo(np.float64).eps))

    def __init__(self, df):
        self._frame = DataFrame.wrap(df)

    def data(self):
        return self._frame

    def _bic_lineargaussian(self, frame, variable, parents):
        from .factors import MLELinearGaussianCPD
        p = MLELinearGaussianCPD().estimate(frame, variable, parents)
        if p.variance < self._MACHINE_TOL or np.isinf(p.variance):
            return None
        rows = frame.valid_rows([variable] + list(parents))
        k = len(parents)
        return 0.5 * (1 + float(k) - float(rows)) - 0.5 * rows * np.log(2 * np.pi) - rows * 0.5 * np.log(p.variance), rows

    def _bic_clg(self, variable, discrete_parents, continuous_parents):
        from . import hybrid
        frame = self._frame
        card, strides = hybrid.create_cardinality_strides(frame, discrete_parents)
        num_configs = int(card.prod())
        order, offsets = hybrid.discrete_slices(frame, discrete_parents, strides, num_configs)
        sub = frame.loc([variable] + list(continuous_parents))
        loglik = 0.0
        for i in range(num_configs):
            rows = order[int(offsets[i]):int(offsets[i + 1])]
            if rows.size == 0:
                continue
            r = self._bic_lineargaussian(sub.take(rows), variable, continuous_parents)
            if r is None:
                return -np.inf
            loglik += r[0]
        valid = frame.valid_rows([variable] + list(discrete_parents) + list(continuous_parents))
        return loglik - np.log(valid) * 0.5 * num_configs * (len(continuous_parents) + 2)

    def _bic_discrete(self, variable, parents):
        from . import hybrid
        f = hybrid.DiscreteFactor(variable, parents)
        card, strides = hybrid.create_cardinality_strides(self._frame, [variable] + list(parents))
        idx, _ = f._indices(self._frame, strides)
        counts = np.bincount(idx, minlength=int(card.prod())).astype(np.int64).reshape(-1, int(card[0]))
        ll = 0.0
        for row in counts:  # per parent configuration, categories in order (bic.cpp:73-92)
            tot = int(row.sum())
            if tot > 0:
                inv = 1.0 / tot
                for c in row:
                    if c > 0:
                        ll += float(c) * np.log(float(c) * inv)
        return ll - np.log(float(counts.sum())) * 0.5 * (int(card[0]) - 1) * counts.shape[0]

    def local_score(self, model, variable, evidence=None):
        if evidence is None:
            evidence = model.parents(variable)
        return self.local_score_node_type(model, model.underlying_node_type(self._frame, variable), variable, evidence)

    def local_score_node_type(self, model, variable_type, variable, evidence):
        from .factors import LinearGaussianCPDType
        from .hybrid import DiscreteFactorType
        evidence = list(evidence)
        if variable_type == LinearGaussianCPDType():
            disc = [p for p in evidence if model.underlying_node_type(self._frame, p) == DiscreteFactorType()]
            cont = [p for p in evidence if p not in disc]
            if not disc:
                r = self._bic_lineargaussian(self._frame, variable, evidence)
                if r is None:
                    return -np.inf
                return float(r[0] - np.log(r[1]) * 0.5 * (len(evidence) + 2))
            return float(self._bic_clg(variable, disc, cont))
        if variable_type == DiscreteFactorType():
            if any(model.underlying_node_type(self._frame, p) != DiscreteFactorType() for p in evidence):
                raise ValueError("Local score for discrete variable " + variable + " cannot be calculated because the "
                                 "parents/evidence contains non-discrete variables.")
            return float(self._bic_discrete(variable, evidence))
        raise ValueError("Bayesian network type \"" + str(model.type()) + "\" not valid for score BIC")

    def __str__(self):
        return "BIC"
