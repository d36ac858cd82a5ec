"""Hill-climbing operators: AddArc / RemoveArc / FlipArc / ChangeNodeType and the operator sets.

Mirrors learning/operators/operators.{hpp,cpp} (pybindings_operators.cpp:860-904 for the Python
surface).  The candidate enumeration order, the parent-list edits (`swap_remove_v` / `push_back`),
the persistent `sorted_idx` vector and its `std::sort` (pbn_sort_desc) are reproduced exactly, because
the selected operator sequence has to be identical to the reference's.  What changes is WHEN the
scores are computed: every set first lists the local scores it needs
(`_cache_requests` / `_update_requests`), the whole list goes to `Score.local_score_batch` in one
call - one fused GPU launch per family size for the likelihood scores - and only then are the deltas
filled in, in the reference's order and with its arithmetic.
"""
import ctypes
import sys

import numpy as np

from ._lib import check, lib
from .factors import UnknownFactorType

LOWEST = -sys.float_info.max  # std::numeric_limits<double>::lowest()


def _swap_remove(v, value):
    """util::swap_remove_v (util/vector.hpp:13-27)."""
    i = v.index(value)
    v[i] = v[-1]
    v.pop()


class Operator:
    """operators.hpp:22-67."""

    def __init__(self, delta):
        self._delta = float(delta)

    def delta(self):
        return self._delta

    def apply(self, model):
        raise NotImplementedError

    def nodes_changed(self, model):
        raise NotImplementedError

    def opposite(self, model):
        raise NotImplementedError

    def _key(self):
        raise NotImplementedError

    def __eq__(self, other):
        return isinstance(other, Operator) and type(self) is type(other) and self._key() == other._key()

    def __ne__(self, other):
        return not self == other

    def __hash__(self):
        return hash((type(self).__name__,) + self._key())

    def __repr__(self):
        return str(self)


class ArcOperator(Operator):
    def __init__(self, source, target, delta):
        super().__init__(delta)
        self._source, self._target = source, target

    def source(self):
        return self._source

    def target(self):
        return self._target

    def _key(self):
        return (self._source, self._target)

    def __str__(self):
        return "%s(%s -> %s; Delta: %f)" % (type(self).__name__, self._source, self._target, self._delta)


class AddArc(ArcOperator):
    def apply(self, model):
        model.add_arc_unsafe(self._source, self._target)

    def nodes_changed(self, model):
        return [self._target]

    def opposite(self, model):
        return RemoveArc(self._source, self._target, -self._delta)


class RemoveArc(ArcOperator):
    def apply(self, model):
        model.remove_arc(self._source, self._target)

    def nodes_changed(self, model):
        return [self._target]

    def opposite(self, model):
        return AddArc(self._source, self._target, -self._delta)


class FlipArc(ArcOperator):
    def apply(self, model):
        model.flip_arc_unsafe(self._source, self._target)

    def nodes_changed(self, model):
        return [self._source, self._target]

    def opposite(self, model):
        return FlipArc(self._target, self._source, -self._delta)


class ChangeNodeType(Operator):
    def __init__(self, node, node_type, delta):
        super().__init__(delta)
        self._node, self._node_type = node, node_type

    def node(self):
        return self._node

    def node_type(self):
        return self._node_type

    def apply(self, model):
        model.set_node_type(self._node, self._node_type)

    def nodes_changed(self, model):
        return [self._node]

    def opposite(self, model):
        return ChangeNodeType(self._node, model.node_type(self._node), -self._delta)

    def _key(self):
        return (self._node, self._node_type)

    def __str__(self):
        return "ChangeNodeType(%s -> %s; Delta: %f)" % (self._node, self._node_type, self._delta)


class OperatorTabuSet:
    """operators.hpp:252-293."""

    def __init__(self):
        self._set = set()

    def insert(self, op):
        self._set.add(op)

    def contains(self, op):
        return op in self._set

    def clear(self):
        self._set.clear()

    def empty(self):
        return not self._set


class LocalScoreCache:
    """operators.hpp:295-338."""

    def __init__(self, model=None):
        self._scores = np.zeros(model.num_nodes() if model is not None else 0)

    def _resize(self, model):
        if self._scores.size != model.num_nodes():
            self._scores = np.zeros(model.num_nodes())

    def cache_local_scores(self, model, score):
        self._resize(model)
        nodes = model.nodes()
        vals = score.local_score_batch(model, [(None, n, model.parents(n)) for n in nodes])
        for n, v in zip(nodes, vals):
            self._scores[model.collapsed_index(n)] = v

    def cache_vlocal_scores(self, model, score):
        self._resize(model)
        nodes = model.nodes()
        vals = score.vlocal_score_batch(model, [(None, n, model.parents(n)) for n in nodes])
        for n, v in zip(nodes, vals):
            self._scores[model.collapsed_index(n)] = v

    def update_local_score(self, model, score, variable):
        self._scores[model.collapsed_index(variable)] = score.local_score(model, variable)

    def update_vlocal_score(self, model, score, variable):
        self._scores[model.collapsed_index(variable)] = score.vlocal_score(model, variable)

    def sum(self):
        return float(self._scores.sum())

    def local_score(self, model, name):
        return float(self._scores[model.collapsed_index(name)])


class _Plan:
    """Score requests of one cache/update step plus the code that turns the answers into deltas."""

    def __init__(self):
        self.requests = []
        self.finish = []  # callables taking the list of answers (for this plan's own slice)

    def ask(self, node_type, variable, evidence):
        self.requests.append((node_type, variable, list(evidence)))
        return len(self.requests) - 1


def _run_plans(model, score, plans):
    """One batched score call for several plans; their finishers run in order."""
    requests = []
    offsets = []
    for p in plans:
        offsets.append(len(requests))
        requests.extend(p.requests)
    answers = score.local_score_batch(model, requests) if requests else []
    for p, off in zip(plans, offsets):
        local = answers[off:off + len(p.requests)]
        for fn in p.finish:
            fn(local)


class OperatorSet:
    """operators.hpp:340-436."""

    def __init__(self):
        self._local_cache = None
        self._owns_local_cache = False

    def set_local_score_cache(self, cache):
        self._local_cache = cache
        self._owns_local_cache = False

    def local_score_cache(self):
        return self._local_cache

    def _initialize_local_cache(self, model):
        if self._local_cache is None:
            self._local_cache = LocalScoreCache(model)
            self._owns_local_cache = True

    def _raise_uninitialized(self):
        if self._local_cache is None:
            raise ValueError("Local cache not initialized. Call cache_scores() before find_max()")

    def set_arc_blacklist(self, blacklist):
        pass

    def set_arc_whitelist(self, whitelist):
        pass

    def set_max_indegree(self, indegree):
        pass

    def set_type_blacklist(self, blacklist):
        pass

    def set_type_whitelist(self, whitelist):
        pass

    def finished(self):
        self._local_cache = None

    # the reference's entry points, expressed through plans
    def cache_scores(self, model, score):
        plans = self._cache_plans(model, score)
        _run_plans(model, score, plans)

    def update_scores(self, model, score, variables):
        plans = self._update_plans(model, score, variables)
        _run_plans(model, score, plans)

    def _local_cache_plan(self, model, nodes):
        """(Re)computes the cached local score of `nodes` (LocalScoreCache::cache/update_local_score)."""
        plan = _Plan()
        cache = self._local_cache
        cache._resize(model)
        slots = [(n, plan.ask(None, n, model.parents(n))) for n in nodes]

        def fin(ans):
            for n, i in slots:
                cache._scores[model.collapsed_index(n)] = ans[i]
        plan.finish.append(fin)
        return plan


def _validate_restrictions(model, blacklist, whitelist):
    """util::validate_restrictions (util/validate_whitelists.hpp:152-180)."""
    for lst in (blacklist, whitelist):
        for s, t in lst:
            for n in (s, t):
                if not model.contains_node(n):
                    raise ValueError("Node " + n + " not present in the graph.")
    white = []
    for s, t in whitelist:
        a = (model.index(s), model.index(t))
        if a not in white:
            white.append(a)
    black = []
    for s, t in blacklist:
        a = (model.index(s), model.index(t))
        if a in white:
            raise ValueError("Arc " + s + " -> " + t + " in blacklist and whitelist")
        if a not in black:
            black.append(a)
    return black, white


class ArcOperatorSet(OperatorSet):
    """operators.hpp:438-683, operators.cpp:19-363."""

    def __init__(self, blacklist=None, whitelist=None, max_indegree=0):
        super().__init__()
        self._blacklist = list(blacklist or [])
        self._whitelist = list(whitelist or [])
        self._max_indegree = int(max_indegree)
        self._delta = np.zeros(0)          # flat, index = source + target * num_nodes (column-major MatrixXd)
        self._valid = np.zeros(0, dtype=bool)
        self._sorted_idx = np.zeros(0, dtype=np.int32)
        self._n = 0

    def set_arc_blacklist(self, blacklist):
        self._blacklist = list(blacklist)

    def set_arc_whitelist(self, whitelist):
        self._whitelist = list(whitelist)

    def set_max_indegree(self, indegree):
        self._max_indegree = int(indegree)

    def _d(self, s, t):
        return s + t * self._n

    def update_valid_ops(self, model):
        n = model.num_nodes()
        if self._n != n or self._delta.size != n * n:
            self._delta = np.zeros(n * n)
            self._valid = np.ones(n * n, dtype=bool)
            self._n = n
        self._valid[:] = True
        black, white = _validate_restrictions(model, self._blacklist, self._whitelist)
        # the restrictions hold raw node indices; delta / valid_op are indexed by COLLAPSED indices
        # (operators.cpp:36-52: model.collapsed_from_index), which differ once nodes have been removed
        cfi = model.collapsed_from_index
        for s, t in white:
            s, t = cfi(s), cfi(t)
            self._valid[self._d(s, t)] = False
            self._valid[self._d(t, s)] = False
            self._delta[self._d(s, t)] = LOWEST
            self._delta[self._d(t, s)] = LOWEST
        for s, t in black:
            s, t = cfi(s), cfi(t)
            self._valid[self._d(s, t)] = False
            self._delta[self._d(s, t)] = LOWEST
        for i in range(n):
            self._valid[self._d(i, i)] = False
            self._delta[self._d(i, i)] = LOWEST
        idx = [i + j * n for i in range(n) for j in range(n) if self._valid[i + j * n]]
        self._sorted_idx = np.array(idx, dtype=np.int32)

    def _cache_plans(self, model, score):
        if not score.compatible_bn(model):
            raise ValueError("BayesianNetwork is not compatible with the score.")
        self._initialize_local_cache(model)
        plans = []
        if self._owns_local_cache:
            plans.append(self._local_cache_plan(model, model.nodes()))
        self.update_valid_ops(model)
        plan = _Plan()
        cache = self._local_cache
        bn_type = model.type()
        todo = []
        for target in model.nodes():
            parents_target = model.parents(target)
            tc = model.collapsed_index(target)
            for source in model.nodes():
                sc = model.collapsed_index(source)
                if not (self._valid[self._d(sc, tc)] and bn_type.can_have_arc(model, source, target)):
                    continue
                # cache_score_operation (operators.cpp:71-101)
                if model.has_arc(source, target):
                    _swap_remove(parents_target, source)
                    a = plan.ask(None, target, parents_target)
                    parents_target.append(source)
                    todo.append((sc, tc, "rm", source, target, a, None))
                elif model.has_arc(target, source):
                    new_parents_source = model.parents(source)
                    _swap_remove(new_parents_source, target)
                    parents_target.append(source)
                    a = plan.ask(None, source, new_parents_source)
                    b = plan.ask(None, target, parents_target)
                    parents_target.pop()
                    todo.append((sc, tc, "flip", source, target, a, b))
                else:
                    parents_target.append(source)
                    a = plan.ask(None, target, parents_target)
                    parents_target.pop()
                    todo.append((sc, tc, "add", source, target, a, None))

        def fin(ans):
            for sc, tc, kind, source, target, a, b in todo:
                if kind == "flip":
                    d = ans[a] + ans[b] - cache.local_score(model, source) - cache.local_score(model, target)
                else:
                    d = ans[a] - cache.local_score(model, target)
                self._delta[self._d(sc, tc)] = d
        plan.finish.append(fin)
        plans.append(plan)
        return plans

    def _update_plans(self, model, score, variables):
        self._raise_uninitialized()
        plans = []
        if self._owns_local_cache:
            plans.append(self._local_cache_plan(model, variables))
        plan = _Plan()
        cache = self._local_cache
        bn_type = model.type()
        todo = []
        for target in variables:  # update_incoming_arcs_scores (operators.cpp:296-347)
            tc = model.collapsed_index(target)
            parents = model.parents(target)
            for source in model.nodes():
                sc = model.collapsed_index(source)
                if not self._valid[self._d(sc, tc)]:
                    continue
                if model.has_arc(source, target):
                    _swap_remove(parents, source)
                    a = plan.ask(None, target, parents)
                    parents.append(source)
                    b = None
                    if self._valid[self._d(tc, sc)] and bn_type.can_have_arc(model, target, source):
                        parents_source = model.parents(source)
                        parents_source.append(target)
                        b = plan.ask(None, source, parents_source)
                    todo.append((sc, tc, "rm", source, target, a, b))
                elif model.has_arc(target, source) and bn_type.can_have_arc(model, source, target):
                    parents_source = model.parents(source)
                    _swap_remove(parents_source, target)
                    parents.append(source)
                    a = plan.ask(None, source, parents_source)
                    b = plan.ask(None, target, parents)
                    parents.pop()
                    todo.append((sc, tc, "flip", source, target, a, b))
                elif bn_type.can_have_arc(model, source, target):
                    parents.append(source)
                    a = plan.ask(None, target, parents)
                    parents.pop()
                    todo.append((sc, tc, "add", source, target, a, None))

        def fin(ans):
            for sc, tc, kind, source, target, a, b in todo:
                if kind == "rm":
                    d = ans[a] - cache.local_score(model, target)
                    self._delta[self._d(sc, tc)] = d
                    if b is not None:
                        self._delta[self._d(tc, sc)] = d + ans[b] - cache.local_score(model, source)
                elif kind == "flip":
                    self._delta[self._d(sc, tc)] = ans[a] + ans[b] - cache.local_score(model, source) - \
                        cache.local_score(model, target)
                else:
                    self._delta[self._d(sc, tc)] = ans[a] - cache.local_score(model, target)
        plan.finish.append(fin)
        plans.append(plan)
        return plans

    def find_max(self, model, tabu_set=None):
        """find_max_indegree<limited> with and without a tabu set (operators.hpp:489-623)."""
        self._raise_uninitialized()
        if self._sorted_idx.size:
            check(lib().pbn_sort_desc(self._sorted_idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), self._sorted_idx.size,
                                      self._delta.ctypes.data_as(ctypes.POINTER(ctypes.c_double))))
        limited = self._max_indegree > 0
        n = model.num_nodes()
        for idx in self._sorted_idx.tolist():
            sc, tc = idx % n, idx // n
            source, target = model.collapsed_name(sc), model.collapsed_name(tc)
            d = self._delta[idx]
            if model.has_arc(source, target):
                op = RemoveArc(source, target, d)
                if tabu_set is None or not tabu_set.contains(op):
                    return op
            elif model.has_arc(target, source) and model.can_flip_arc(target, source):
                if limited and model.num_parents(target) >= self._max_indegree:
                    continue
                op = FlipArc(target, source, d)
                if tabu_set is None or not tabu_set.contains(op):
                    return op
            elif model.can_add_arc(source, target):
                if limited and model.num_parents(target) >= self._max_indegree:
                    continue
                op = AddArc(source, target, d)
                if tabu_set is None or not tabu_set.contains(op):
                    return op
        return None


class ChangeNodeTypeSet(OperatorSet):
    """operators.hpp:685-749, operators.cpp:439-599."""

    def __init__(self, type_blacklist=None, type_whitelist=None):
        super().__init__()
        self._delta = []
        self._is_whitelisted = np.zeros(0, dtype=bool)
        self._type_blacklist = set((n, t) for n, t in (type_blacklist or []))
        self._type_whitelist = list(type_whitelist or [])

    def set_type_blacklist(self, blacklist):
        self._type_blacklist = set((n, t) for n, t in blacklist)

    def set_type_whitelist(self, whitelist):
        self._type_whitelist = list(whitelist)

    def update_whitelisted(self, model):
        self._is_whitelisted = np.zeros(model.num_nodes(), dtype=bool)
        for name, _ in self._type_whitelist:
            self._is_whitelisted[model.collapsed_index(name)] = True

    def _cache_plans(self, model, score):
        if model.type().is_homogeneous():
            raise ValueError("ChangeNodeTypeSet can only be used with non-homogeneous Bayesian networks.")
        if not score.compatible_bn(model):
            raise ValueError("BayesianNetwork is not compatible with the score.")
        self._initialize_local_cache(model)
        plans = []
        if self._owns_local_cache:
            plans.append(self._local_cache_plan(model, model.nodes()))
        self._delta = []
        self.update_whitelisted(model)
        plan = _Plan()
        cache = self._local_cache
        bn_type = model.type()
        todo = []
        for i in range(model.num_nodes()):
            if self._is_whitelisted[i]:
                continue
            name = model.collapsed_name(i)
            t = model.node_type(name)
            if t == UnknownFactorType():
                raise ValueError("Cannot calculate ChangeNodeType delta score for " + str(t) +
                                 ". Set appropiate node types for the model")
            alt = bn_type.alternative_node_type(model, name)
            self._delta.append(np.full(len(alt), LOWEST))
            slot = len(self._delta) - 1  # the reference appends (it does not index by node): operators.cpp:474-477
            for k, at in enumerate(alt):
                if (name, at) not in self._type_blacklist and bn_type.compatible_node_type(model, name, at):
                    todo.append((slot, k, name, plan.ask(at, name, model.parents(name))))

        def fin(ans):
            for slot, k, name, a in todo:
                self._delta[slot][k] = ans[a] - cache.local_score(model, name)
        plan.finish.append(fin)
        plans.append(plan)
        return plans

    def _update_plans(self, model, score, variables):
        self._raise_uninitialized()
        plans = []
        if self._owns_local_cache:
            plans.append(self._local_cache_plan(model, variables))
        plan = _Plan()
        cache = self._local_cache
        bn_type = model.type()
        todo = []
        for n in variables:
            ci = model.collapsed_index(n)
            if self._is_whitelisted[ci]:
                continue
            alt = bn_type.alternative_node_type(model, n)
            if len(self._delta[ci]) < len(alt):
                self._delta[ci] = np.zeros(len(alt))
            if len(self._delta[ci]) > len(alt):
                self._delta[ci][len(alt):] = LOWEST
            for k, at in enumerate(alt):
                if bn_type.compatible_node_type(model, n, at) and (n, at) not in self._type_blacklist:
                    todo.append((ci, k, n, plan.ask(at, n, model.parents(n))))
                else:
                    self._delta[ci][k] = LOWEST

        def fin(ans):
            for ci, k, n, a in todo:
                self._delta[ci][k] = ans[a] - cache.local_score(model, n)
        plan.finish.append(fin)
        plans.append(plan)
        return plans

    def find_max(self, model, tabu_set=None):
        self._raise_uninitialized()
        max_score, max_node, max_type = LOWEST, -1, -1
        for i in range(len(self._delta)):
            if self._is_whitelisted[i] or len(self._delta[i]) == 0:
                continue
            if tabu_set is None:
                k = int(np.argmax(self._delta[i]))
                if self._delta[i][k] > max_score:
                    max_score, max_node, max_type = self._delta[i][k], i, k
            else:
                name = model.collapsed_name(i)
                alt = model.type().alternative_node_type(model, name)
                for k in range(len(self._delta[i])):
                    if self._delta[i][k] > max_score:
                        op = ChangeNodeType(name, alt[k], self._delta[i][k])
                        if not tabu_set.contains(op):
                            max_score, max_node, max_type = self._delta[i][k], i, k
        if max_score > LOWEST:
            name = model.collapsed_name(max_node)
            alt = model.type().alternative_node_type(model, name)
            return ChangeNodeType(name, alt[max_type], self._delta[max_node][max_type])
        return None


class OperatorPool(OperatorSet):
    """operators.hpp:751-906.  `set_type_blacklist` is deliberately NOT forwarded to the pooled sets: the
    reference's OperatorPool does not override it (operators.hpp:800-830)."""

    def __init__(self, opsets):
        super().__init__()
        self._op_sets = list(opsets)
        if not self._op_sets:
            raise ValueError("op_sets argument cannot be empty.")

    def set_arc_blacklist(self, blacklist):
        for s in self._op_sets:
            s.set_arc_blacklist(blacklist)

    def set_arc_whitelist(self, whitelist):
        for s in self._op_sets:
            s.set_arc_whitelist(whitelist)

    def set_max_indegree(self, indegree):
        for s in self._op_sets:
            s.set_max_indegree(indegree)

    def set_type_whitelist(self, whitelist):
        for s in self._op_sets:
            s.set_type_whitelist(whitelist)

    def finished(self):
        for s in self._op_sets:
            s.finished()
        super().finished()

    def _plans_of(self, op_set, kind, model, score, variables=None):
        fn = getattr(op_set, "_cache_plans" if kind == "cache" else "_update_plans", None)
        if fn is None:  # an operator set implemented outside this module: run it the reference's way
            return None
        return fn(model, score) if kind == "cache" else fn(model, score, variables)

    def cache_scores(self, model, score):
        if self._local_cache is None:
            self._initialize_local_cache(model)
            for s in self._op_sets:
                s.set_local_score_cache(self._local_cache)
        # one batched score call: the finishers run in order, so the local scores are in the cache before
        # the sets' finishers read them
        plans, foreign = [self._local_cache_plan(model, model.nodes())], []
        for s in self._op_sets:
            p = self._plans_of(s, "cache", model, score)
            if p is None:
                foreign.append(s)
            else:
                plans.extend(p)
        _run_plans(model, score, plans)
        for s in foreign:
            s.cache_scores(model, score)

    def update_scores(self, model, score, variables):
        self._raise_uninitialized()
        plans, foreign = [], []
        if self._owns_local_cache:
            plans.append(self._local_cache_plan(model, variables))
        for s in self._op_sets:
            p = self._plans_of(s, "update", model, score, variables)
            if p is None:
                foreign.append(s)
            else:
                plans.extend(p)
        _run_plans(model, score, plans)
        for s in foreign:
            s.update_scores(model, score, variables)

    def find_max(self, model, tabu_set=None):
        self._raise_uninitialized()
        if tabu_set is not None and tabu_set.empty():
            tabu_set = None
        max_delta, max_op = LOWEST, None
        for s in self._op_sets:
            op = s.find_max(model) if tabu_set is None else s.find_max(model, tabu_set)
            if op is not None and op.delta() > max_delta:
                max_op, max_delta = op, op.delta()
        return max_op
