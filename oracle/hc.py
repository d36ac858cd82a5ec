"""CPU oracle of greedy hill climbing with CVLikelihood — TEST INFRASTRUCTURE ONLY.

A plain, serial restatement of the reference's structure search for a SemiparametricBN over
continuous columns; every local score is one call of the oracle's CVLikelihood restatement
(oracle.cv_score), one (candidate, fold) after another, exactly as the reference runs it:

  learning/algorithms/hillclimbing.hpp:62-199      estimate_hc (plain Score, patience = 0 or > 0)
  learning/operators/operators.cpp:19-132          ArcOperatorSet::update_valid_ops / cache_scores
  learning/operators/operators.cpp:296-363         update_incoming_arcs_scores / update_scores
  learning/operators/operators.hpp:489-530         find_max_indegree (std::sort of the persistent index vector)
  learning/operators/operators.cpp:439-599         ChangeNodeTypeSet
  learning/operators/operators.hpp:853-869         OperatorPool::find_max
  models/SemiparametricBN.hpp:43-119               node types of continuous data: LinearGaussianCPD <-> CKDE
  graph/generic_graph.hpp:2557-2740                has_path / can_add_arc / can_flip_arc
  graph/graph_types.hpp:12-51                      parents kept in std::unordered_set<int> (iteration order!)

Nothing here is shared with pybnesian_b200: no batching, no GPU.  The only liberty is a memo of
local scores keyed on the ORDERED parent list, which SURVEY.md §7(vii) shows is result-identical.
"""
import ctypes
import sys

import numpy as np

import oracle

LOWEST = -sys.float_info.max
MACHINE_TOL = 1.4901161193847656e-08
LG, CKDE = "LinearGaussianFactor", "CKDEFactor"


class _USet:
    def __init__(self, h=None):
        L = oracle.lib()
        L.orc_uset_new.restype = ctypes.c_void_p
        L.orc_uset_clone.restype = ctypes.c_void_p
        L.orc_uset_clone.argtypes = [ctypes.c_void_p]
        for f in (L.orc_uset_free, L.orc_uset_size):
            f.argtypes = [ctypes.c_void_p]
        L.orc_uset_insert.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.orc_uset_erase.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.orc_uset_list.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int)]
        self.L = L
        self.h = ctypes.c_void_p(L.orc_uset_new()) if h is None else h

    def clone(self):
        return _USet(ctypes.c_void_p(self.L.orc_uset_clone(self.h)))

    def insert(self, v):
        self.L.orc_uset_insert(self.h, v)

    def erase(self, v):
        self.L.orc_uset_erase(self.h, v)

    def __len__(self):
        return self.L.orc_uset_size(self.h)

    def list(self):
        n = len(self)
        out = (ctypes.c_int * max(n, 1))()
        if n:
            self.L.orc_uset_list(self.h, out)
        return list(out)[:n]


class _Graph:
    def __init__(self, n):
        self.n = n
        self.pa = [_USet() for _ in range(n)]
        self.ch = [_USet() for _ in range(n)]
        self.arcs = set()

    def clone(self):
        g = _Graph.__new__(_Graph)
        g.n = self.n
        g.pa = [p.clone() for p in self.pa]
        g.ch = [c.clone() for c in self.ch]
        g.arcs = set(self.arcs)
        return g

    def add(self, s, t):
        self.arcs.add((s, t)); self.pa[t].insert(s); self.ch[s].insert(t)

    def remove(self, s, t):
        self.arcs.discard((s, t)); self.pa[t].erase(s); self.ch[s].erase(t)

    def has_path(self, s, t, skip_direct=False):
        if not skip_direct and (s, t) in self.arcs:
            return True
        seen, stack = {s}, []
        for c in self.ch[s].list():
            if skip_direct and c == t:
                continue
            stack.append(c); seen.add(c)
        while stack:
            v = stack.pop()
            kids = self.ch[v].list()
            if t in kids:
                return True
            for c in kids:
                if c not in seen:
                    seen.add(c); stack.append(c)
        return False

    def can_add(self, s, t):
        return s != t and (len(self.pa[s]) == 0 or len(self.ch[t]) == 0 or not self.has_path(t, s))

    def can_flip(self, s, t):
        if s == t:
            return False
        if (s, t) in self.arcs:
            if len(self.pa[t]) == 1 or len(self.ch[s]) == 1:
                return True
            return not self.has_path(s, t, skip_direct=True)
        if len(self.pa[t]) == 0 or len(self.ch[s]) == 0:
            return True
        return not self.has_path(s, t)


def _swap_remove(v, x):
    i = v.index(x); v[i] = v[-1]; v.pop()


def hill_climb(X, k=10, seed=0, max_indegree=0, max_iters=2 ** 31 - 1, epsilon=0.0, operators=("arcs", "node_type"),
               start_types=None, rule="normal_reference"):
    """X: (rows, nodes) float array without nulls.  Returns (operator list, final arcs, final node types,
    list of all computed deltas); operators are tuples (kind, a, b, delta) with node indices:
    ("AddArc"|"RemoveArc"|"FlipArc", source, target, delta) or ("ChangeNodeType", node, new_type, delta)."""
    X = np.asfortranarray(X)
    rows, n = X.shape
    indices, limits = oracle.cv_indices(np.arange(rows), k, seed)
    types = list(start_types) if start_types is not None else [LG] * n
    memo = {}

    def local_score(node, parents, t=None):
        t = types[node] if t is None else t
        key = (t, node, tuple(parents))
        if key not in memo:
            cols = np.asfortranarray(X[:, [node] + list(parents)])
            memo[key] = oracle.cv_score(cols, indices, limits, "ckde" if t == CKDE else "lg", rule)
        return memo[key]

    g = _Graph(n)
    prev_g, prev_types = g.clone(), list(types)
    use_arcs, use_types = "arcs" in operators, "node_type" in operators

    # OperatorPool::cache_scores: local cache, then every set
    cache = [local_score(v, g.pa[v].list()) for v in range(n)]
    delta = np.zeros(n * n)      # index = source + target * n
    valid = np.ones(n * n, dtype=bool)
    for i in range(n):
        valid[i + i * n] = False
        delta[i + i * n] = LOWEST
    sorted_idx = np.array([i + j * n for i in range(n) for j in range(n) if valid[i + j * n]], dtype=np.int32)
    all_deltas = []
    if use_arcs:
        for t in range(n):
            pt = g.pa[t].list()
            for s in range(n):
                if not valid[s + t * n]:
                    continue
                if (s, t) in g.arcs:
                    _swap_remove(pt, s)
                    d = local_score(t, pt) - cache[t]
                    pt.append(s)
                elif (t, s) in g.arcs:
                    ps = g.pa[s].list()
                    _swap_remove(ps, t)
                    pt.append(s)
                    d = local_score(s, ps) + local_score(t, pt) - cache[s] - cache[t]
                    pt.pop()
                else:
                    pt.append(s)
                    d = local_score(t, pt) - cache[t]
                    pt.pop()
                delta[s + t * n] = d
                all_deltas.append(d)
    tdelta = [LOWEST] * n
    alt = lambda v: CKDE if types[v] == LG else LG
    if use_types:
        for v in range(n):
            tdelta[v] = local_score(v, g.pa[v].list(), alt(v)) - cache[v]
            all_deltas.append(tdelta[v])

    L = oracle.lib()
    L.orc_sort_desc.argtypes = [ctypes.POINTER(ctypes.c_int), ctypes.c_int64, ctypes.POINTER(ctypes.c_double)]

    def find_max_arcs():
        L.orc_sort_desc(sorted_idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), sorted_idx.size,
                        delta.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
        for idx in sorted_idx.tolist():
            s, t = idx % n, idx // n
            if (s, t) in g.arcs:
                return ("RemoveArc", s, t, delta[idx])
            elif (t, s) in g.arcs and g.can_flip(t, s):
                if max_indegree > 0 and len(g.pa[t]) >= max_indegree:
                    continue
                return ("FlipArc", t, s, delta[idx])
            elif g.can_add(s, t):
                if max_indegree > 0 and len(g.pa[t]) >= max_indegree:
                    continue
                return ("AddArc", s, t, delta[idx])
        return None

    def find_max_types():
        best, node = LOWEST, -1
        for v in range(n):
            if tdelta[v] > best:
                best, node = tdelta[v], v
        return ("ChangeNodeType", node, alt(node), tdelta[node]) if best > LOWEST else None

    ops = []
    it = 0
    while it < max_iters:
        it += 1
        best, best_delta = None, LOWEST
        for finder, on in ((find_max_arcs, use_arcs), (find_max_types, use_types)):
            if on:
                op = finder()
                if op is not None and op[3] > best_delta:
                    best, best_delta = op, op[3]
        if best is None or (best[3] - epsilon) < MACHINE_TOL:
            break
        kind, a, b, d = best
        if kind == "AddArc":
            g.add(a, b); changed = [b]
        elif kind == "RemoveArc":
            g.remove(a, b); changed = [b]
        elif kind == "FlipArc":
            g.remove(a, b); g.add(b, a); changed = [a, b]
        else:
            types[a] = b; changed = [a]
        if not (d > MACHINE_TOL):  # plain Score: validation delta = operator delta (zero patience)
            g, types = prev_g, prev_types
            break
        if kind == "AddArc":
            prev_g.add(a, b)
        elif kind == "RemoveArc":
            prev_g.remove(a, b)
        elif kind == "FlipArc":
            prev_g.remove(a, b); prev_g.add(b, a)
        else:
            prev_types[a] = b
        ops.append(best)
        # OperatorPool::update_scores
        for v in changed:
            cache[v] = local_score(v, g.pa[v].list())
        if use_arcs:
            for t in changed:
                parents = g.pa[t].list()
                for s in range(n):
                    if not valid[s + t * n]:
                        continue
                    if (s, t) in g.arcs:
                        _swap_remove(parents, s)
                        dd = local_score(t, parents) - cache[t]
                        parents.append(s)
                        delta[s + t * n] = dd
                        if valid[t + s * n]:
                            ps = g.pa[s].list()
                            ps.append(t)
                            delta[t + s * n] = dd + local_score(s, ps) - cache[s]
                            all_deltas.append(delta[t + s * n])
                    elif (t, s) in g.arcs:
                        ps = g.pa[s].list()
                        _swap_remove(ps, t)
                        parents.append(s)
                        dd = local_score(s, ps) + local_score(t, parents) - cache[s] - cache[t]
                        parents.pop()
                        delta[s + t * n] = dd
                    else:
                        parents.append(s)
                        dd = local_score(t, parents) - cache[t]
                        parents.pop()
                        delta[s + t * n] = dd
                    all_deltas.append(dd)
        if use_types:
            for v in changed:
                tdelta[v] = local_score(v, g.pa[v].list(), alt(v)) - cache[v]
                all_deltas.append(tdelta[v])
    return ops, sorted(g.arcs), types, all_deltas
