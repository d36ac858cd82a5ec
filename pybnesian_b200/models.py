"""Graph and Bayesian-network containers used by the score / hill-climbing path.

Mirrors the slice of the reference that `estimate_hc` touches:
graph/generic_graph.hpp (Dag: arcs, parents, can_add_arc / can_flip_arc / has_path),
graph/graph_types.hpp:12-51 (DNode: parents and children are `std::unordered_set<int>`, whose
iteration order fixes the order in which `parents(node)` lists the evidence of a CPD),
models/BayesianNetwork.hpp (node types, clone, whitelists), models/SemiparametricBN.hpp:17-169,
models/GaussianNetwork.hpp, models/KDENetwork.hpp.  The parent / children sets are REAL
libstdc++ unordered_sets held behind the C ABI (pbn_intset_*), so the evidence order is the
reference's, not an emulation of it.
"""
import ctypes
import pickle

import numpy as np
import pyarrow as pa

from ._lib import check, lib
from .dataset import DataFrame
from .factors import (CKDEType, FactorType, LinearGaussianCPDType, UnknownFactorType)
from .hybrid import DiscreteFactorType


class _IntSet:
    """std::unordered_set<int> (pbn_intset)."""

    __slots__ = ("h",)

    def __init__(self, handle=None):
        if handle is None:
            handle = ctypes.c_void_p()
            check(lib().pbn_intset_new(ctypes.byref(handle)))
        self.h = handle

    def clone(self):
        h = ctypes.c_void_p()
        check(lib().pbn_intset_clone(self.h, ctypes.byref(h)))
        return _IntSet(h)

    def insert(self, v):
        lib().pbn_intset_insert(self.h, int(v))

    def erase(self, v):
        lib().pbn_intset_erase(self.h, int(v))

    def __contains__(self, v):
        return bool(lib().pbn_intset_contains(self.h, int(v)))

    def __len__(self):
        return lib().pbn_intset_size(self.h)

    def list(self):
        n = len(self)
        if n == 0:
            return []
        out = (ctypes.c_int * n)()
        check(lib().pbn_intset_list(self.h, out))
        return list(out)

    def __del__(self):
        try:
            if self.h:
                lib().pbn_intset_free(self.h)
                self.h = None
        except Exception:
            pass


class Dag:
    """graph::Dag over a fixed node list (graph/generic_graph.hpp:1185-1256, 2557-2740)."""

    def __init__(self, nodes=None, arcs=None):
        nodes = list(nodes or [])
        for a in (arcs or []):
            if not (isinstance(a, (tuple, list)) and len(a) == 2 and all(isinstance(x, str) for x in a)):
                # what pybind11 answers when no overload takes the arguments (BayesianNetwork_test.py:28-30)
                raise TypeError("__init__(): incompatible constructor arguments. An arc is a (source, target) pair of node names.")
        if arcs and not nodes:
            for s, t in arcs:
                for n in (s, t):
                    if n not in nodes:
                        nodes.append(n)
        if len(set(nodes)) != len(nodes):
            raise ValueError("Graph cannot be created with repeated names.")
        # raw slots (None once removed) / name -> raw index / names in collapsed order / name -> collapsed index / free
        # raw slots: GraphBase::m_nodes, m_indices, m_string_nodes (a BidirectionalMapIndex), m_free_indices
        # (graph/generic_graph.hpp:395-507)
        self._names = nodes
        self._index = {n: i for i, n in enumerate(nodes)}
        self._collapsed = list(nodes)
        self._cindex = {n: i for i, n in enumerate(nodes)}
        self._free = []
        self._parents = [_IntSet() for _ in nodes]
        self._children = [_IntSet() for _ in nodes]
        self._arcs = set()
        # ArcGraph::m_roots (generic_graph.hpp:979-993, 1185-1248): an unordered_set<int> whose iteration order
        # (insertion / erase history) seeds the stack of topological_sort, hence the per-node seeds of sample()
        self._roots = _IntSet()
        for i in range(len(nodes)):
            self._roots.insert(i)
        # Dag(nodes, arcs) (graph/generic_graph.hpp: DagImpl constructor): arcs go in unchecked, then one topological sort
        # validates the whole graph ("Graph must be a DAG to obtain a topological sort.")
        for s, t in (arcs or []):
            si, ti = self.index(s), self.index(t)
            if (si, ti) not in self._arcs:
                self._add_arc_unsafe(si, ti)
        if arcs:
            self.topological_sort()

    # -- nodes ---------------------------------------------------------------------------
    def nodes(self):
        return list(self._collapsed)

    def num_nodes(self):
        return len(self._collapsed)

    def num_raw_nodes(self):
        return len(self._names)

    def contains_node(self, name):
        return name in self._index

    def index(self, name):
        try:
            return self._index[name]
        except KeyError:
            raise IndexError("Node " + str(name) + " not present in the graph.")

    def collapsed_index(self, name):
        try:
            return self._cindex[name]
        except KeyError:
            raise IndexError("Node " + str(name) + " not present in the graph.")

    def name(self, idx):
        if idx < 0 or idx >= len(self._names) or self._names[idx] is None:
            raise IndexError("Node index " + str(idx) + " not present in the graph.")
        return self._names[idx]

    def collapsed_name(self, idx):
        return self._collapsed[idx]

    def collapsed_from_index(self, idx):
        return self._cindex[self.name(idx)]

    def index_from_collapsed(self, cidx):
        return self._index[self._collapsed[cidx]]

    def add_node(self, name):
        """GraphBase::add_node / create_node (generic_graph.hpp:509-544): a freed raw slot is reused, last freed first."""
        if name in self._index:
            raise ValueError("Cannot add node " + str(name) + " because a node with the same name already exists.")
        if self._free:
            idx = self._free.pop()
            self._names[idx] = name
            self._parents[idx], self._children[idx] = _IntSet(), _IntSet()
        else:
            idx = len(self._names)
            self._names.append(name)
            self._parents.append(_IntSet())
            self._children.append(_IntSet())
        self._index[name] = idx
        self._cindex[name] = len(self._collapsed)
        self._collapsed.append(name)
        self._roots.insert(idx)
        return idx

    def remove_node(self, name):
        """GraphBase::remove_node_unsafe (generic_graph.hpp:546-580): arcs of the node go first, the raw slot is freed, and
        the collapsed order closes the gap with the LAST node (BidirectionalMapIndex::remove = swap_remove)."""
        idx = self.index(name)
        if idx in self._roots:
            self._roots.erase(idx)
        for p in self._parents[idx].list():
            self._remove_arc_unsafe(p, idx)
        for ch in self._children[idx].list():
            self._remove_arc_unsafe(idx, ch)
        if idx in self._roots:  # _remove_arc_unsafe re-inserts a node that lost its last parent
            self._roots.erase(idx)
        ci = self._cindex.pop(name)
        last = self._collapsed.pop()
        if ci < len(self._collapsed):
            self._collapsed[ci] = last
            self._cindex[last] = ci
        del self._index[name]
        self._names[idx] = None
        self._free.append(idx)

    # -- arcs ----------------------------------------------------------------------------
    def num_arcs(self):
        return len(self._arcs)

    def arcs(self):
        return [(self._names[s], self._names[t]) for s, t in sorted(self._arcs)]

    def has_arc(self, source, target):
        return (self.index(source), self.index(target)) in self._arcs

    def parents(self, node):
        return [self._names[p] for p in self._parents[self.index(node)].list()]

    def parent_indices(self, node):
        return self._parents[self.index(node)].list()

    def children(self, node):
        return [self._names[c] for c in self._children[self.index(node)].list()]

    def num_parents(self, node):
        return len(self._parents[self.index(node)])

    def num_children(self, node):
        return len(self._children[self.index(node)])

    def _add_arc_unsafe(self, s, t):
        if t in self._roots:
            self._roots.erase(t)
        self._arcs.add((s, t))
        self._parents[t].insert(s)
        self._children[s].insert(t)

    def _remove_arc_unsafe(self, s, t):
        self._arcs.discard((s, t))
        self._parents[t].erase(s)
        self._children[s].erase(t)
        if len(self._parents[t]) == 0:
            self._roots.insert(t)

    def add_arc_unsafe(self, source, target):
        self._add_arc_unsafe(self.index(source), self.index(target))

    def add_arc(self, source, target):
        s, t = self.index(source), self.index(target)
        if (s, t) in self._arcs:
            return
        if not self._can_add(s, t):
            raise ValueError("Cannot add arc " + source + " -> " + target + ".")
        self._add_arc_unsafe(s, t)

    def remove_arc(self, source, target):
        s, t = self.index(source), self.index(target)
        if (s, t) in self._arcs:
            self._remove_arc_unsafe(s, t)

    def flip_arc_unsafe(self, source, target):
        s, t = self.index(source), self.index(target)
        self._remove_arc_unsafe(s, t)
        self._add_arc_unsafe(t, s)

    def flip_arc(self, source, target):
        s, t = self.index(source), self.index(target)
        if not self._can_flip(s, t):
            raise ValueError("Cannot flip arc " + source + " -> " + target + ".")
        if (s, t) in self._arcs:
            self._remove_arc_unsafe(s, t)
            self._add_arc_unsafe(t, s)

    # -- acyclicity (DirectedImpl::has_path_unsafe*, DagImpl::can_add/flip_arc_unsafe) --------------
    def _has_path(self, s, t, skip_direct=False):
        if not skip_direct and (s, t) in self._arcs:
            return True
        seen = {s}
        stack = []
        for ch in self._children[s].list():
            if skip_direct and ch == t:
                continue
            stack.append(ch)
            seen.add(ch)
        while stack:
            v = stack.pop()
            ch = self._children[v]
            if t in ch:
                return True
            for c in ch.list():
                if c not in seen:
                    seen.add(c)
                    stack.append(c)
        return False

    def has_path(self, source, target):
        return self._has_path(self.index(source), self.index(target))

    def _can_add(self, s, t):
        return s != t and (len(self._parents[s]) == 0 or len(self._children[t]) == 0 or not self._has_path(t, s))

    def _can_flip(self, s, t):
        if s == t:
            return False
        if (s, t) in self._arcs:
            if len(self._parents[t]) == 1 or len(self._children[s]) == 1:
                return True
            return not self._has_path(s, t, skip_direct=True)
        if len(self._parents[t]) == 0 or len(self._children[s]) == 0:
            return True
        return not self._has_path(s, t)

    def can_add_arc(self, source, target):
        return self._can_add(self.index(source), self.index(target))

    def can_flip_arc(self, source, target):
        return self._can_flip(self.index(source), self.index(target))

    def roots(self):
        """ArcGraph::roots (graph/generic_graph.hpp:1133): names of the nodes without parents."""
        return {self._names[i] for i in self._roots.list()}

    def leaves(self):
        return {n for i, n in enumerate(self._names) if n is not None and len(self._children[i]) == 0}

    def save(self, filename):
        if not filename.endswith(".pickle"):
            filename += ".pickle"
        with open(filename, "wb") as f:
            pickle.dump(self, f, protocol=2)

    def topological_sort(self):
        indeg = [len(p) for p in self._parents]
        stack = self._roots.list()  # DagImpl::topological_sort (generic_graph.hpp:2659-2708)
        order = []
        while stack:
            i = stack.pop()
            order.append(self._names[i])
            for ch in self._children[i].list():
                indeg[ch] -= 1
                if indeg[ch] == 0:
                    stack.append(ch)
        if len(order) != len(self._collapsed):
            raise ValueError("Graph must be a DAG to obtain a topological sort.")
        return order

    def clone(self):
        g = Dag.__new__(Dag)
        g._names = list(self._names)
        g._index = dict(self._index)
        g._collapsed = list(self._collapsed)
        g._cindex = dict(self._cindex)
        g._free = list(self._free)
        g._parents = [p.clone() for p in self._parents]  # unordered_set copy keeps the iteration order
        g._children = [c.clone() for c in self._children]
        g._arcs = set(self._arcs)
        g._roots = self._roots.clone()
        return g

    def __getstate__(self):
        # graph::__getstate__ (generic_graph.hpp:282-330) saves the collapsed node list: removed slots do not survive
        return (self.nodes(), self.arcs())

    def __setstate__(self, t):
        self.__init__(t[0], t[1])


# ----------------------------------------------------------------------------------------------
# Bayesian network types (models/BayesianNetwork.hpp:170-260 BayesianNetworkType)
# ----------------------------------------------------------------------------------------------
class BayesianNetworkType:
    _instances = {}

    def __new__(cls, *a, **k):
        inst = BayesianNetworkType._instances.get(cls)
        if inst is None:
            inst = super().__new__(cls)
            BayesianNetworkType._instances[cls] = inst
        return inst

    def is_homogeneous(self):
        raise NotImplementedError

    def default_node_type(self):
        raise NotImplementedError

    def data_default_node_type(self, datatype):
        return []

    def compatible_node_type(self, model, variable, node_type):
        return True

    def can_have_arc(self, model, source, target):
        return True

    def alternative_node_type(self, model, variable):
        return []

    def new_bn(self, nodes):
        raise NotImplementedError

    def __eq__(self, other):
        return type(self) is type(other)

    def __hash__(self):
        return hash(type(self).__name__)

    def __str__(self):
        return type(self).__name__

    __repr__ = __str__

    def __reduce__(self):
        return (type(self), ())


class GaussianNetworkType(BayesianNetworkType):
    """models/GaussianNetwork.hpp."""

    def is_homogeneous(self):
        return True

    def default_node_type(self):
        return LinearGaussianCPDType()

    def new_bn(self, nodes):
        return GaussianNetwork(nodes)

    def __str__(self):
        return "GaussianNetworkType"

    __repr__ = __str__


class KDENetworkType(BayesianNetworkType):
    """models/KDENetwork.hpp."""

    def is_homogeneous(self):
        return True

    def default_node_type(self):
        return CKDEType()

    def new_bn(self, nodes):
        return KDENetwork(nodes)

    def __str__(self):
        return "KDENetworkType"

    __repr__ = __str__


def _is_continuous(datatype):
    return datatype in (pa.float64(), pa.float32())


class SemiparametricBNType(BayesianNetworkType):
    """models/SemiparametricBN.hpp:17-136."""

    def is_homogeneous(self):
        return False

    def default_node_type(self):
        raise RuntimeError("default_node_type() for SemiparametricBN is not defined.")

    def data_default_node_type(self, datatype):
        if _is_continuous(datatype):
            return [LinearGaussianCPDType(), CKDEType()]
        if pa.types.is_dictionary(datatype):
            return [DiscreteFactorType()]
        raise ValueError("Data type [" + str(datatype) + "] not compatible with SemiparametricBNType")

    def compatible_node_type(self, model, variable, node_type):
        # SemiparametricBN.hpp:62-78: a discrete node may only have discrete parents
        if node_type not in (LinearGaussianCPDType(), CKDEType(), DiscreteFactorType()):
            return False
        if node_type == DiscreteFactorType():
            return all(model.node_type(p) == DiscreteFactorType() for p in model.parents(variable))
        return True

    def can_have_arc(self, model, source, target):
        # SemiparametricBN.hpp:96-101
        return model.node_type(target) != DiscreteFactorType() or model.node_type(source) == DiscreteFactorType()

    def alternative_node_type(self, model, variable):
        t = model.node_type(variable)
        if t == LinearGaussianCPDType():
            return [CKDEType()]
        if t == CKDEType():
            return [LinearGaussianCPDType()]
        return []

    def new_bn(self, nodes):
        return SemiparametricBN(nodes)

    def __str__(self):
        return "SemiparametricNetworkType"

    __repr__ = __str__


class HeterogeneousBNType(BayesianNetworkType):
    """models/HeterogeneousBN.hpp:28-196: the default factor types of a node are given by the user, either one list for
    every data type or one list per Arrow data type (first entry = default, compared by type id as DataTypeEqualTo)."""

    def __new__(cls, *a, **k):  # parametrised: not a per-class singleton
        return object.__new__(cls)

    def __init__(self, default_factor_types):
        if isinstance(default_factor_types, dict):
            self._single = False
            self._ftype = []
            self._ftypes = {dt: list(fts) for dt, fts in default_factor_types.items() if list(fts)}
            if not self._ftypes:
                raise ValueError("Default factor_type cannot be empty.")
            for dt, fts in self._ftypes.items():
                if dt is None:
                    raise ValueError("Default factor_types cannot contain null DataType.")
                if any(f is None for f in fts):
                    raise ValueError("Default factor_type cannot contain null FactorType.")
        else:
            self._single = True
            self._ftype = list(default_factor_types)
            self._ftypes = {}
            if not self._ftype:
                raise ValueError("Default factor_type cannot be empty.")
            if any(f is None for f in self._ftype):
                raise ValueError("Default factor_type cannot contain null FactorType.")

    def is_homogeneous(self):
        return False

    def default_node_type(self):
        raise RuntimeError("default_node_type() for HeterogeneousBN is not defined.")

    def data_default_node_type(self, datatype):
        if self._single:
            return list(self._ftype)
        for dt, fts in self._ftypes.items():
            if dt.id == datatype.id:
                return list(fts)
        raise ValueError("Not valid FactorType for DataType " + str(datatype))

    def single_default(self):
        return self._single

    def default_node_types(self):
        return {None: list(self._ftype)} if self._single else {dt: list(f) for dt, f in self._ftypes.items()}

    def new_bn(self, nodes):
        return HeterogeneousBN(self._ftype if self._single else self._ftypes, nodes)

    def _key(self):
        if self._single:
            return ("single", tuple(str(f) for f in self._ftype))
        return ("multi", frozenset((dt.id, tuple(str(f) for f in fts)) for dt, fts in self._ftypes.items()))

    def __eq__(self, other):
        return isinstance(other, HeterogeneousBNType) and self._key() == other._key()

    def __hash__(self):
        return hash(self._key())

    def __str__(self):
        if self._single:
            return "HeterogeneousBNType([" + ", ".join(str(f) for f in self._ftype) + "])"
        return "HeterogeneousBNType({" + ", ".join(str(dt) + ": [" + ", ".join(str(f) for f in fts) + "]"
                                                    for dt, fts in self._ftypes.items()) + "})"

    __repr__ = __str__

    def __reduce__(self):
        return (HeterogeneousBNType, (self._ftype if self._single else self._ftypes,))


# ----------------------------------------------------------------------------------------------
# BayesianNetwork (models/BayesianNetwork.hpp:262-1000, unconditional networks only)
# ----------------------------------------------------------------------------------------------
class BayesianNetwork:
    def __init__(self, bn_type, nodes=None, arcs=None, node_types=None, graph=None):
        self._type = bn_type
        if isinstance(nodes, Dag):
            graph, nodes = nodes, None
        if graph is not None:
            self._g = graph.clone()
        else:
            nodes = list(nodes) if nodes is not None else None
            # a list of 2-tuples in the first position is an arc list (BNGeneric(type, arcs))
            if nodes and arcs is None and all(isinstance(n, tuple) and len(n) == 2 for n in nodes) \
                    and not all(isinstance(n[1], FactorType) for n in nodes):
                arcs, nodes = nodes, None
            self._g = Dag(nodes, arcs)
        n = self._g.num_nodes()
        self._cpds = []
        if bn_type.is_homogeneous():
            self._node_types = []
            for name, t in (node_types or []):  # a homogeneous network only takes its own default type
                if t != bn_type.default_node_type():
                    raise self._wrong_type(name, t)
        else:
            self._node_types = [UnknownFactorType() for _ in range(n)]
            for name, t in (node_types or []):
                self.set_node_type(name, t)

    # -- graph delegation ----------------------------------------------------------------------
    def graph(self):
        return self._g

    def type(self):
        return self._type

    def nodes(self):
        return self._g.nodes()

    def num_nodes(self):
        return self._g.num_nodes()

    def num_arcs(self):
        return self._g.num_arcs()

    def arcs(self):
        return self._g.arcs()

    def contains_node(self, name):
        return self._g.contains_node(name)

    def index(self, node):
        return self._g.index(node)

    def collapsed_index(self, node):
        return self._g.collapsed_index(node)

    def collapsed_name(self, idx):
        return self._g.collapsed_name(idx)

    def collapsed_from_index(self, idx):
        return self._g.collapsed_from_index(idx)

    def indices(self):
        """name -> raw index of every node (GraphBase::indices, graph/generic_graph.hpp:431)."""
        return dict(self._g._index)

    def collapsed_indices(self):
        return dict(self._g._cindex)

    def index_from_collapsed(self, cidx):
        return self._g.index_from_collapsed(cidx)

    def add_node(self, node):
        """BNGeneric::add_node (models/BayesianNetwork.hpp:503-515)."""
        idx = self._g.add_node(node)
        if idx == self._g.num_raw_nodes() - 1:
            if self._cpds:
                self._cpds.append(None)
            if not self._type.is_homogeneous():
                self._node_types.append(UnknownFactorType())
        return idx

    def remove_node(self, node):
        """BNGeneric::remove_node (models/BayesianNetwork.hpp:517-527)."""
        idx = self._g.index(node)
        if self._cpds:
            self._cpds[idx] = None
        if not self._type.is_homogeneous():
            self._node_types[idx] = UnknownFactorType()
        self._g.remove_node(node)

    def name(self, idx):
        return self._g.name(idx)

    def parents(self, node):
        return self._g.parents(node)

    def children(self, node):
        return self._g.children(node)

    def num_parents(self, node):
        return self._g.num_parents(node)

    def num_children(self, node):
        return self._g.num_children(node)

    def has_arc(self, source, target):
        return self._g.has_arc(source, target)

    def has_path(self, source, target):
        return self._g.has_path(source, target)

    def can_add_arc(self, source, target):
        return self._g.can_add_arc(source, target) and self._type.can_have_arc(self, source, target)

    def can_flip_arc(self, source, target):
        return self._g.can_flip_arc(source, target) and self._type.can_have_arc(self, target, source)

    def add_arc(self, source, target):
        if self.can_add_arc(source, target):
            self._g.add_arc_unsafe(source, target)
        elif not self.has_arc(source, target):
            raise ValueError("Cannot add arc " + source + " -> " + target + ".")

    def add_arc_unsafe(self, source, target):
        self._g.add_arc_unsafe(source, target)

    def remove_arc(self, source, target):
        self._g.remove_arc(source, target)

    def flip_arc(self, source, target):
        if self.can_flip_arc(source, target):
            self._g.flip_arc_unsafe(source, target)
        else:
            raise ValueError("Cannot flip arc " + source + " -> " + target + ".")

    def flip_arc_unsafe(self, source, target):
        self._g.flip_arc_unsafe(source, target)

    def check_blacklist(self, arc_blacklist):
        for s, t in arc_blacklist:
            if self.has_arc(s, t):
                raise ValueError("Arc " + s + " -> " + t + " in blacklist, but it is present in the Bayesian Network.")

    def force_whitelist(self, arc_whitelist):
        for s, t in arc_whitelist:
            if not self.has_arc(s, t):
                if self.has_arc(t, s):
                    raise ValueError("Arc " + s + " -> " + t + " in whitelist, but arc " + t + " -> " + s +
                                     " is present in the Bayesian Network.")
                elif self.can_add_arc(s, t):
                    self.add_arc_unsafe(s, t)
                else:
                    raise ValueError("Arc " + s + " -> " + t + " not allowed in this Bayesian network.")
        self._g.topological_sort()

    # -- node types (BayesianNetwork.hpp:640-800) ------------------------------------------------------
    def node_type(self, node):
        if self._type.is_homogeneous():
            return self._type.default_node_type()
        return self._node_types[self.index(node)]

    def node_types(self):
        return {n: self.node_type(n) for n in self.nodes()}

    def underlying_node_type(self, df, node):
        if self._type.is_homogeneous():
            return self._type.default_node_type()
        t = self._node_types[self.index(node)]
        if t != UnknownFactorType():
            return t
        frame = DataFrame.wrap(df)
        nt = self._type.data_default_node_type(frame._col(node).type)
        if not nt:
            raise ValueError("There is no underlying FactorType for node " + node)
        return nt[0]

    def _wrong_type(self, node, t):
        return ValueError("Wrong factor type \"" + str(t) + "\" for node \"" + node + "\" in Bayesian network type \"" +
                          str(self._type) + "\".")

    def set_node_type(self, node, new_type):
        if self._type.is_homogeneous():
            if new_type != self._type.default_node_type():
                raise self._wrong_type(node, new_type)
            return
        if new_type != UnknownFactorType() and not self._type.compatible_node_type(self, node, new_type):
            raise self._wrong_type(node, new_type)
        i = self.index(node)
        self._node_types[i] = new_type
        if self._cpds and self._cpds[i] is not None and self._cpds[i].type() != new_type:
            self._cpds[i] = None

    def has_unknown_node_types(self):
        if self._type.is_homogeneous():
            return False
        return any(self._node_types[self.index(n)] == UnknownFactorType() for n in self.nodes())

    def set_unknown_node_types(self, df, type_blacklist=()):
        if self._type.is_homogeneous():
            return
        frame = DataFrame.wrap(df)
        black = set((n, t) for n, t in type_blacklist)
        new_types = []
        for nn in self.nodes():
            if self.node_type(nn) == UnknownFactorType():
                for t in self._type.data_default_node_type(frame._col(nn).type):
                    if (nn, t) not in black:
                        new_types.append((nn, t))
                        break
                else:
                    raise ValueError("A valid FactorType for node " + nn + " could not be inferred.")
        self.force_type_whitelist(new_types)

    def force_type_whitelist(self, type_whitelist):
        if self._type.is_homogeneous():
            for n, t in type_whitelist:
                if t != self._type.default_node_type():
                    raise self._wrong_type(n, t)
            return
        old = []
        for n, t in type_whitelist:
            i = self.index(n)
            old.append((i, self._node_types[i]))
            self._node_types[i] = t
        for n, t in type_whitelist:
            if t != UnknownFactorType() and not self._type.compatible_node_type(self, n, t):
                for i, o in old:
                    self._node_types[i] = o
                raise self._wrong_type(n, t)
        if self._cpds:
            for n, _ in type_whitelist:
                i = self.index(n)
                if self._cpds[i] is not None and self._cpds[i].type() != self._node_types[i]:
                    self._cpds[i] = None

    # -- CPDs --------------------------------------------------------------------------------
    def fitted(self):
        if not self._cpds:
            return False
        cpds = [self._cpds[self.index(n)] for n in self.nodes()]
        return all(c is not None and c.fitted() for c in cpds)

    def cpd(self, node):
        i = self.index(node)
        if not self._cpds or self._cpds[i] is None:
            raise ValueError("CPD of variable \"" + node + "\" not added. Call add_cpds() or fit() to add the CPD.")
        return self._cpds[i]

    def fit(self, df, construction_args=None):
        frame = DataFrame.wrap(df)
        if not self._cpds:
            self._cpds = [None] * self._g.num_raw_nodes()
        # BayesianNetwork.hpp:965-976: nodes without a type take their data's default type first, so that
        # new_factor sees the (possibly discrete) types of the parents
        if not self._type.is_homogeneous():
            self.force_type_whitelist([(n, self.underlying_node_type(frame, n)) for n in self.nodes()
                                       if self.node_type(n) == UnknownFactorType()])
        for node in self.nodes():
            i = self.index(node)
            t = self.node_type(node)
            parents = self.parents(node)
            cur = self._cpds[i]
            if cur is None or cur.type() != t or cur.evidence() != parents:
                args, kwargs = construction_args.args(node, t) if construction_args is not None else ((), {})
                cur = t.new_factor(self, node, parents, *args, **kwargs)
                self._cpds[i] = cur
            if not cur.fitted():
                cur.fit(frame)

    def logl(self, df):
        frame = DataFrame.wrap(df)
        total = np.zeros(frame.num_rows)
        for node in self.nodes():
            total = total + self.cpd(node).logl(frame)
        return total

    def slogl(self, df):
        frame = DataFrame.wrap(df)
        total = 0.0  # plain additions in node order like the reference (the built-in sum() compensates since Python 3.12)
        for node in self.nodes():
            total += self.cpd(node).slogl(frame)
        return float(total)

    def sample(self, n, seed=None, ordered=False):
        """BNGeneric::sample (models/BayesianNetwork.hpp:1024-1065): ancestral sampling in topological order; node i of
        the order is sampled with seed + i given the columns sampled so far.  Returns a pandas DataFrame (columns in
        topological order, or in nodes() order when `ordered`)."""
        import pyarrow as pa
        from .factors import _random_seed
        if n < 0:
            raise ValueError("n should be a non-negative number")
        if not self.fitted():
            raise ValueError("Model not fitted.")
        seed = _random_seed(seed)
        names, arrays = [], []
        for i, node in enumerate(self._g.topological_sort()):
            parents = DataFrame(pa.RecordBatch.from_arrays(arrays, names=names)) if arrays else None
            arr = self.cpd(node).sample(n, parents, (seed + i) & 0xFFFFFFFF)
            names.append(node)
            arrays.append(arr)
        if ordered:
            order = [names.index(v) for v in self.nodes()]
            names, arrays = [names[j] for j in order], [arrays[j] for j in order]
        # a pyarrow.RecordBatch, what the reference's DataFrame type caster returns (dataset/dataset.hpp:2120-2143)
        return pa.RecordBatch.from_arrays(arrays, names=names)

    def clone(self):
        m = type(self).__new__(type(self))
        m._type = self._type
        m._g = self._g.clone()
        m._node_types = list(self._node_types)
        m._cpds = list(self._cpds)
        m._include_cpd = self.include_cpd()
        # a Python-derived network keeps its own state across clone(), as the reference's trampoline does through
        # __getstate_extra__ / __setstate_extra__ (models/BayesianNetwork.hpp:1139-1167, pybindings_models.cpp)
        for key, val in self.__dict__.items():
            if key not in m.__dict__:
                m.__dict__[key] = val
        if hasattr(self, "__getstate_extra__") and hasattr(m, "__setstate_extra__"):
            m.__setstate_extra__(self.__getstate_extra__())
        return m

    def check_compatible_cpd(self, cpd):
        """BNGeneric::check_compatible_cpd (models/BayesianNetwork.hpp:863-912)."""
        if not self.contains_node(cpd.variable()):
            raise ValueError("CPD defined on variable which is not present in the model:\n" + str(cpd))
        evidence = cpd.evidence()
        for ev in evidence:
            if not self.contains_node(ev):
                raise ValueError("Evidence variable " + ev + " is not present in the model:\n" + str(cpd))
        pa = self.parents(cpd.variable())
        if len(pa) != len(evidence) or set(pa) != set(evidence):
            raise ValueError("CPD do not have the model's parent set as evidence:\n" + str(cpd) + "\nParents: ["
                             + ", ".join(pa) + "]")
        t = self.node_type(cpd.variable())
        if t != UnknownFactorType() and cpd.type() != t:
            raise ValueError("Factor " + str(cpd) + " is of type " + str(cpd.type()) + ". Bayesian network expects type "
                             + str(t))

    def add_cpds(self, cpds):
        """BNGeneric::add_cpds (models/BayesianNetwork.hpp:914-940)."""
        cpds = list(cpds)
        for cpd in cpds:
            self.check_compatible_cpd(cpd)
        if not self._type.is_homogeneous():
            self.force_type_whitelist([(c.variable(), c.type()) for c in cpds
                                       if self.node_type(c.variable()) == UnknownFactorType()])
        if not self._cpds:
            self._cpds = [None] * self._g.num_raw_nodes()
        for cpd in cpds:
            self._cpds[self.index(cpd.variable())] = cpd

    def include_cpd(self):
        return getattr(self, "_include_cpd", False)

    def set_include_cpd(self, include_cpd):
        self._include_cpd = bool(include_cpd)

    def save(self, filename, include_cpd=False):
        """BNGeneric::save (models/BayesianNetwork.hpp:1127-1137, util/pickle.hpp): pickle into `filename`.pickle;
        the CPDs travel only when include_cpd (GPU-resident factors are read back through their own __getstate__)."""
        self._include_cpd = bool(include_cpd)
        if not filename.endswith(".pickle"):
            filename += ".pickle"
        with open(filename, "wb") as f:
            pickle.dump(self, f, protocol=2)

    # BNGeneric::__getstate__ (models/BayesianNetwork.hpp:1139-1167): (graph, type, node types, include_cpd, cpds)
    def __getstate__(self):
        cpds = []
        if self.include_cpd() and self._cpds:
            for node in self.nodes():
                c = self._cpds[self.index(node)]
                if c is not None:
                    try:
                        self.check_compatible_cpd(c)
                        cpds.append(c)
                    except ValueError:
                        pass
        node_types = []
        if not self._type.is_homogeneous():
            node_types = [(n, self._node_types[self.index(n)]) for n in self.nodes()
                          if self._node_types[self.index(n)] != UnknownFactorType()]
        return (self._type, self._g.nodes(), self._g.arcs(), node_types, self.include_cpd(), cpds)

    def __setstate__(self, t):
        self._type = t[0]
        self._g = Dag(t[1], t[2])
        self._cpds = []
        if len(t) == 4:  # pickles written before the CPDs were part of the state
            self._node_types = list(t[3])
            return
        self._node_types = [] if self._type.is_homogeneous() else [UnknownFactorType() for _ in t[1]]
        for name, ft in t[3]:
            self._node_types[self.index(name)] = ft
        self._include_cpd = bool(t[4])
        if t[4] and t[5]:
            self.add_cpds(t[5])

    def __str__(self):
        return type(self).__name__ + " with %d nodes and %d arcs" % (self.num_nodes(), self.num_arcs())

    __repr__ = __str__


class GaussianNetwork(BayesianNetwork):
    def __init__(self, nodes=None, arcs=None, graph=None):
        super().__init__(GaussianNetworkType(), nodes, arcs, None, graph)


class KDENetwork(BayesianNetwork):
    def __init__(self, nodes=None, arcs=None, graph=None):
        super().__init__(KDENetworkType(), nodes, arcs, None, graph)


class SemiparametricBN(BayesianNetwork):
    """pybnesian.SemiparametricBN (models/SemiparametricBN.hpp:138-169): ctor forms (nodes), (arcs),
    (nodes, arcs), (graph), each optionally followed by node_types = [(name, FactorType), ...]."""

    def __init__(self, nodes=None, arcs=None, node_types=None, graph=None):
        # (nodes, node_types) and (arcs, node_types): the second positional is a FactorType list
        if arcs is not None and node_types is None and len(arcs) > 0 and \
                all(isinstance(a, tuple) and len(a) == 2 and isinstance(a[1], FactorType) for a in arcs):
            arcs, node_types = None, arcs
        super().__init__(SemiparametricBNType(), nodes, arcs, node_types, graph)


def load(filename):
    """pybnesian.load (util/pickle.hpp, lib.cpp:38-43): the object saved by any `.save()` of this package."""
    with open(filename, "rb") as f:
        return pickle.load(f)


class HeterogeneousBN(BayesianNetwork):
    """pybnesian.HeterogeneousBN (models/HeterogeneousBN.hpp:198-290): HeterogeneousBN(factor_types, nodes | arcs | graph
    [, node_types]) with factor_types a list of FactorType or {pyarrow.DataType: [FactorType]}."""

    def __init__(self, factor_types, nodes=None, arcs=None, node_types=None, graph=None):
        bn_type = factor_types if isinstance(factor_types, HeterogeneousBNType) else HeterogeneousBNType(factor_types)
        # (nodes, node_types) and (arcs, node_types) positional forms of the reference constructors
        if arcs is not None and node_types is None and arcs and all(
                isinstance(a, tuple) and len(a) == 2 and isinstance(a[1], FactorType) for a in arcs):
            arcs, node_types = None, arcs
        super().__init__(bn_type, nodes, arcs, node_types, graph)
