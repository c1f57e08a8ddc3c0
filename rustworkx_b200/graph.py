"""Minimal stand-in for rustworkx's ``PyGraph`` / ``PyDiGraph`` containers.

The real containers (reference ``src/graph.rs:152-189``, ``src/digraph.rs:192-200``,
``src/lib.rs:201``) wrap petgraph's ``StableGraph`` and are OUT OF SCOPE for this repository
(SURVEY.md section 2b); no rustworkx wheel exists in this image, so this module supplies just the
semantics the all-sources-BFS hot path depends on:

* stable node / edge indices with holes after ``remove_node`` / ``remove_edge``
  (``src/graph.rs:883-888`` -- removing a missing node is silently ignored),
* LIFO reuse of vacated indices by later ``add_node`` / ``add_edge`` (petgraph 0.8 StableGraph free
  lists, documented at ``rustworkx/__init__.py:34-49``),
* the ``multigraph`` flag (``src/graph.rs:178-189``: a parallel edge updates the payload instead),
* the accessors the device snapshot is built from: ``node_indices()``, ``edge_indices()``,
  ``edge_list()`` (``src/graph.rs:418,462,815``).

Storage is numpy-backed so that 10^6-node / 10^7-edge benchmark inputs stay cheap.  Every edge also
carries an insertion sequence number so the CPU oracle can rebuild petgraph's adjacency iteration
order (most recently added edge first).
"""

from __future__ import annotations

import itertools

import numpy as np

_END = 0xFFFFFFFF


class NoEdgeBetweenNodes(Exception):
    """Mirror of ``rustworkx.NoEdgeBetweenNodes``."""


class InvalidNode(Exception):
    """Mirror of ``rustworkx.InvalidNode``."""


class NodeIndices(list):
    """List-like mirror of ``rustworkx.NodeIndices`` (``src/iterators.rs:627-648``)."""

    def __array__(self, dtype=None, copy=None):
        return np.asarray(list(self), dtype=dtype if dtype is not None else np.uint64)


class EdgeIndices(NodeIndices):
    """List-like mirror of ``rustworkx.EdgeIndices``."""


class EdgeList(list):
    """List-like mirror of ``rustworkx.EdgeList`` (``src/iterators.rs:469-480``)."""

    def __array__(self, dtype=None, copy=None):
        arr = np.asarray(list(self), dtype=dtype if dtype is not None else np.uint64)
        return arr.reshape(-1, 2)


def _check_index(x, what="index"):
    if isinstance(x, (bool, np.bool_)) or not isinstance(x, (int, np.integer)):
        raise TypeError(f"'{type(x).__name__}' object cannot be interpreted as an integer")
    if x < 0:
        raise OverflowError("can't convert negative int to unsigned")
    return int(x)


# Mutation stamps come from one process-wide counter, so a stamp is never handed out twice: not after
# clear() re-initialises a graph, not across graphs.  The device-snapshot cache (rustworkx_b200.
# device_snapshot) keys on it.
_VERSIONS = itertools.count(1)


class _StableGraph:
    _directed = False

    def __init__(self, multigraph=True, attrs=None, *, node_count_hint=0, edge_count_hint=0):
        self.multigraph = bool(multigraph)
        self.attrs = attrs
        ncap = max(16, int(node_count_hint))
        ecap = max(16, int(edge_count_hint))
        self._nalive = np.zeros(ncap, dtype=np.uint8)
        self._nlen = 0
        self._ndata = {}  # index -> payload (sparse: bulk generators leave payload = index)
        self._ndata_default_index = True  # missing payload means "payload == index"
        self._esrc = np.zeros(ecap, dtype=np.uint32)
        self._edst = np.zeros(ecap, dtype=np.uint32)
        self._ealive = np.zeros(ecap, dtype=np.uint8)
        self._eseq = np.zeros(ecap, dtype=np.uint64)
        self._elen = 0
        self._edata = {}  # index -> payload (missing means None)
        self._free_nodes = []  # LIFO
        self._free_edges = []  # LIFO
        self._seq = 0
        self._node_count = 0
        self._edge_count = 0
        self._edge_lookup = None  # (a, b) -> index, only for multigraph=False
        self._version = next(_VERSIONS)

    # ------------------------------------------------------------------ internals
    def _touch(self):
        self._version = next(_VERSIONS)

    def _grow_nodes(self, need):
        if need > self._nalive.shape[0]:
            cap = max(need, 2 * self._nalive.shape[0])
            self._nalive = np.concatenate(
                [self._nalive, np.zeros(cap - self._nalive.shape[0], dtype=np.uint8)]
            )

    def _grow_edges(self, need):
        cap0 = self._esrc.shape[0]
        if need > cap0:
            cap = max(need, 2 * cap0)
            pad = cap - cap0
            self._esrc = np.concatenate([self._esrc, np.zeros(pad, dtype=np.uint32)])
            self._edst = np.concatenate([self._edst, np.zeros(pad, dtype=np.uint32)])
            self._ealive = np.concatenate([self._ealive, np.zeros(pad, dtype=np.uint8)])
            self._eseq = np.concatenate([self._eseq, np.zeros(pad, dtype=np.uint64)])

    def _has_node(self, i):
        return 0 <= i < self._nlen and self._nalive[i] != 0

    def _key(self, a, b):
        if self._directed or a <= b:
            return (a, b)
        return (b, a)

    def _lookup(self):
        if self._edge_lookup is None:
            lut = {}
            for e in np.flatnonzero(self._ealive[: self._elen]):
                lut.setdefault(self._key(int(self._esrc[e]), int(self._edst[e])), int(e))
            self._edge_lookup = lut
        return self._edge_lookup

    def _find_edge(self, a, b):
        """Index of an edge a->b (either orientation when undirected), or None."""
        if not self.multigraph:
            return self._lookup().get(self._key(a, b))
        src = self._esrc[: self._elen]
        dst = self._edst[: self._elen]
        alive = self._ealive[: self._elen] != 0
        hit = (src == a) & (dst == b) & alive
        if not self._directed:
            hit |= (src == b) & (dst == a) & alive
        idx = np.flatnonzero(hit)
        if idx.size == 0:
            return None
        # petgraph find_edge walks a's outgoing list (most recent first)
        return int(idx[np.argmax(self._eseq[idx])])

    # ------------------------------------------------------------------ nodes
    def add_node(self, obj=None):
        if self._free_nodes:
            idx = self._free_nodes.pop()
        else:
            idx = self._nlen
            self._grow_nodes(idx + 1)
            self._nlen += 1
        self._nalive[idx] = 1
        self._ndata[idx] = obj
        self._node_count += 1
        self._touch()
        return idx

    def add_nodes_from(self, obj_list):
        return NodeIndices(self.add_node(o) for o in obj_list)

    def _bulk_add_nodes(self, count):
        """Append ``count`` nodes whose payload is their index (generator fast path)."""
        if self._free_nodes:
            return NodeIndices(self.add_node(None) for _ in range(count))
        start = self._nlen
        self._grow_nodes(start + count)
        self._nalive[start : start + count] = 1
        self._nlen += count
        self._node_count += count
        self._touch()
        return range(start, start + count)

    def remove_node(self, node):
        node = _check_index(node)
        if not self._has_node(node):
            return  # src/graph.rs:883-888: silently ignored
        # petgraph removes the outgoing list first, then the incoming list, each from its head
        # (most recently added edge first); this fixes the order of the edge free list.
        src = self._esrc[: self._elen]
        dst = self._edst[: self._elen]
        alive = self._ealive[: self._elen] != 0
        out_e = np.flatnonzero((src == node) & alive)
        out_e = out_e[np.argsort(-self._eseq[out_e].astype(np.int64), kind="stable")]
        for e in out_e:
            self._remove_edge_index(int(e))
        alive = self._ealive[: self._elen] != 0
        in_e = np.flatnonzero((dst == node) & alive)
        in_e = in_e[np.argsort(-self._eseq[in_e].astype(np.int64), kind="stable")]
        for e in in_e:
            self._remove_edge_index(int(e))
        self._nalive[node] = 0
        self._ndata.pop(node, None)
        self._free_nodes.append(node)
        self._node_count -= 1
        self._touch()

    def remove_nodes_from(self, index_list):
        for i in index_list:
            self.remove_node(i)

    def has_node(self, node):
        return self._has_node(_check_index(node))

    def num_nodes(self):
        return self._node_count

    def __len__(self):
        return self._node_count

    def node_indices(self):
        return NodeIndices(int(i) for i in np.flatnonzero(self._nalive[: self._nlen]))

    node_indexes = node_indices

    def nodes(self):
        return [self._node_payload(int(i)) for i in np.flatnonzero(self._nalive[: self._nlen])]

    def _node_payload(self, i):
        if i in self._ndata:
            return self._ndata[i]
        return i if self._ndata_default_index else None

    def __getitem__(self, idx):
        idx = _check_index(idx)
        if not self._has_node(idx):
            raise IndexError("No node found for index")
        return self._node_payload(idx)

    def __setitem__(self, idx, value):
        idx = _check_index(idx)
        if not self._has_node(idx):
            raise IndexError("No node found for index")
        self._ndata[idx] = value

    # ------------------------------------------------------------------ edges
    def _new_edge_slot(self):
        if self._free_edges:
            return self._free_edges.pop()
        idx = self._elen
        self._grow_edges(idx + 1)
        self._elen += 1
        return idx

    def add_edge(self, node_a, node_b, edge=None):
        a = _check_index(node_a)
        b = _check_index(node_b)
        if not self._has_node(a) or not self._has_node(b):
            raise IndexError("One of the endpoints of the edge does not exist in graph")
        if not self.multigraph:
            existing = self._find_edge(a, b)
            if existing is not None:  # src/graph.rs:178-189: update in place
                self._edata[existing] = edge
                return existing
        e = self._new_edge_slot()
        self._esrc[e] = a
        self._edst[e] = b
        self._ealive[e] = 1
        self._eseq[e] = self._seq
        self._seq += 1
        if edge is not None:
            self._edata[e] = edge
        self._edge_count += 1
        if self._edge_lookup is not None:
            self._edge_lookup.setdefault(self._key(a, b), e)
        self._touch()
        return e

    def add_edges_from(self, obj_list):
        return EdgeIndices(self.add_edge(a, b, w) for (a, b, w) in obj_list)

    def add_edges_from_no_data(self, obj_list):
        return EdgeIndices(self.add_edge(a, b, None) for (a, b) in obj_list)

    def extend_from_edge_list(self, edge_list):
        """``src/graph.rs:983-998``: missing node indices are created (payload ``None``)."""
        arr = np.asarray(edge_list)
        if arr.size == 0:
            return
        arr = arr.reshape(-1, 2)
        if arr.min() < 0:
            raise OverflowError("can't convert negative int to unsigned")
        max_index = int(arr.max())
        while max_index >= self._node_count:  # the reference compares against node_count()
            self.add_node(None)
        if not (self._nalive[np.unique(arr)] != 0).all():
            raise IndexError("One of the endpoints of the edge does not exist in graph")
        self._bulk_add_edges(arr[:, 0], arr[:, 1])

    def extend_from_weighted_edge_list(self, edge_list):
        pairs = [(a, b) for (a, b, _w) in edge_list]
        weights = [w for (_a, _b, w) in edge_list]
        if not pairs:
            return
        max_index = max(max(a, b) for a, b in pairs)
        while max_index >= self._node_count:
            self.add_node(None)
        for (a, b), w in zip(pairs, weights):
            self.add_edge(a, b, w)

    def _bulk_add_edges(self, src, dst):
        """Vectorised ``add_edge`` for payload-free edges (generators, edge lists)."""
        src = np.ascontiguousarray(src, dtype=np.uint32)
        dst = np.ascontiguousarray(dst, dtype=np.uint32)
        m = src.shape[0]
        if m == 0:
            return
        if self._free_edges or not self.multigraph:
            for a, b in zip(src.tolist(), dst.tolist()):
                self.add_edge(a, b, None)
            return
        start = self._elen
        self._grow_edges(start + m)
        self._esrc[start : start + m] = src
        self._edst[start : start + m] = dst
        self._ealive[start : start + m] = 1
        self._eseq[start : start + m] = np.arange(self._seq, self._seq + m, dtype=np.uint64)
        self._seq += m
        self._elen += m
        self._edge_count += m
        self._touch()

    def _remove_edge_index(self, e):
        if self._edge_lookup is not None:
            key = self._key(int(self._esrc[e]), int(self._edst[e]))
            if self._edge_lookup.get(key) == e:
                del self._edge_lookup[key]
        self._ealive[e] = 0
        self._edata.pop(e, None)
        self._free_edges.append(e)
        self._edge_count -= 1
        self._touch()

    def remove_edge(self, node_a, node_b):
        a = _check_index(node_a)
        b = _check_index(node_b)
        e = self._find_edge(a, b) if (self._has_node(a) and self._has_node(b)) else None
        if e is None:
            raise NoEdgeBetweenNodes("No edge found between nodes")
        self._remove_edge_index(e)

    def remove_edge_from_index(self, edge):
        e = _check_index(edge)
        if 0 <= e < self._elen and self._ealive[e]:
            self._remove_edge_index(e)

    def remove_edges_from(self, index_list):
        for a, b in index_list:
            self.remove_edge(a, b)

    def has_edge(self, node_a, node_b):
        a = _check_index(node_a)
        b = _check_index(node_b)
        return self._find_edge(a, b) is not None

    def get_edge_data(self, node_a, node_b):
        e = self._find_edge(_check_index(node_a), _check_index(node_b))
        if e is None:
            raise NoEdgeBetweenNodes("No edge found between nodes")
        return self._edata.get(e)

    def get_edge_data_by_index(self, edge_index):
        e = _check_index(edge_index)
        if not (0 <= e < self._elen and self._ealive[e]):
            raise IndexError(f"Provided edge index {e} is not present in the graph")
        return self._edata.get(e)

    def get_edge_endpoints_by_index(self, edge_index):
        e = _check_index(edge_index)
        if not (0 <= e < self._elen and self._ealive[e]):
            raise IndexError(f"Provided edge index {e} is not present in the graph")
        return (int(self._esrc[e]), int(self._edst[e]))

    def num_edges(self):
        return self._edge_count

    def _live_edges(self):
        return np.flatnonzero(self._ealive[: self._elen])

    def edge_indices(self):
        return EdgeIndices(int(e) for e in self._live_edges())

    def edge_list(self):
        idx = self._live_edges()
        return EdgeList(zip(self._esrc[idx].tolist(), self._edst[idx].tolist()))

    def weighted_edge_list(self):
        idx = self._live_edges()
        return [
            (int(self._esrc[e]), int(self._edst[e]), self._edata.get(int(e))) for e in idx
        ]

    def edges(self):
        return [self._edata.get(int(e)) for e in self._live_edges()]

    def edge_index_map(self):
        return {
            int(e): (int(self._esrc[e]), int(self._edst[e]), self._edata.get(int(e)))
            for e in self._live_edges()
        }

    # ------------------------------------------------------------------ structure queries
    def _live_node_array(self):
        """Ascending indices of the live nodes (uint32); no scan when no slot was ever vacated."""
        if self._node_count == self._nlen:
            return np.arange(self._nlen, dtype=np.uint32)
        return np.flatnonzero(self._nalive[: self._nlen]).astype(np.uint32)

    def _live_edge_array(self):
        if self._edge_count == self._elen:
            return np.arange(self._elen, dtype=np.uint32)
        return np.flatnonzero(self._ealive[: self._elen]).astype(np.uint32)

    def _bounds(self):
        """(node_bound, edge_bound) = highest live index + 1 (petgraph StableGraph)."""
        if self._node_count == self._nlen:
            nb = self._nlen
        else:
            live_n = np.flatnonzero(self._nalive[: self._nlen])
            nb = int(live_n[-1]) + 1 if live_n.size else 0
        if self._edge_count == self._elen:
            eb = self._elen
        else:
            live_e = np.flatnonzero(self._ealive[: self._elen])
            eb = int(live_e[-1]) + 1 if live_e.size else 0
        return nb, eb

    def _snapshot_arrays(self):
        """Arrays the C-ABI snapshot call consumes.  When no edge slot was
        ever vacated the endpoint arrays are read-only views of the container's own storage (the
        snapshot call copies them to the device before it returns): 1 ms instead of 260 ms for
        BA(10^6, 8)."""
        live_nodes = self._live_node_array()
        live_edges = self._live_edge_array()
        nb, eb = self._bounds()
        dense = self._edge_count == self._elen

        def take(a):
            if not dense:
                return np.ascontiguousarray(a[live_edges])
            v = a[: self._elen]
            v.setflags(write=False)
            return v

        return {
            "n_bound": nb,
            "e_bound": eb,
            "live_nodes": live_nodes,
            "edge_src": take(self._esrc),
            "edge_dst": take(self._edst),
            "edge_id": live_edges,
            # no edge slot was ever vacated: the edge index is the position in the list
            "edge_id_is_position": bool(dense),
            "edge_seq": take(self._eseq),
            "directed": self._directed,
        }

    def neighbors(self, node):
        node = _check_index(node)
        src = self._esrc[: self._elen]
        dst = self._edst[: self._elen]
        alive = self._ealive[: self._elen] != 0
        out = dst[(src == node) & alive]
        if self._directed:
            vals = out
        else:
            vals = np.concatenate([out, src[(dst == node) & alive]])
        seen = []
        for v in vals.tolist():
            if v not in seen:
                seen.append(v)
        return NodeIndices(seen)

    def degree(self, node):
        node = _check_index(node)
        src = self._esrc[: self._elen]
        dst = self._edst[: self._elen]
        alive = self._ealive[: self._elen] != 0
        return int(((src == node) & alive).sum() + ((dst == node) & alive).sum())

    def copy(self):
        import copy as _copy

        new = self.__class__.__new__(self.__class__)
        new.__dict__ = {}
        for k, v in self.__dict__.items():
            if isinstance(v, np.ndarray):
                new.__dict__[k] = v.copy()
            elif isinstance(v, (dict, list)):
                new.__dict__[k] = _copy.copy(v)
            else:
                new.__dict__[k] = v
        return new

    def clear(self):
        self.__init__(multigraph=self.multigraph, attrs=self.attrs)

    def clear_edges(self):
        for e in self._live_edges():
            self._remove_edge_index(int(e))


class PyGraph(_StableGraph):
    """Undirected multigraph stand-in (reference ``src/graph.rs:152-160``)."""

    _directed = False

    def __init__(self, multigraph=True, attrs=None, *, node_count_hint=None, edge_count_hint=None):
        super().__init__(
            multigraph,
            attrs,
            node_count_hint=node_count_hint or 0,
            edge_count_hint=edge_count_hint or 0,
        )

    def to_directed(self):
        """Each undirected edge (a, b) becomes a->b then b->a (``src/graph.rs`` ``to_directed``)."""
        new = PyDiGraph(multigraph=self.multigraph, attrs=self.attrs)
        nb, _ = self._bounds()
        new._grow_nodes(self._nlen)
        new._nalive[: self._nlen] = self._nalive[: self._nlen]
        new._nlen = self._nlen
        new._ndata = dict(self._ndata)
        new._ndata_default_index = self._ndata_default_index
        new._node_count = self._node_count
        new._free_nodes = list(self._free_nodes)
        for e in self._live_edges():
            a, b = int(self._esrc[e]), int(self._edst[e])
            w = self._edata.get(int(e))
            new.add_edge(a, b, w)
            new.add_edge(b, a, w)
        return new


class PyDiGraph(_StableGraph):
    """Directed multigraph stand-in (reference ``src/digraph.rs:192-200``)."""

    _directed = True

    def __init__(
        self,
        check_cycle=False,
        multigraph=True,
        attrs=None,
        *,
        node_count_hint=None,
        edge_count_hint=None,
    ):
        super().__init__(
            multigraph,
            attrs,
            node_count_hint=node_count_hint or 0,
            edge_count_hint=edge_count_hint or 0,
        )
        self.check_cycle = bool(check_cycle)

    def in_degree(self, node):
        node = _check_index(node)
        alive = self._ealive[: self._elen] != 0
        return int(((self._edst[: self._elen] == node) & alive).sum())

    def out_degree(self, node):
        node = _check_index(node)
        alive = self._ealive[: self._elen] != 0
        return int(((self._esrc[: self._elen] == node) & alive).sum())

    def successor_indices(self, node):
        return self.neighbors(node)

    def predecessor_indices(self, node):
        node = _check_index(node)
        alive = self._ealive[: self._elen] != 0
        vals = self._esrc[: self._elen][(self._edst[: self._elen] == node) & alive]
        seen = []
        for v in vals.tolist():
            if v not in seen:
                seen.append(v)
        return NodeIndices(seen)

    def to_undirected(self, multigraph=True):
        new = PyGraph(multigraph=multigraph, attrs=self.attrs)
        new._grow_nodes(self._nlen)
        new._nalive[: self._nlen] = self._nalive[: self._nlen]
        new._nlen = self._nlen
        new._ndata = dict(self._ndata)
        new._ndata_default_index = self._ndata_default_index
        new._node_count = self._node_count
        new._free_nodes = list(self._free_nodes)
        for e in self._live_edges():
            new.add_edge(int(self._esrc[e]), int(self._edst[e]), self._edata.get(int(e)))
        return new

    def clear(self):
        self.__init__(check_cycle=self.check_cycle, multigraph=self.multigraph, attrs=self.attrs)


class PyDAG(PyDiGraph):
    """``rustworkx.PyDAG`` is a plain subclass of ``PyDiGraph`` (``rustworkx/__init__.py:52-135``)."""
