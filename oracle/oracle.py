"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE ONLY -- see rx_oracle.h).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module.  The product package ``rustworkx_b200`` never does.

Parity status: pinned against the reference's golden vectors (tests/test_oracle_golden.py).
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "librxoracle.so")


def build(force=False):
    """Compile oracle/rx_oracle.c with gcc (recipe: oracle/Makefile)."""
    src = os.path.join(_HERE, "rx_oracle.c")
    hdr = os.path.join(_HERE, "rx_oracle.h")
    if (
        not force
        and os.path.exists(_LIB_PATH)
        and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(src), os.path.getmtime(hdr))
    ):
        return _LIB_PATH
    env = dict(os.environ)
    env.pop("CC", None)
    subprocess.run(["make", "-C", _HERE, "-B", "librxoracle.so"], check=True, env=env,
                   stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return _LIB_PATH


class Stats(C.Structure):
    _fields_ = [
        ("te", C.c_uint64),
        ("te_dag", C.c_uint64),
        ("v", C.c_uint64),
        ("sources", C.c_uint64),
        ("max_level", C.c_uint64),
    ]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    vp, u8p, u32p, u64p, f64p = (
        C.c_void_p,
        C.POINTER(C.c_uint8),
        C.POINTER(C.c_uint32),
        C.POINTER(C.c_uint64),
        C.POINTER(C.c_double),
    )
    sp = C.POINTER(Stats)
    L.rxo_graph_new.restype = vp
    L.rxo_graph_new.argtypes = [C.c_int]
    L.rxo_graph_free.argtypes = [vp]
    L.rxo_add_node.restype = C.c_uint32
    L.rxo_add_node.argtypes = [vp]
    L.rxo_add_edge.restype = C.c_uint32
    L.rxo_add_edge.argtypes = [vp, C.c_uint32, C.c_uint32]
    L.rxo_remove_edge.argtypes = [vp, C.c_uint32]
    L.rxo_remove_node.argtypes = [vp, C.c_uint32]
    for name in ("rxo_node_count", "rxo_edge_count", "rxo_node_bound", "rxo_edge_bound"):
        getattr(L, name).restype = C.c_size_t
        getattr(L, name).argtypes = [vp]
    L.rxo_graph_from_snapshot.restype = vp
    L.rxo_graph_from_snapshot.argtypes = [C.c_int, C.c_size_t, u8p, C.c_size_t, u32p, u32p, u8p, u64p]
    L.rxo_betweenness_centrality.argtypes = [vp, C.c_int, C.c_int, C.c_size_t, f64p, u8p, sp]
    L.rxo_edge_betweenness_centrality.argtypes = [vp, C.c_int, C.c_size_t, f64p, u8p, sp]
    L.rxo_closeness_centrality.argtypes = [vp, C.c_int, C.c_size_t, f64p, u8p, sp]
    L.rxo_distance_matrix.argtypes = [vp, C.c_size_t, C.c_int, C.c_double, f64p, sp]
    L.rxo_betweenness_partial.argtypes = [vp, u32p, C.c_size_t, C.c_int, C.c_int, f64p, sp]
    L.rxo_edge_betweenness_partial.argtypes = [vp, u32p, C.c_size_t, C.c_int, f64p, sp]
    L.rxo_closeness_partial.argtypes = [vp, u32p, C.c_size_t, C.c_int, C.c_int, f64p, sp]
    L.rxo_average_shortest_path_length.restype = C.c_double
    L.rxo_average_shortest_path_length.argtypes = [vp, C.c_size_t, C.c_int, C.c_int]
    L.rxo_distance_sum.argtypes = [vp, C.c_size_t, C.c_int, u64p, u64p]
    L.rxo_group_closeness_centrality.restype = C.c_double
    L.rxo_group_closeness_centrality.argtypes = [vp, u32p, C.c_size_t]
    L.rxo_group_betweenness_centrality.restype = C.c_double
    L.rxo_group_betweenness_centrality.argtypes = [vp, u32p, C.c_size_t, C.c_int, C.c_size_t]
    L.rxo_group_betweenness_partial.restype = C.c_double
    L.rxo_group_betweenness_partial.argtypes = [vp, u32p, C.c_size_t, u32p, C.c_size_t, C.c_int, sp]
    L.rxo_newman_weighted_closeness_centrality.argtypes = [vp, f64p, C.c_int, C.c_size_t, f64p, u8p, sp]
    L.rxo_weighted_closeness_partial.argtypes = [vp, f64p, C.c_int, C.c_int, u32p, C.c_size_t, C.c_int,
                                                 C.c_int, f64p, sp]
    L.rxo_dijkstra_lengths_rows.argtypes = [vp, f64p, u32p, C.c_size_t, C.c_int, f64p, sp]
    L.rxo_rescale.argtypes = [f64p, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_int]
    L.rxo_max_threads.restype = C.c_int
    _lib = L
    return L


def _p(arr, ctype):
    return arr.ctypes.data_as(C.POINTER(ctype))


class OracleGraph:
    """A petgraph-StableGraph-like graph living inside the oracle library."""

    def __init__(self, directed=False, _handle=None):
        self._L = lib()
        self._h = _handle if _handle is not None else self._L.rxo_graph_new(1 if directed else 0)
        self.directed = bool(directed)

    def __del__(self):
        try:
            if self._h:
                self._L.rxo_graph_free(self._h)
                self._h = None
        except Exception:
            pass

    # -- construction -------------------------------------------------------------------------
    @classmethod
    def from_graph(cls, graph):
        """Snapshot a ``rustworkx_b200`` stand-in graph, keeping petgraph's adjacency order."""
        snap = graph._snapshot_arrays()
        return cls.from_arrays(
            snap["directed"], snap["n_bound"], snap["live_nodes"], snap["e_bound"],
            snap["edge_src"], snap["edge_dst"], snap["edge_id"], snap["edge_seq"],
        )

    @classmethod
    def from_arrays(cls, directed, n_bound, live_nodes, e_bound, esrc, edst, eid=None, eseq=None):
        L = lib()
        n_bound = int(n_bound)
        e_bound = int(e_bound)
        nalive = np.zeros(max(n_bound, 1), dtype=np.uint8)
        nalive[np.asarray(live_nodes, dtype=np.int64)] = 1
        esrc = np.asarray(esrc, dtype=np.uint32)
        edst = np.asarray(edst, dtype=np.uint32)
        if eid is None:
            eid = np.arange(esrc.shape[0], dtype=np.uint32)
        eid = np.asarray(eid, dtype=np.int64)
        d_src = np.zeros(max(e_bound, 1), dtype=np.uint32)
        d_dst = np.zeros(max(e_bound, 1), dtype=np.uint32)
        d_alive = np.zeros(max(e_bound, 1), dtype=np.uint8)
        d_seq = np.zeros(max(e_bound, 1), dtype=np.uint64)
        d_src[eid] = esrc
        d_dst[eid] = edst
        d_alive[eid] = 1
        d_seq[eid] = np.asarray(eseq, dtype=np.uint64) if eseq is not None else eid.astype(np.uint64)
        h = L.rxo_graph_from_snapshot(
            1 if directed else 0, n_bound, _p(nalive, C.c_uint8), e_bound,
            _p(d_src, C.c_uint32), _p(d_dst, C.c_uint32), _p(d_alive, C.c_uint8),
            _p(d_seq, C.c_uint64),
        )
        return cls(directed, _handle=h)

    @classmethod
    def from_edges(cls, directed, edges, num_nodes=None):
        """petgraph ``Graph::from_edges``: nodes 0..=max index, edges in the given order."""
        g = cls(directed)
        edges = list(edges)
        n = (max(max(a, b) for a, b in edges) + 1) if edges else 0
        if num_nodes is not None:
            n = max(n, num_nodes)
        for _ in range(n):
            g.add_node()
        for a, b in edges:
            g.add_edge(a, b)
        return g

    def add_node(self):
        return int(self._L.rxo_add_node(self._h))

    def add_edge(self, a, b):
        e = int(self._L.rxo_add_edge(self._h, a, b))
        if e == 0xFFFFFFFF:
            raise IndexError("endpoint does not exist")
        return e

    def remove_edge(self, e):
        return self._L.rxo_remove_edge(self._h, e)

    def remove_node(self, a):
        return self._L.rxo_remove_node(self._h, a)

    def node_count(self):
        return int(self._L.rxo_node_count(self._h))

    def edge_count(self):
        return int(self._L.rxo_edge_count(self._h))

    def node_bound(self):
        return int(self._L.rxo_node_bound(self._h))

    def edge_bound(self):
        return int(self._L.rxo_edge_bound(self._h))

    # -- the four hot-path functions: return Vec<Option<f64>> as a list with None holes ----------
    @staticmethod
    def _opt(vals, present):
        return [float(v) if p else None for v, p in zip(vals, present)]

    def betweenness_centrality(self, include_endpoints, normalized, parallel_threshold=50,
                               return_stats=False, as_arrays=False):
        nb = self.node_bound()
        out = np.zeros(max(nb, 1), dtype=np.float64)
        present = np.zeros(max(nb, 1), dtype=np.uint8)
        st = Stats()
        self._L.rxo_betweenness_centrality(
            self._h, int(include_endpoints), int(normalized), int(parallel_threshold),
            _p(out, C.c_double), _p(present, C.c_uint8), C.byref(st))
        res = (out[:nb], present[:nb]) if as_arrays else self._opt(out[:nb], present[:nb])
        return (res, st.as_dict()) if return_stats else res

    def edge_betweenness_centrality(self, normalized, parallel_threshold=50, return_stats=False,
                                    as_arrays=False):
        eb = self.edge_bound()
        out = np.zeros(max(eb, 1), dtype=np.float64)
        present = np.zeros(max(eb, 1), dtype=np.uint8)
        st = Stats()
        self._L.rxo_edge_betweenness_centrality(
            self._h, int(normalized), int(parallel_threshold), _p(out, C.c_double),
            _p(present, C.c_uint8), C.byref(st))
        res = (out[:eb], present[:eb]) if as_arrays else self._opt(out[:eb], present[:eb])
        return (res, st.as_dict()) if return_stats else res

    def closeness_centrality(self, wf_improved, parallel_threshold=50, return_stats=False,
                             as_arrays=False):
        nb = self.node_bound()
        out = np.zeros(max(nb, 1), dtype=np.float64)
        present = np.zeros(max(nb, 1), dtype=np.uint8)
        st = Stats()
        self._L.rxo_closeness_centrality(
            self._h, int(wf_improved), int(parallel_threshold), _p(out, C.c_double),
            _p(present, C.c_uint8), C.byref(st))
        res = (out[:nb], present[:nb]) if as_arrays else self._opt(out[:nb], present[:nb])
        return (res, st.as_dict()) if return_stats else res

    def distance_matrix(self, parallel_threshold=300, as_undirected=False, null_value=0.0,
                        return_stats=False):
        n = self.node_count()
        out = np.zeros((n, n), dtype=np.float64)
        st = Stats()
        if n:
            self._L.rxo_distance_matrix(
                self._h, int(parallel_threshold), int(as_undirected), float(null_value),
                _p(out, C.c_double), C.byref(st))
        return (out, st.as_dict()) if return_stats else out

    def average_shortest_path_length(self, parallel_threshold=300, as_undirected=False,
                                     disconnected=False):
        return float(self._L.rxo_average_shortest_path_length(
            self._h, int(parallel_threshold), int(as_undirected), int(disconnected)))

    def distance_sum(self, parallel_threshold=300, as_undirected=False):
        a, b = C.c_uint64(), C.c_uint64()
        self._L.rxo_distance_sum(self._h, int(parallel_threshold), int(as_undirected),
                                 C.byref(a), C.byref(b))
        return int(a.value), int(b.value)

    # -- group centralities ------------------------------------------------------
    def group_closeness_centrality(self, group):
        grp = np.ascontiguousarray(group, dtype=np.uint32)
        return float(self._L.rxo_group_closeness_centrality(self._h, _p(grp, C.c_uint32), grp.shape[0]))

    def group_betweenness_centrality(self, group, normalized=True, parallel_threshold=50):
        grp = np.ascontiguousarray(group, dtype=np.uint32)
        return float(self._L.rxo_group_betweenness_centrality(
            self._h, _p(grp, C.c_uint32), grp.shape[0], int(normalized), int(parallel_threshold)))

    def group_betweenness_partial(self, group, sources, threads=1):
        grp = np.ascontiguousarray(group, dtype=np.uint32)
        src = np.ascontiguousarray(sources, dtype=np.uint32)
        st = Stats()
        val = float(self._L.rxo_group_betweenness_partial(
            self._h, _p(grp, C.c_uint32), grp.shape[0], _p(src, C.c_uint32), src.shape[0],
            int(threads), C.byref(st)))
        return val, st.as_dict()

    # -- weighted shortest path lengths -------------------------------------------
    def _edge_array(self, weight_by_edge):
        w = np.zeros(max(self.edge_bound(), 1), dtype=np.float64)
        w[: self.edge_bound()] = np.asarray(weight_by_edge, dtype=np.float64)[: self.edge_bound()]
        return w

    def newman_weighted_closeness_centrality(self, weight_by_edge, wf_improved=True,
                                             parallel_threshold=50, as_arrays=False):
        nb = self.node_bound()
        w = self._edge_array(weight_by_edge)
        out = np.zeros(max(nb, 1), dtype=np.float64)
        present = np.zeros(max(nb, 1), dtype=np.uint8)
        st = Stats()
        self._L.rxo_newman_weighted_closeness_centrality(
            self._h, _p(w, C.c_double), int(wf_improved), int(parallel_threshold),
            _p(out, C.c_double), _p(present, C.c_uint8), C.byref(st))
        return (out[:nb], present[:nb]) if as_arrays else self._opt(out[:nb], present[:nb])

    def weighted_closeness_partial(self, weight_by_edge, sources, invert=True, reversed_=True,
                                   wf_improved=True, threads=1):
        w = self._edge_array(weight_by_edge)
        src = np.ascontiguousarray(sources, dtype=np.uint32)
        out = np.zeros(max(self.node_bound(), 1), dtype=np.float64)
        st = Stats()
        self._L.rxo_weighted_closeness_partial(
            self._h, _p(w, C.c_double), int(invert), int(reversed_), _p(src, C.c_uint32), src.shape[0],
            int(wf_improved), int(threads), _p(out, C.c_double), C.byref(st))
        return out[: self.node_bound()], st.as_dict()

    def dijkstra_lengths_rows(self, cost_by_edge, sources, threads=1):
        """rows[i, v] = length sources[i] -> v by ORIGINAL node index, NaN where there is no entry."""
        w = self._edge_array(cost_by_edge)
        src = np.ascontiguousarray(sources, dtype=np.uint32)
        nb = self.node_bound()
        out = np.zeros((src.shape[0], max(nb, 1)), dtype=np.float64)
        st = Stats()
        self._L.rxo_dijkstra_lengths_rows(self._h, _p(w, C.c_double), _p(src, C.c_uint32), src.shape[0],
                                          int(threads), _p(out, C.c_double), C.byref(st))
        return out[:, :nb]

    # -- partial (source subset, un-rescaled) --------------------------------------------------
    def betweenness_partial(self, sources, include_endpoints=False, threads=1):
        src = np.ascontiguousarray(sources, dtype=np.uint32)
        out = np.zeros(max(self.node_bound(), 1), dtype=np.float64)
        st = Stats()
        self._L.rxo_betweenness_partial(self._h, _p(src, C.c_uint32), src.shape[0],
                                        int(include_endpoints), int(threads),
                                        _p(out, C.c_double), C.byref(st))
        return out[: self.node_bound()], st.as_dict()

    def edge_betweenness_partial(self, sources, threads=1):
        src = np.ascontiguousarray(sources, dtype=np.uint32)
        out = np.zeros(max(self.edge_bound(), 1), dtype=np.float64)
        st = Stats()
        self._L.rxo_edge_betweenness_partial(self._h, _p(src, C.c_uint32), src.shape[0],
                                             int(threads), _p(out, C.c_double), C.byref(st))
        return out[: self.edge_bound()], st.as_dict()

    def closeness_partial(self, sources, wf_improved=True, threads=1):
        src = np.ascontiguousarray(sources, dtype=np.uint32)
        out = np.zeros(max(self.node_bound(), 1), dtype=np.float64)
        st = Stats()
        self._L.rxo_closeness_partial(self._h, _p(src, C.c_uint32), src.shape[0],
                                      int(wf_improved), int(threads), _p(out, C.c_double),
                                      C.byref(st))
        return out[: self.node_bound()], st.as_dict()


def rescale(x, node_count, normalized, directed, include_endpoints):
    x = np.ascontiguousarray(x, dtype=np.float64).copy()
    lib().rxo_rescale(_p(x, C.c_double), x.shape[0], int(node_count), int(normalized),
                      int(directed), int(include_endpoints))
    return x


def max_threads():
    return int(lib().rxo_max_threads())
