"""ctypes binding of ``librxcuda.so`` (the C-ABI declared in ``include/rxcuda.h``).

This is the Python-side stub a maintainer would otherwise write in Rust inside a ``rustworkx-cuda``
crate.  The shared library is built in-tree by
``rustworkx_b200/csrc/Makefile`` (``__graft_entry__.build()``); if it is missing, or no CUDA device
is usable, every hot-path call raises -- there is NO CPU fallback.
"""

from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RXCUDA_LIB") or os.path.join(_HERE, "librxcuda.so")

RXC_OK = 0
RXC_ERR_INVALID = -1
RXC_ERR_CUDA = -2
RXC_ERR_OOM = -3
RXC_ERR_NCCL = -4
RXC_ERR_LIMIT = -5


class RxCudaError(RuntimeError):
    """Raised when librxcuda reports a CUDA / NCCL / argument error."""


class Stats(C.Structure):
    _fields_ = [
        ("te", C.c_uint64),
        ("te_dag", C.c_uint64),
        ("v", C.c_uint64),
        ("sources", C.c_uint64),
        ("levels", C.c_uint64),
        ("launches", C.c_uint64),
        ("batches", C.c_uint64),
        ("ms_total", C.c_double),
        ("ms_forward", C.c_double),
        ("ms_backward", C.c_double),
        ("ms_snapshot", C.c_double),
        ("ms_d2h", C.c_double),
        ("ms_reduce", C.c_double),
        ("ms_reduce_wait", C.c_double),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


# every symbol include/rxcuda.h declares: name -> (restype, argtypes)
_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)
_f64p = C.POINTER(C.c_double)
_vp = C.c_void_p
SYMBOLS = {
    "rxc_version": (C.c_char_p, []),
    "rxc_last_error": (C.c_char_p, []),
    "rxc_device_count": (C.c_int, []),
    "rxc_set_device": (C.c_int, [C.c_int]),
    "rxc_get_device": (C.c_int, []),
    "rxc_set_devices": (C.c_int, [C.POINTER(C.c_int), C.c_int]),
    "rxc_graph_snapshot": (C.c_int, [C.c_uint32, _u32p, C.c_uint32, C.c_uint32, _u32p, _u32p, _u32p,
                                     C.c_uint64, C.c_int, C.POINTER(_vp)]),
    "rxc_graph_free": (None, [_vp]),
    "rxc_graph_snapshot_cached": (C.c_int, [C.c_uint32, _u32p, C.c_uint32, C.c_uint32, _u32p, _u32p, _u32p,
                                            C.c_uint64, C.c_int, C.POINTER(_vp), C.POINTER(C.c_int)]),
    "rxc_snapshot_cache_clear": (C.c_int, []),
    "rxc_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(_vp)]),
    "rxc_host_free": (None, [_vp]),
    "rxc_graph_info": (C.c_int, [_vp, _u32p, _u32p, _u32p, C.POINTER(C.c_uint64),
                                 C.POINTER(C.c_int)]),
    "rxc_betweenness": (C.c_int, [_vp, C.c_int, C.c_int, _f64p]),
    "rxc_edge_betweenness": (C.c_int, [_vp, C.c_int, _f64p]),
    "rxc_closeness": (C.c_int, [_vp, C.c_int, _f64p]),
    "rxc_distance_matrix": (C.c_int, [_vp, C.c_int, C.c_double, _f64p]),
    "rxc_distance_sums": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "rxc_group_betweenness": (C.c_int, [_vp, _u32p, C.c_uint64, C.c_int, _f64p]),
    "rxc_group_closeness": (C.c_int, [_vp, _u32p, C.c_uint64, _f64p]),
    "rxc_group_betweenness_partial": (C.c_int, [_vp, _u32p, C.c_uint64, _u32p, C.c_uint64, _f64p]),
    "rxc_newman_weighted_closeness": (C.c_int, [_vp, _f64p, C.c_int, _f64p]),
    "rxc_all_pairs_dijkstra_lengths_rows": (C.c_int, [_vp, _f64p, C.c_double, C.c_uint32, C.c_uint32, _f64p]),
    "rxc_weighted_closeness_partial": (C.c_int, [_vp, _f64p, C.c_int, C.c_int, _u32p, C.c_uint64, C.c_int,
                                                 _f64p]),
    "rxc_betweenness_partial": (C.c_int, [_vp, _u32p, C.c_uint64, C.c_int, _f64p]),
    "rxc_edge_betweenness_partial": (C.c_int, [_vp, _u32p, C.c_uint64, _f64p]),
    "rxc_closeness_partial": (C.c_int, [_vp, _u32p, C.c_uint64, C.c_int, _f64p]),
    "rxc_distance_matrix_rows": (C.c_int, [_vp, C.c_int, C.c_double, C.c_uint32, C.c_uint32, _f64p]),
    "rxc_rescale": (None, [_f64p, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int]),
    "rxc_comm_unique_id": (C.c_int, [_u8p]),
    "rxc_comm_init": (C.c_int, [C.c_int, C.c_int, _u8p]),
    "rxc_comm_reduce_sum_f64": (C.c_int, [_f64p, C.c_uint64, C.c_int]),
    "rxc_betweenness_partial_reduce": (C.c_int, [_vp, _u32p, C.c_uint64, C.c_int, C.c_int, _f64p]),
    "rxc_betweenness_source_order": (C.c_int, [_vp, _u32p, C.c_uint64, _u32p]),
    "rxc_edge_betweenness_partial_reduce": (C.c_int, [_vp, _u32p, C.c_uint64, C.c_int, _f64p]),
    "rxc_comm_barrier": (C.c_int, []),
    "rxc_comm_destroy": (None, []),
    "rxc_get_stats": (C.c_int, [_vp, C.POINTER(Stats)]),
    "rxc_set_option": (C.c_int, [C.c_char_p, C.c_int64]),
}

_lib = None

# include/rxgen.h: the benchmark input generators live in their own host-only library
GEN_LIB_PATH = os.environ.get("RXGEN_LIB") or os.path.join(_HERE, "librxgen.so")
GEN_SYMBOLS = {
    "rxc_gen_gnp": (C.c_int64, [C.c_uint32, C.c_double, C.c_uint64, C.c_int, _u32p, _u32p,
                                C.c_uint64]),
    "rxc_gen_gnm": (C.c_int64, [C.c_uint32, C.c_uint64, C.c_uint64, C.c_int, _u32p, _u32p,
                                C.c_uint64]),
    "rxc_gen_barabasi_albert": (C.c_int64, [C.c_uint32, C.c_uint32, C.c_uint64, _u32p, _u32p,
                                            C.c_uint64]),
}
_gen = None


def gen_lib():
    """Load librxgen.so (once): seeded random-graph generators for benchmark inputs, no CUDA."""
    global _gen
    if _gen is None:
        if not os.path.exists(GEN_LIB_PATH):
            raise RxCudaError(f"{GEN_LIB_PATH} is missing: build it with `make -C rustworkx_b200/csrc`")
        G = C.CDLL(GEN_LIB_PATH)
        for name, (restype, argtypes) in GEN_SYMBOLS.items():
            fn = getattr(G, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _gen = G
    return _gen


def lib():
    """Load librxcuda.so (once).  Fails loudly when the extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RxCudaError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (or `make -C rustworkx_b200/csrc`).  rustworkx_b200 has no CPU fallback."
        )
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (restype, argtypes) in SYMBOLS.items():
        fn = getattr(L, name)
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = L
    env = os.environ.get("RXCUDA_DEVICES")
    if env:
        set_devices("all" if env.strip().lower() == "all" else [int(x) for x in env.split(",") if x.strip()])
    return L


def check(rc):
    if rc == RXC_OK:
        return
    msg = lib().rxc_last_error().decode("utf-8", "replace")
    if rc == RXC_ERR_OOM:
        raise MemoryError(msg)
    raise RxCudaError(f"librxcuda error {rc}: {msg}")


def p_u32(a):
    return a.ctypes.data_as(_u32p)


def p_f64(a):
    return a.ctypes.data_as(_f64p)


def p_u8(a):
    return a.ctypes.data_as(_u8p)


def device_count():
    return int(lib().rxc_device_count())


# Page-locked blocks that went out of use, by (rounded) size.  Large RESULTS (a 333 MB distance matrix,
# 134 MB of edge scores) are handed out in page-locked memory so that the DMA engine writes them at
# PCIe speed; page-locking itself costs more than the copy, so released blocks are kept for the next
# result of that size (at most POOL_CAP bytes in total; RXCUDA_PINNED_RESULTS=0 turns it all off).
_POOL = {}
_POOL_BYTES = 0
POOL_CAP = int(os.environ.get("RXCUDA_PINNED_POOL_BYTES", 4 << 30))
POOL_GRAIN = 1 << 20
PINNED_RESULTS = os.environ.get("RXCUDA_PINNED_RESULTS", "1") != "0"
PINNED_RESULT_MIN = 4 << 20


class _PinnedBlock:
    """Owner of one ``rxc_host_alloc`` block; arrays made by :func:`pinned_empty` keep it alive."""

    def __init__(self, nbytes, pooled=False):
        global _POOL_BYTES
        self.pooled = bool(pooled)
        self.nbytes = int(nbytes)
        self.ptr = _vp()
        if self.pooled:
            self.nbytes = (self.nbytes + POOL_GRAIN - 1) // POOL_GRAIN * POOL_GRAIN
            free = _POOL.get(self.nbytes)
            if free:
                self.ptr = _vp(free.pop())
                _POOL_BYTES -= self.nbytes
                return
        check(lib().rxc_host_alloc(self.nbytes, C.byref(self.ptr)))

    def __del__(self):
        global _POOL_BYTES
        try:
            if self.ptr:
                if self.pooled and _POOL_BYTES + self.nbytes <= POOL_CAP:
                    _POOL.setdefault(self.nbytes, []).append(self.ptr.value)
                    _POOL_BYTES += self.nbytes
                else:
                    lib().rxc_host_free(self.ptr)
                self.ptr = None
        except Exception:
            pass


def pinned_pool_clear():
    """Give every pooled page-locked block back to the driver."""
    global _POOL_BYTES
    for blocks in _POOL.values():
        for ptr in blocks:
            lib().rxc_host_free(_vp(ptr))
    _POOL.clear()
    _POOL_BYTES = 0


def result_empty(shape, dtype=np.float64):
    """Buffer for a result the library writes: page-locked (pooled) when it is large and a device is
    there, plain numpy otherwise."""
    dtype = np.dtype(dtype)
    shape = (int(shape),) if np.isscalar(shape) else tuple(int(x) for x in shape)
    nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
    if PINNED_RESULTS and nbytes >= PINNED_RESULT_MIN:
        try:
            return pinned_empty(shape, dtype, pooled=True)
        except (RxCudaError, MemoryError):
            pass
    return np.empty(shape, dtype=dtype)


def pinned_empty(shape, dtype, pooled=False):
    """numpy array in page-locked host memory (``rxc_host_alloc``): the library copies such buffers
    by direct DMA instead of staging them through its own pinned slots."""
    dtype = np.dtype(dtype)
    shape = (int(shape),) if np.isscalar(shape) else tuple(int(x) for x in shape)
    nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
    block = _PinnedBlock(max(nbytes, 1), pooled)
    buf = (C.c_char * max(nbytes, 1)).from_address(block.ptr.value)  # (a pooled block may be larger)
    buf._rxc_owner = block  # the ctypes buffer is the array's base: ties the block's lifetime to it
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape, dtype=np.int64))).reshape(shape)


def pinned_copy(a):
    """Copy of ``a`` in page-locked host memory."""
    a = np.ascontiguousarray(a)
    out = pinned_empty(a.shape, a.dtype)
    out[...] = a
    return out


def set_device(device):
    check(lib().rxc_set_device(int(device)))


def set_devices(devices):
    """In-process multi-GPU: replicate later snapshots on these devices and shard full calls over
    them ([] or None: back to the single current device; "all": every visible device)."""
    if devices == "all":
        devices = list(range(device_count()))
    devices = list(devices or [])
    arr = (C.c_int * max(len(devices), 1))(*devices)
    check(lib().rxc_set_devices(arr, len(devices)))


def set_option(name, value):
    check(lib().rxc_set_option(name.encode(), int(value)))


class DeviceGraph:
    """An ``rxc_graph*``: the StableGraph snapshot as compact CSR/CSC resident in HBM."""

    def __init__(self, n_bound, live_nodes, e_bound, edge_src, edge_dst, edge_id, directed, cached=False):
        """``edge_id=None``: the edge index is the position in the list (no edge was ever removed);
        the C-ABI then gets NULL and nothing is copied for it.  ``cached=True`` goes through
        ``rxc_graph_snapshot_cached`` (content-keyed; ``self.cache_hit`` tells what happened)."""
        L = lib()
        self.live_nodes = np.ascontiguousarray(live_nodes, dtype=np.uint32)
        src = np.ascontiguousarray(edge_src, dtype=np.uint32)
        dst = np.ascontiguousarray(edge_dst, dtype=np.uint32)
        self._edge_id = None if edge_id is None else np.ascontiguousarray(edge_id, dtype=np.uint32)
        self._n_edges = int(src.shape[0])
        self.n_bound = int(n_bound)
        self.e_bound = int(e_bound)
        self.n_live = int(self.live_nodes.shape[0])
        self.directed = bool(directed)
        h = _vp()
        self.cache_hit = False
        args = (self.n_bound, p_u32(self.live_nodes), self.n_live, self.e_bound, p_u32(src), p_u32(dst),
                None if self._edge_id is None else p_u32(self._edge_id), src.shape[0], 1 if directed else 0)
        if cached:
            hit = C.c_int(0)
            check(L.rxc_graph_snapshot_cached(*args, C.byref(h), C.byref(hit)))
            self.cache_hit = bool(hit.value)
        else:
            check(L.rxc_graph_snapshot(*args, C.byref(h)))
        self._h = h

    @property
    def edge_id(self):
        """original edge index of every listed edge"""
        if self._edge_id is None:
            self._edge_id = np.arange(self._n_edges, dtype=np.uint32)
        return self._edge_id

    @classmethod
    def from_graph(cls, graph, cached=False):
        """Snapshot a PyGraph / PyDiGraph (stand-in or, when importable, a real rustworkx graph --
        both expose node_indices() / edge_indices() / edge_list(), src/graph.rs:418,462,815)."""
        if hasattr(graph, "_snapshot_arrays"):
            s = graph._snapshot_arrays()
            return cls(s["n_bound"], s["live_nodes"], s["e_bound"], s["edge_src"], s["edge_dst"],
                       None if s.get("edge_id_is_position") else s["edge_id"], s["directed"], cached=cached)
        nodes = np.asarray(graph.node_indices(), dtype=np.uint32)
        eids = np.asarray(graph.edge_indices(), dtype=np.uint32)
        el = np.asarray(graph.edge_list(), dtype=np.uint32).reshape(-1, 2)
        directed = type(graph).__name__ in ("PyDiGraph", "PyDAG")
        nb = int(nodes[-1]) + 1 if nodes.size else 0
        eb = int(eids[-1]) + 1 if eids.size else 0
        return cls(nb, nodes, eb, el[:, 0], el[:, 1], eids, directed, cached=cached)

    def close(self):
        if getattr(self, "_h", None):
            lib().rxc_graph_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def stats(self):
        st = Stats()
        check(lib().rxc_get_stats(self._h, C.byref(st)))
        return st.as_dict()

    # ---- full calls -------------------------------------------------------------------------
    def betweenness(self, include_endpoints, normalized):
        out = np.zeros(max(self.n_bound, 1), dtype=np.float64)
        check(lib().rxc_betweenness(self._h, int(include_endpoints), int(normalized), p_f64(out)))
        return out[: self.n_bound]

    def edge_betweenness(self, normalized):
        out = result_empty(max(self.e_bound, 1))  # the library writes all e_bound entries
        check(lib().rxc_edge_betweenness(self._h, int(normalized), p_f64(out)))
        return out[: self.e_bound]

    def closeness(self, wf_improved):
        out = np.zeros(max(self.n_bound, 1), dtype=np.float64)
        check(lib().rxc_closeness(self._h, int(wf_improved), p_f64(out)))
        return out[: self.n_bound]

    def distance_matrix(self, as_undirected, null_value):
        out = result_empty((self.n_live, self.n_live))
        if self.n_live:
            check(lib().rxc_distance_matrix(self._h, int(as_undirected), float(null_value), p_f64(out)))
        return out

    def distance_sums(self, as_undirected):
        total, pairs = C.c_uint64(), C.c_uint64()
        check(lib().rxc_distance_sums(self._h, int(as_undirected), C.byref(total), C.byref(pairs)))
        return int(total.value), int(pairs.value)

    def group_betweenness(self, group, normalized):
        grp = np.ascontiguousarray(group, dtype=np.uint32)
        out = C.c_double()
        check(lib().rxc_group_betweenness(self._h, p_u32(grp), grp.shape[0], int(normalized), C.byref(out)))
        return float(out.value)

    def group_closeness(self, group):
        grp = np.ascontiguousarray(group, dtype=np.uint32)
        out = C.c_double()
        check(lib().rxc_group_closeness(self._h, p_u32(grp), grp.shape[0], C.byref(out)))
        return float(out.value)

    def group_betweenness_partial(self, group, sources):
        grp = np.ascontiguousarray(group, dtype=np.uint32)
        src = np.ascontiguousarray(sources, dtype=np.uint32)
        out = C.c_double()
        check(lib().rxc_group_betweenness_partial(self._h, p_u32(grp), grp.shape[0], p_u32(src),
                                                  src.shape[0], C.byref(out)))
        return float(out.value)

    # ---- weighted shortest path lengths (edge_weight: e_bound doubles by original edge index) ----
    def _edge_weights(self, edge_weight):
        w = np.zeros(max(self.e_bound, 1), dtype=np.float64)
        w[: self.e_bound] = np.asarray(edge_weight, dtype=np.float64)[: self.e_bound]
        return w

    def newman_weighted_closeness(self, edge_weight, wf_improved):
        w = self._edge_weights(edge_weight)
        out = np.zeros(max(self.n_bound, 1), dtype=np.float64)
        check(lib().rxc_newman_weighted_closeness(self._h, p_f64(w), int(wf_improved), p_f64(out)))
        return out[: self.n_bound]

    def weighted_closeness_partial(self, edge_weight, sources, invert=True, reversed_=True, wf_improved=True):
        w = self._edge_weights(edge_weight)
        src = np.ascontiguousarray(sources, dtype=np.uint32)
        out = np.zeros(max(src.shape[0], 1), dtype=np.float64)
        check(lib().rxc_weighted_closeness_partial(self._h, p_f64(w), int(invert), int(reversed_), p_u32(src),
                                                   src.shape[0], int(wf_improved), p_f64(out)))
        return out[: src.shape[0]]

    def all_pairs_dijkstra_lengths_rows(self, edge_weight, row_begin=0, row_end=None,
                                        unreachable_value=float("nan")):
        w = self._edge_weights(edge_weight)
        row_end = self.n_live if row_end is None else row_end
        out = result_empty((row_end - row_begin, self.n_live))
        if out.size:
            check(lib().rxc_all_pairs_dijkstra_lengths_rows(self._h, p_f64(w), float(unreachable_value),
                                                            int(row_begin), int(row_end), p_f64(out)))
        return out

    # ---- partial calls (source shards; un-rescaled) ---------------------------------------------
    def betweenness_partial(self, sources, include_endpoints=False, out=None):
        """``out``: optional caller buffer of at least n_bound doubles (e.g. ``pinned_empty``)."""
        src = np.ascontiguousarray(sources, dtype=np.uint32)
        if out is None:
            out = np.zeros(max(self.n_bound, 1), dtype=np.float64)
        elif out.dtype != np.float64 or not out.flags.c_contiguous or out.size < self.n_bound:
            raise ValueError("out must be a C-contiguous float64 array of at least n_bound elements")
        check(lib().rxc_betweenness_partial(self._h, p_u32(src), src.shape[0],
                                            int(include_endpoints), p_f64(out)))
        return out[: self.n_bound]

    def betweenness_source_order(self, sources):
        """The library's grouping order of ``sources`` (siblings of a hub next to each other)."""
        src = np.ascontiguousarray(sources, dtype=np.uint32)
        out = np.empty(max(src.shape[0], 1), dtype=np.uint32)
        check(lib().rxc_betweenness_source_order(self._h, p_u32(src), src.shape[0], p_u32(out)))
        return out[: src.shape[0]]

    def betweenness_partial_reduce(self, sources, include_endpoints=False, root=0, out=None):
        """This rank's shard; the partial vector is reduced in HBM over the ranks of ``rxc_comm_init``
        and returned on ``root`` only (``None`` elsewhere)."""
        src = np.ascontiguousarray(sources, dtype=np.uint32)
        if out is None:
            out = np.zeros(max(self.n_bound, 1), dtype=np.float64)
        check(lib().rxc_betweenness_partial_reduce(self._h, p_u32(src), src.shape[0],
                                                   int(include_endpoints), int(root), p_f64(out)))
        return out[: self.n_bound]

    def edge_betweenness_partial_reduce(self, sources, root=0):
        src = np.ascontiguousarray(sources, dtype=np.uint32)
        out = np.zeros(max(self.e_bound, 1), dtype=np.float64)
        check(lib().rxc_edge_betweenness_partial_reduce(self._h, p_u32(src), src.shape[0], int(root),
                                                        p_f64(out)))
        return out[: self.e_bound]

    def edge_betweenness_partial(self, sources):
        src = np.ascontiguousarray(sources, dtype=np.uint32)
        out = result_empty(max(self.e_bound, 1))  # the library writes all e_bound entries
        check(lib().rxc_edge_betweenness_partial(self._h, p_u32(src), src.shape[0], p_f64(out)))
        return out[: self.e_bound]

    def closeness_partial(self, sources, wf_improved=True):
        src = np.ascontiguousarray(sources, dtype=np.uint32)
        out = np.zeros(max(src.shape[0], 1), dtype=np.float64)
        check(lib().rxc_closeness_partial(self._h, p_u32(src), src.shape[0], int(wf_improved),
                                          p_f64(out)))
        return out[: src.shape[0]]

    def distance_matrix_rows(self, row_begin, row_end, as_undirected=False, null_value=0.0):
        out = result_empty((row_end - row_begin, self.n_live))
        if out.size:
            check(lib().rxc_distance_matrix_rows(self._h, int(as_undirected), float(null_value),
                                                 int(row_begin), int(row_end), p_f64(out)))
        return out


def rescale(x, node_count, normalized, directed, include_endpoints):
    x = np.ascontiguousarray(x, dtype=np.float64)
    lib().rxc_rescale(p_f64(x), x.shape[0], int(node_count), int(normalized), int(directed),
                      int(include_endpoints))
    return x
