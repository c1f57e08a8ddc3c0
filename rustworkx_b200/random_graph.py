"""Seeded random graph generators named by BASELINE.json's configs (benchmark inputs only).

Mirrors the Python names of the reference (``src/random_graph.rs:182,240,626``); the edge sets come
from ``librxgen.so``'s host-side restatement of the algorithms in
``rustworkx-core/src/generators/random_graph.rs`` (see ``csrc/generators.cpp``).  The reference's
RNG stream (rand_pcg / rand, un-vendored) cannot be reproduced here, so a given seed yields a graph
from the same distribution, not the same graph.
"""

from __future__ import annotations

import numpy as np

from . import _lib
from .graph import PyDiGraph, PyGraph


def _seed(seed):
    if seed is None:
        return int(np.random.SeedSequence().generate_state(1, dtype=np.uint64)[0])
    if isinstance(seed, (bool, np.bool_)) or not isinstance(seed, (int, np.integer)):
        raise TypeError(f"'{type(seed).__name__}' object cannot be interpreted as an integer")
    if seed < 0:
        raise OverflowError("can't convert negative int to unsigned")
    return int(seed)


def _make(directed, n, src, dst, count):
    g = PyDiGraph() if directed else PyGraph()
    g._bulk_add_nodes(int(n))
    g._bulk_add_edges(src[:count], dst[:count])
    return g


def _gnp(num_nodes, probability, seed, directed):
    n = int(num_nodes)
    if n <= 0:
        raise ValueError("num_nodes must be > 0")
    p = float(probability)
    if not (0.0 <= p <= 1.0):
        raise ValueError("Probability out of range, must be 0 <= p <= 1")
    pairs = n * (n - 1) if directed else n * (n - 1) // 2
    mean = pairs * p
    cap = int(min(pairs, mean + 8.0 * np.sqrt(mean + 1.0) + 64))
    L = _lib.gen_lib()
    while True:
        src = np.empty(max(cap, 1), dtype=np.uint32)
        dst = np.empty(max(cap, 1), dtype=np.uint32)
        m = L.rxc_gen_gnp(n, p, _seed(seed), int(directed), _lib.p_u32(src), _lib.p_u32(dst), cap)
        if m >= 0:
            return _make(directed, n, src, dst, int(m))
        if cap >= pairs:
            raise ValueError("gnp generation failed")
        cap = min(pairs, cap * 2)


def undirected_gnp_random_graph(num_nodes, probability, seed=None):
    return _gnp(num_nodes, probability, seed, False)


def directed_gnp_random_graph(num_nodes, probability, seed=None):
    return _gnp(num_nodes, probability, seed, True)


def _gnm(num_nodes, num_edges, seed, directed):
    n, m = int(num_nodes), int(num_edges)
    if n <= 0:
        raise ValueError("num_nodes must be > 0")
    if m < 0:
        raise OverflowError("can't convert negative int to unsigned")
    pairs = n * (n - 1) if directed else n * (n - 1) // 2
    cap = min(m, pairs)
    src = np.empty(max(cap, 1), dtype=np.uint32)
    dst = np.empty(max(cap, 1), dtype=np.uint32)
    got = _lib.gen_lib().rxc_gen_gnm(n, m, _seed(seed), int(directed), _lib.p_u32(src), _lib.p_u32(dst), cap)
    if got < 0:
        raise ValueError("gnm generation failed")
    return _make(directed, n, src, dst, int(got))


def undirected_gnm_random_graph(num_nodes, num_edges, seed=None):
    return _gnm(num_nodes, num_edges, seed, False)


def directed_gnm_random_graph(num_nodes, num_edges, seed=None):
    return _gnm(num_nodes, num_edges, seed, True)


def barabasi_albert_graph(n, m, seed=None, initial_graph=None):
    """Undirected Barabasi-Albert graph grown from a star of ``m`` nodes."""
    if initial_graph is not None:
        raise NotImplementedError("initial_graph is not supported by this benchmark generator")
    n, m = int(n), int(m)
    if m < 1 or m >= n:
        raise ValueError("m must be >= 1 and < n")
    cap = (m - 1) + m * (n - m)
    src = np.empty(max(cap, 1), dtype=np.uint32)
    dst = np.empty(max(cap, 1), dtype=np.uint32)
    got = _lib.gen_lib().rxc_gen_barabasi_albert(n, m, _seed(seed), _lib.p_u32(src), _lib.p_u32(dst), cap)
    if got < 0:
        raise ValueError("barabasi_albert generation failed")
    return _make(False, n, src, dst, int(got))
