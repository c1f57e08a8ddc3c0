"""Benchmark / test input generators (mirror of the ``rustworkx.generators`` names this path uses).

Generators are OUT OF SCOPE as a product; these restate just the ones the
reference's hot-path tests and BASELINE.json's configs name, with the same node numbering and edge
insertion order as the reference so edge indices line up:

* ``path_graph`` / ``directed_path_graph``   -- rustworkx-core/src/generators/path_graph.rs:93-99
* ``cycle_graph`` / ``directed_cycle_graph`` -- .../cycle_graph.rs:93-106
* ``mesh_graph`` / ``complete_graph`` (+directed) -- .../complete_graph.rs:83-90
* ``star_graph``                             -- .../star_graph.rs:98-108
* ``full_rary_tree``                         -- .../full_rary_tree_graph.rs:104-120
* ``grid_graph`` / ``directed_grid_graph``   -- .../grid_graph.rs:64-147
* ``heavy_hex_graph``                        -- .../heavy_hex_graph.rs:107-257

Node payloads are the node index (as the reference's Python wrappers do, ``src/generators.rs``); edge
payloads are ``None``.
"""

from __future__ import annotations

import numpy as np

from .graph import PyDiGraph, PyGraph


def _build(directed, n, src, dst, multigraph=True):
    g = PyDiGraph(multigraph=multigraph) if directed else PyGraph(multigraph=multigraph)
    g._bulk_add_nodes(int(n))
    g._bulk_add_edges(np.asarray(src, dtype=np.uint32), np.asarray(dst, dtype=np.uint32))
    return g


def _interleave(a_src, a_dst):
    """(a,b) followed immediately by (b,a) for every edge (the ``bidirectional`` pattern)."""
    src = np.empty(2 * len(a_src), dtype=np.uint32)
    dst = np.empty(2 * len(a_src), dtype=np.uint32)
    src[0::2], dst[0::2] = a_src, a_dst
    src[1::2], dst[1::2] = a_dst, a_src
    return src, dst


def _num_nodes(num_nodes, weights):
    if num_nodes is None and weights is None:
        raise IndexError("num_nodes and weights list not specified")
    return len(weights) if num_nodes is None else int(num_nodes)


def path_graph(num_nodes=None, weights=None, multigraph=True):
    n = _num_nodes(num_nodes, weights)
    a = np.arange(max(n - 1, 0), dtype=np.uint32)
    return _build(False, n, a, a + 1, multigraph)


def directed_path_graph(num_nodes=None, weights=None, bidirectional=False, multigraph=True):
    n = _num_nodes(num_nodes, weights)
    a = np.arange(max(n - 1, 0), dtype=np.uint32)
    src, dst = (a, a + 1) if not bidirectional else _interleave(a, a + 1)
    return _build(True, n, src, dst, multigraph)


def _cycle_edges(n):
    if n == 0:
        return np.zeros(0, np.uint32), np.zeros(0, np.uint32)
    a = np.arange(n, dtype=np.uint32)
    return a, (a + 1) % np.uint32(n)


def cycle_graph(num_nodes=None, weights=None, multigraph=True):
    n = _num_nodes(num_nodes, weights)
    src, dst = _cycle_edges(n)
    return _build(False, n, src, dst, multigraph)


def directed_cycle_graph(num_nodes=None, weights=None, bidirectional=False, multigraph=True):
    n = _num_nodes(num_nodes, weights)
    src, dst = _cycle_edges(n)
    if bidirectional:
        src, dst = _interleave(src, dst)
    return _build(True, n, src, dst, multigraph)


def _complete_edges(n):
    i, j = np.triu_indices(n, k=1)
    return i.astype(np.uint32), j.astype(np.uint32)


def mesh_graph(num_nodes=None, weights=None, multigraph=True):
    n = _num_nodes(num_nodes, weights)
    src, dst = _complete_edges(n)
    return _build(False, n, src, dst, multigraph)


complete_graph = mesh_graph


def directed_mesh_graph(num_nodes=None, weights=None, multigraph=True):
    n = _num_nodes(num_nodes, weights)
    src, dst = _interleave(*_complete_edges(n))
    return _build(True, n, src, dst, multigraph)


directed_complete_graph = directed_mesh_graph


def star_graph(num_nodes=None, weights=None, multigraph=True):
    n = _num_nodes(num_nodes, weights)
    a = np.arange(1, max(n, 1), dtype=np.uint32)
    return _build(False, n, np.zeros_like(a), a, multigraph)


def directed_star_graph(
    num_nodes=None, weights=None, inward=False, bidirectional=False, multigraph=True
):
    n = _num_nodes(num_nodes, weights)
    a = np.arange(1, max(n, 1), dtype=np.uint32)
    z = np.zeros_like(a)
    if bidirectional:
        src, dst = _interleave(a, z)
    elif inward:
        src, dst = a, z
    else:
        src, dst = z, a
    return _build(True, n, src, dst, multigraph)


def full_rary_tree(branching_factor, num_nodes, weights=None, multigraph=True):
    n = int(num_nodes)
    r = int(branching_factor)
    child = np.arange(1, max(n, 1), dtype=np.uint32)
    parent = ((child - 1) // max(r, 1)).astype(np.uint32) if r > 0 else child[:0]
    if r == 0:
        child = child[:0]
    return _build(False, n, parent, child, multigraph)


def _grid_edges(rows, cols):
    idx = np.arange(rows * cols, dtype=np.int64).reshape(rows, cols)
    # per cell (i, j): first the "down" edge, then the "right" edge (grid_graph.rs:124-145)
    down_ok = np.zeros((rows, cols), dtype=bool)
    down_ok[: rows - 1, :] = True
    right_ok = np.zeros((rows, cols), dtype=bool)
    right_ok[:, : cols - 1] = True
    src = np.stack([idx, idx], axis=-1)
    dst = np.stack([idx + cols, idx + 1], axis=-1)
    ok = np.stack([down_ok, right_ok], axis=-1)
    return src[ok].astype(np.uint32), dst[ok].astype(np.uint32)


def grid_graph(rows=None, cols=None, weights=None, multigraph=True):
    if weights is None and (rows is None or cols is None):
        raise IndexError("dimensions and weights list not specified")
    rows, cols = int(rows or 0), int(cols or 0)
    if weights is not None and rows * cols < len(weights) and rows == 0:
        rows, cols = 1, len(weights)
    src, dst = _grid_edges(rows, cols) if rows * cols else (np.zeros(0), np.zeros(0))
    return _build(False, rows * cols, src, dst, multigraph)


def directed_grid_graph(rows=None, cols=None, weights=None, bidirectional=False, multigraph=True):
    if weights is None and (rows is None or cols is None):
        raise IndexError("dimensions and weights list not specified")
    rows, cols = int(rows or 0), int(cols or 0)
    src, dst = _grid_edges(rows, cols) if rows * cols else (np.zeros(0), np.zeros(0))
    if bidirectional:
        src, dst = _interleave(src, dst)
    return _build(True, rows * cols, src, dst, multigraph)


def _heavy_hex_edges(d):
    num_data = d * d
    half = (d + 1) // 2
    num_syndrome = (d - 1) * half
    num_flag = d * (d - 1)
    data = np.arange(num_data)
    syn = num_data + np.arange(num_syndrome)
    flag = num_data + num_syndrome + np.arange(num_flag)
    edges = []
    # connect data and flags (heavy_hex_graph.rs:141-151)
    for i in range(d):
        for j in range(d - 1):
            f = flag[i * (d - 1) + j]
            edges.append((data[i * d + j], f))
            edges.append((data[i * d + j + 1], f))
    # connect data and syndromes (:153-191)
    for i in range(d - 1):
        chunk = syn[i * half : (i + 1) * half]
        if i % 2 == 0:
            edges.append((data[i * d], chunk[0]))
            edges.append((data[(i + 1) * d], chunk[0]))
        else:
            edges.append((data[i * d + (d - 1)], chunk[-1]))
            edges.append((data[i * d + (2 * d - 1)], chunk[-1]))
    # connect flags and syndromes (:193-255)
    for i in range(d - 1):
        chunk = syn[i * half : (i + 1) * half]
        if i % 2 == 0:
            for j, s in enumerate(chunk):
                if j != 0:
                    edges.append((s, flag[i * (d - 1) + 2 * (j - 1) + 1]))
                    edges.append((s, flag[(i + 1) * (d - 1) + 2 * (j - 1) + 1]))
        else:
            for j, s in enumerate(chunk):
                if j != len(chunk) - 1:
                    edges.append((s, flag[i * (d - 1) + 2 * j]))
                    edges.append((s, flag[(i + 1) * (d - 1) + 2 * j]))
    arr = np.asarray(edges, dtype=np.uint32).reshape(-1, 2)
    return (5 * d * d - 2 * d - 1) // 2, arr[:, 0], arr[:, 1]


def heavy_hex_graph(d, multigraph=True):
    d = int(d)
    if d % 2 == 0:
        raise IndexError("d must be an odd number.")
    if d == 1:
        return _build(False, 1, np.zeros(0), np.zeros(0), multigraph)
    n, src, dst = _heavy_hex_edges(d)
    return _build(False, n, src, dst, multigraph)


def directed_heavy_hex_graph(d, bidirectional=False, multigraph=True):
    d = int(d)
    if d % 2 == 0:
        raise IndexError("d must be an odd number.")
    if d == 1:
        return _build(True, 1, np.zeros(0), np.zeros(0), multigraph)
    n, src, dst = _heavy_hex_edges(d)
    if bidirectional:
        # (a, f) (b, f) then (f, a) (f, b): pairs of forward edges followed by their reverses
        s4 = src.reshape(-1, 2)
        d4 = dst.reshape(-1, 2)
        src = np.concatenate([s4, d4], axis=1).reshape(-1)
        dst = np.concatenate([d4, s4], axis=1).reshape(-1)
    return _build(True, n, src, dst, multigraph)
