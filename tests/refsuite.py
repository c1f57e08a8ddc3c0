"""Running the REFERENCE'S OWN Python tests against this package (missing #1).

The reference's test files import ``rustworkx``; here that name is aliased to ``rustworkx_b200`` and
the selected ``unittest`` classes are run unmodified from where they lie under
``/root/reference/tests`` (never copied into this repository).  Two back ends:

* **CPU container** (``/root/reference`` present, no GPU): the package's device layer is swapped for
  :class:`OracleDeviceGraph` -- the CPU oracle behind the same interface as ``_lib.DeviceGraph`` --
  so the reference's assertions exercise the host API (dispatch, argument checks, mappings, the
  stand-in containers, generators) AND pin the oracle on every value those tests assert.  Every call
  that reaches the device layer is recorded: the graph as it was at that moment, the operation, its
  arguments and the result the reference's assertions then accepted.
* **GPU box** (no ``/root/reference``): ``tests/golden/reference_suite_calls.json`` -- that recording,
  committed together with ``tests/golden/make_reference_suite_calls.py`` -- is replayed through the
  real C-ABI on the GPU and compared with the recorded results (distance work bit-exact,
  betweenness within 1e-9).  When the reference checkout *is* present next to a GPU the test classes
  themselves run against the CUDA path as well.

Test infrastructure only: nothing under ``rustworkx_b200/`` imports this module.
"""

from __future__ import annotations

import importlib.util
import io
import json
import os
import sys
import unittest

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE_TESTS = os.environ.get("RX_REFERENCE_TESTS", "/root/reference/tests")
GOLDEN = os.path.join(ROOT, "tests", "golden", "reference_suite_calls.json")

# reference test file -> test classes (and, optionally, the methods) that are on the SURVEY 8 path
SELECTION = {
    "graph/test_centrality.py": {
        "TestCentralityGraph": None,
        "TestCentralityGraphDeletedNode": None,
        "TestEdgeBetweennessCentrality": None,
        "TestGroupClosenessCentralityGraph": None,
        "TestGroupBetweennessCentralityGraph": None,
        "TestGroupCentralityExpectedValuesGraph": ("closeness", "betweenness"),
    },
    "digraph/test_centrality.py": {
        "TestCentralityDiGraph": None,
        "TestCentralityDiGraphDeletedNode": None,
        "TestEdgeBetweennessCentrality": None,
        "TestGroupClosenessCentralityDiGraph": None,
        "TestGroupBetweennessCentralityDiGraph": None,
        "TestGroupCentralityExpectedValuesDiGraph": ("closeness", "betweenness"),
    },
    "graph/test_dist_matrix.py": None,       # every class
    "digraph/test_dist_matrix.py": None,
    "graph/test_avg_shortest_path.py": None,
    "digraph/test_avg_shortest_path.py": None,
    "test_dispatch.py": {
        "TestDispatchPyGraph": ("test_distance_matrix", "test_distance_matrix_as_undirected",
                                "test_betweenness_centrality", "test_all_pairs_dijkstra_path_lengths"),
        "TestDispatchPyDiGraph": ("test_distance_matrix", "test_distance_matrix_as_undirected",
                                  "test_betweenness_centrality", "test_all_pairs_dijkstra_path_lengths"),
    },
}


def reference_available():
    return os.path.isdir(REFERENCE_TESTS) and os.path.isfile(os.path.join(REFERENCE_TESTS, "test_dispatch.py"))


# --------------------------------------------------------------------------------------------------
# the oracle behind the DeviceGraph interface (+ recorder)
# --------------------------------------------------------------------------------------------------
def _oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle

    return oracle


class Recorder:
    def __init__(self):
        self.graphs = []   # distinct graph snapshots
        self._index = {}
        self.calls = []    # {"graph": i, "op": ..., "args": [...], "result": ...}

    def graph_id(self, snap):
        key = json.dumps(snap, sort_keys=True)
        if key not in self._index:
            self._index[key] = len(self.graphs)
            self.graphs.append(snap)
        return self._index[key]

    def add(self, gid, op, args, result):
        self.calls.append({"graph": gid, "op": op, "args": args, "result": encode(result)})

    def dump(self, path):
        with open(path, "w") as f:
            json.dump({"graphs": self.graphs, "calls": self.calls}, f, separators=(",", ":"))
            f.write("\n")


def encode(x):
    """JSON-safe, loss-free: floats travel as their hex representation (NaN / inf included)."""
    if isinstance(x, np.ndarray):
        return {"shape": list(x.shape), "hex": [float(v).hex() for v in x.reshape(-1)]}
    if isinstance(x, (tuple, list)):
        return {"list": [encode(v) for v in x]}
    if isinstance(x, (float, np.floating)):
        return {"f": float(x).hex()}
    if isinstance(x, (int, np.integer)):
        return {"i": int(x)}
    raise TypeError(type(x))


def decode(x):
    if "hex" in x:
        return np.array([float.fromhex(v) for v in x["hex"]], dtype=np.float64).reshape(x["shape"])
    if "list" in x:
        return [decode(v) for v in x["list"]]
    if "f" in x:
        return float.fromhex(x["f"])
    return x["i"]


RECORDER = None  # set by the generating script / the CPU test to capture calls


class OracleDeviceGraph:
    """``_lib.DeviceGraph``'s interface answered by the CPU oracle (tests only)."""

    def __init__(self, graph):
        oracle = _oracle()
        s = graph._snapshot_arrays()
        self.live_nodes = np.ascontiguousarray(s["live_nodes"], dtype=np.uint32)
        self.edge_id = np.ascontiguousarray(s["edge_id"], dtype=np.uint32)
        self.n_bound, self.e_bound = int(s["n_bound"]), int(s["e_bound"])
        self.n_live = int(self.live_nodes.shape[0])
        self.directed = bool(s["directed"])
        self._og = oracle.OracleGraph.from_graph(graph)
        self._gid = None
        if RECORDER is not None:
            self._gid = RECORDER.graph_id({
                "directed": self.directed, "n_bound": self.n_bound, "e_bound": self.e_bound,
                "live_nodes": self.live_nodes.tolist(), "edge_src": np.asarray(s["edge_src"]).tolist(),
                "edge_dst": np.asarray(s["edge_dst"]).tolist(), "edge_id": self.edge_id.tolist()})

    @classmethod
    def from_graph(cls, graph, cached=False):
        return cls(graph)

    def _rec(self, op, args, result):
        if RECORDER is not None:
            RECORDER.add(self._gid, op, args, result)
        return result

    @staticmethod
    def _dense(pair):
        vals, present = pair
        return np.where(present.astype(bool), vals, 0.0)

    def betweenness(self, include_endpoints, normalized):
        out = self._dense(self._og.betweenness_centrality(include_endpoints, normalized, 50, as_arrays=True))
        return self._rec("betweenness", [bool(include_endpoints), bool(normalized)], out)

    def edge_betweenness(self, normalized):
        out = self._dense(self._og.edge_betweenness_centrality(normalized, 50, as_arrays=True))
        return self._rec("edge_betweenness", [bool(normalized)], out)

    def closeness(self, wf_improved):
        out = self._dense(self._og.closeness_centrality(wf_improved, 50, as_arrays=True))
        return self._rec("closeness", [bool(wf_improved)], out)

    def distance_matrix(self, as_undirected, null_value):
        out = self._og.distance_matrix(300, as_undirected, null_value)
        return self._rec("distance_matrix", [bool(as_undirected), float(null_value).hex()], out)

    def distance_sums(self, as_undirected):
        total, pairs = self._og.distance_sum(300, as_undirected)
        self._rec("distance_sums", [bool(as_undirected)], [total, pairs])
        return total, pairs

    def group_betweenness(self, group, normalized):
        out = self._og.group_betweenness_centrality(np.asarray(group, dtype=np.uint32), normalized, 50)
        return self._rec("group_betweenness", [[int(x) for x in group], bool(normalized)], out)

    def group_closeness(self, group):
        out = self._og.group_closeness_centrality(np.asarray(group, dtype=np.uint32))
        return self._rec("group_closeness", [[int(x) for x in group]], out)

    def newman_weighted_closeness(self, edge_weight, wf_improved):
        w = np.asarray(edge_weight, dtype=np.float64)
        out = self._dense(self._og.newman_weighted_closeness_centrality(w, wf_improved, 50, as_arrays=True))
        return self._rec("newman_weighted_closeness", [[float(x).hex() for x in w], bool(wf_improved)], out)

    def all_pairs_dijkstra_lengths_rows(self, edge_weight, row_begin=0, row_end=None,
                                        unreachable_value=float("nan")):
        w = np.asarray(edge_weight, dtype=np.float64)
        row_end = self.n_live if row_end is None else row_end
        rows = self._og.dijkstra_lengths_rows(w, self.live_nodes[row_begin:row_end])
        out = np.ascontiguousarray(rows[:, self.live_nodes.astype(np.int64)])
        return self._rec("all_pairs_dijkstra_lengths_rows", [[float(x).hex() for x in w]], out)


# --------------------------------------------------------------------------------------------------
# loading and running the reference's test classes
# --------------------------------------------------------------------------------------------------
def _alias_rustworkx():
    import rustworkx_b200

    saved = {k: sys.modules.get(k) for k in ("rustworkx", "rustworkx.generators")}
    sys.modules["rustworkx"] = rustworkx_b200
    sys.modules["rustworkx.generators"] = rustworkx_b200.generators
    return saved


def _restore(saved):
    for k, v in saved.items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v


def load_reference_suite():
    """unittest.TestSuite of the selected classes / methods, loaded from REFERENCE_TESTS."""
    loader = unittest.TestLoader()
    suite = unittest.TestSuite()
    picked = []
    for rel, classes in SELECTION.items():
        path = os.path.join(REFERENCE_TESTS, rel)
        name = "rxref_" + rel.replace("/", "_").replace(".py", "")
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        for attr in sorted(dir(mod)):
            obj = getattr(mod, attr)
            if not (isinstance(obj, type) and issubclass(obj, unittest.TestCase)):
                continue
            if classes is not None and attr not in classes:
                continue
            want = None if classes is None else classes[attr]
            for method in loader.getTestCaseNames(obj):
                if want is not None and not any(method == w or w in method.split("_") for w in want):
                    continue
                suite.addTest(obj(method))
                picked.append(f"{rel}::{attr}::{method}")
    return suite, picked


def run_reference_suite(oracle_backend, recorder=None):
    """Run the selected reference tests against rustworkx_b200; returns (result, names, report)."""
    global RECORDER
    import rustworkx_b200

    saved_modules = _alias_rustworkx()
    saved_dg = rustworkx_b200.DeviceGraph
    saved_cache = rustworkx_b200._SNAPSHOT_CACHE
    RECORDER = recorder
    try:
        if oracle_backend:
            rustworkx_b200.DeviceGraph = OracleDeviceGraph
            rustworkx_b200.set_snapshot_cache(False)  # record every call with the graph of that moment
        suite, picked = load_reference_suite()
        stream = io.StringIO()
        result = unittest.TextTestRunner(stream=stream, verbosity=1).run(suite)
        return result, picked, stream.getvalue()
    finally:
        RECORDER = None
        rustworkx_b200.DeviceGraph = saved_dg
        rustworkx_b200.set_snapshot_cache(saved_cache)
        _restore(saved_modules)


# --------------------------------------------------------------------------------------------------
# replay of the recording through the real C-ABI (GPU)
# --------------------------------------------------------------------------------------------------
def replay(path=GOLDEN, rtol=1e-9):
    """Re-issue every recorded call on the GPU; returns (number of calls, list of mismatches)."""
    from rustworkx_b200 import _lib

    with open(path) as f:
        rec = json.load(f)
    handles = {}
    bad = []

    def handle(i):
        if i not in handles:
            s = rec["graphs"][i]
            handles[i] = _lib.DeviceGraph(s["n_bound"], np.asarray(s["live_nodes"], dtype=np.uint32),
                                          s["e_bound"], np.asarray(s["edge_src"], dtype=np.uint32),
                                          np.asarray(s["edge_dst"], dtype=np.uint32),
                                          np.asarray(s["edge_id"], dtype=np.uint32), s["directed"])
        return handles[i]

    exact_ops = {"closeness", "distance_matrix", "distance_sums", "group_closeness",
                 "all_pairs_dijkstra_lengths_rows"}
    for k, call in enumerate(rec["calls"]):
        dg, op, args = handle(call["graph"]), call["op"], call["args"]
        want = decode(call["result"])
        if op == "distance_matrix":
            got = dg.distance_matrix(args[0], float.fromhex(args[1]))
        elif op == "newman_weighted_closeness":
            got = dg.newman_weighted_closeness(np.array([float.fromhex(x) for x in args[0]]), args[1])
        elif op == "all_pairs_dijkstra_lengths_rows":
            got = dg.all_pairs_dijkstra_lengths_rows(np.array([float.fromhex(x) for x in args[0]]))
        elif op in ("group_betweenness", "group_closeness"):
            got = getattr(dg, op)(np.asarray(args[0], dtype=np.uint32), *args[1:])
        elif op == "distance_sums":
            got = list(dg.distance_sums(*args))
        else:
            got = getattr(dg, op)(*args)
        got_a, want_a = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
        if got_a.shape != want_a.shape:
            bad.append((k, op, "shape", got_a.shape, want_a.shape))
        elif op in exact_ops:
            if not np.array_equal(got_a, want_a, equal_nan=True):
                bad.append((k, op, "not bit-exact"))
        else:
            err = np.abs(got_a - want_a)
            tol = rtol * np.abs(want_a) + 1e-12
            if not np.all((err <= tol) | (np.isnan(got_a) & np.isnan(want_a))):
                bad.append((k, op, "max err %.3e" % float(np.nanmax(err))))
    for h in handles.values():
        h.close()
    return len(rec["calls"]), bad
