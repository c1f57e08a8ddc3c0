"""Golden cases of the reference for the weighted shortest-path functions, shared by
the oracle pin (test_oracle_golden.py) and the GPU parity tests (test_gpu_parity.py).

Newman closeness: rustworkx-core/src/centrality.rs:1382-1760 (assert_almost_equal 1e-4; unit-weight
cases assert_eq against closeness_centrality).  Each case: (name, directed, edges (a, b, weight),
wf_improved, expected list over node index).
All-pairs Dijkstra: tests/graph/test_dijkstra.py:20-36,129-176 and tests/digraph/test_dijkstra.py:
20-39,198-258 (assertEqual).
"""

PATH7 = [(i, i + 1, 1.0) for i in range(6)]
TWO_CC = [(0, 1, 1.0), (1, 2, 0.5), (2, 3, 0.25), (4, 5, 1.0), (5, 6, 0.5), (6, 7, 0.25)]
MANY_TO_ONE = [(i, 0, 0.1) for i in range(1, 8)] + [(0, 8, 1.0)]

NEWMAN = [
    ("two_cc_digraph", True, TWO_CC, False,
     [0.0, 1.0, 0.4, 0.176470, 0.0, 1.0, 0.4, 0.176470]),
    ("two_cc_improved_digraph", True, TWO_CC, True,
     [0.0, 0.14285714, 0.11428571, 0.07563025, 0.0, 0.14285714, 0.11428571, 0.07563025]),
    ("two_cc_improved_cardinality_digraph", True, TWO_CC + [(7, 8, 0.125)], True,
     [0.0, 0.125, 0.1, 0.06617647, 0.0, 0.125, 0.1, 0.06617647, 0.04081632]),
    ("small_ungraph", False, [(0, 1, 0.7), (1, 2, 0.2), (2, 3, 0.5)], False,
     [0.1842105, 0.2234042, 0.2234042, 0.1721311]),
    ("small_digraph", True, [(0, 1, 0.7), (1, 2, 0.2), (2, 3, 0.5)], False,
     [0.0, 0.7, 0.175, 0.172131]),
    ("many_to_one_digraph", True, MANY_TO_ONE, False,
     [0.1, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.10256]),
    ("many_to_one_ungraph", False, MANY_TO_ONE, False,
     [0.112676056] + [0.056737588] * 7 + [0.102564102]),
]

# unit weights: identical (==) to closeness_centrality, both wf settings
NEWMAN_UNIT = [("path7_ungraph", False, PATH7), ("path7_digraph", True, PATH7)]

DIJKSTRA_EDGES = [(0, 1, 7), (2, 0, 9), (0, 3, 14), (1, 2, 10), (3, 2, 2), (3, 4, 9), (1, 5, 15),
                  (2, 5, 11), (4, 5, 6)]

ALL_PAIRS = [
    ("graph", False, None, {
        0: {1: 7.0, 2: 9.0, 3: 11.0, 4: 20.0, 5: 20.0},
        1: {0: 7.0, 2: 10.0, 3: 12.0, 4: 21.0, 5: 15.0},
        2: {0: 9.0, 1: 10.0, 3: 2.0, 4: 11.0, 5: 11.0},
        3: {0: 11.0, 1: 12.0, 2: 2.0, 4: 9.0, 5: 13.0},
        4: {0: 20.0, 1: 21.0, 2: 11.0, 3: 9.0, 5: 6.0},
        5: {0: 20.0, 1: 15.0, 2: 11.0, 3: 13.0, 4: 6.0},
    }),
    ("graph_node_removed", False, 3, {
        0: {1: 7.0, 2: 9.0, 4: 26.0, 5: 20.0},
        1: {0: 7.0, 2: 10.0, 4: 21.0, 5: 15.0},
        2: {0: 9.0, 1: 10.0, 4: 17.0, 5: 11.0},
        4: {0: 26.0, 1: 21.0, 2: 17.0, 5: 6.0},
        5: {0: 20.0, 1: 15.0, 2: 11.0, 4: 6.0},
    }),
    ("digraph", True, None, {
        0: {1: 7.0, 2: 16.0, 3: 14.0, 4: 23.0, 5: 22.0},
        1: {0: 19.0, 2: 10.0, 3: 33.0, 4: 42.0, 5: 15.0},
        2: {0: 9.0, 1: 16.0, 3: 23.0, 4: 32.0, 5: 11.0},
        3: {0: 11.0, 1: 18.0, 2: 2.0, 4: 9.0, 5: 13.0},
        4: {5: 6.0},
        5: {},
    }),
    ("digraph_node_removed", True, 3, {
        0: {1: 7.0, 2: 17.0, 5: 22.0},
        1: {0: 19.0, 2: 10.0, 5: 15.0},
        2: {0: 9.0, 1: 16.0, 5: 11.0},
        4: {5: 6.0},
        5: {},
    }),
]
