"""Golden cases of the reference for the group centralities, shared by the oracle pin
(test_oracle_golden.py) and the GPU parity tests (test_gpu_parity.py).

Sources: rustworkx-core/src/centrality.rs:2155-2360 (unit tests, 1e-10), tests/graph/test_centrality.py:
341-567 and tests/digraph/test_centrality.py:374-565 (assertAlmostEqual, places 7 / 10; the expected
values were obtained by the reference's authors with NetworkX 3.6.1).
Each case: (name, directed, n_nodes, edges, group, expected).
"""


def path(n):
    return [(i, i + 1) for i in range(n - 1)]


def cycle(n):
    return [(i, (i + 1) % n) for i in range(n)]


def complete(n, directed=False):
    if directed:
        return [(i, j) for i in range(n) for j in range(n) if i != j]
    return [(i, j) for i in range(n) for j in range(i + 1, n)]


def star(n):
    return [(0, i) for i in range(1, n)]


def both(edges):
    return [e for a, b in edges for e in ((a, b), (b, a))]


BARBELL = (complete(4) + [(i, j) for i in range(5, 9) for j in range(i + 1, 9)] + [(3, 4), (4, 5)])

CLOSENESS = [
    ("core_path", False, 5, path(5), [0, 1], 0.5),
    ("core_center", False, 5, [(0, 2), (1, 2), (2, 3), (2, 4)], [2], 1.0),
    ("core_disconnected", False, 3, [(0, 1)], [0], 2.0),
    ("core_duplicates", False, 3, path(3), [1, 1], 1.0),
    ("py_star_center", False, 5, star(5), [0], 1.0),
    ("py_path_0", False, 5, path(5), [0], 0.4),
    ("py_path_2", False, 5, path(5), [2], 2 / 3),
    ("py_path_13", False, 5, path(5), [1, 3], 1.0),
    ("py_path_024", False, 5, path(5), [0, 2, 4], 1.0),
    ("py_k6_0", False, 6, complete(6), [0], 1.0),
    ("py_k6_01", False, 6, complete(6), [0, 1], 1.0),
    ("py_k6_024", False, 6, complete(6), [0, 2, 4], 1.0),
    ("py_c8_0", False, 8, cycle(8), [0], 0.4375),
    ("py_c8_04", False, 8, cycle(8), [0, 4], 0.75),
    ("py_c8_0246", False, 8, cycle(8), [0, 2, 4, 6], 1.0),
    ("py_dpath4_3", True, 4, path(4), [3], 0.5),
    ("py_dpath6_0", True, 6, path(6), [0], 0.0),
    ("py_dpath6_2", True, 6, path(6), [2], 5 / 3),
    ("py_dpath6_01", True, 6, path(6), [0, 1], 0.0),
    ("py_dpath6_13", True, 6, path(6), [1, 3], 2.0),
    ("py_dcycle6_0", True, 6, cycle(6), [0], 1 / 3),
    ("py_dcycle6_03", True, 6, cycle(6), [0, 3], 2 / 3),
    ("py_dcycle6_024", True, 6, cycle(6), [0, 2, 4], 1.0),
]

# (name, directed, n_nodes, edges, group, normalized, expected)
BETWEENNESS = [
    ("core_path_center", False, 5, path(5), [2], False, 4.0),
    ("core_path_center_norm", False, 5, path(5), [2], True, 2 / 3),
    ("core_empty", False, 3, path(3), [], False, 0.0),
    ("core_star", False, 5, star(5), [0], False, 6.0),
    ("core_dpath", True, 5, path(5), [2], False, 4.0),
    ("core_dpath_norm", True, 5, path(5), [2], True, 1 / 3),
    ("core_dstar", True, 5, star(5), [0], False, 0.0),
    ("core_bistar", True, 5, both(star(5)), [0], False, 12.0),
    ("core_duplicates", False, 5, path(5), [2, 2], True, 2 / 3),
    ("py_path_13", False, 5, path(5), [1, 3], True, 1.0),
    ("py_path_04", False, 5, path(5), [0, 4], True, 0.0),
    ("py_k6_0", False, 6, complete(6), [0], True, 0.0),
    ("py_k6_01", False, 6, complete(6), [0, 1], True, 0.0),
    ("py_k6_024", False, 6, complete(6), [0, 2, 4], True, 0.0),
    ("py_star_0", False, 5, star(5), [0], True, 1.0),
    ("py_star_1", False, 5, star(5), [1], True, 0.0),
    ("py_star_01", False, 5, star(5), [0, 1], True, 1.0),
    ("py_barbell_4", False, 9, BARBELL, [4], True, 4 / 7),
    ("py_barbell_345", False, 9, BARBELL, [3, 4, 5], True, 0.6),
    ("py_barbell_08", False, 9, BARBELL, [0, 8], True, 0.0),
    ("py_bipath_2", True, 6, both(path(6)), [2], True, 0.6),
    ("py_bipath_14", True, 6, both(path(6)), [1, 4], True, 5 / 6),
    ("py_bipath_05", True, 6, both(path(6)), [0, 5], True, 0.0),
    ("py_bistar_0", True, 5, both(star(5)), [0], True, 1.0),
    ("py_bistar_12", True, 5, both(star(5)), [1, 2], True, 0.0),
    ("py_dcycle_0", True, 6, cycle(6), [0], True, 0.5),
    ("py_dcycle_03", True, 6, cycle(6), [0, 3], True, 5 / 6),
    ("py_dk5_0", True, 5, complete(5, True), [0], True, 0.0),
    ("py_dk5_01", True, 5, complete(5, True), [0, 1], True, 0.0),
    ("py_dk5_024", True, 5, complete(5, True), [0, 2, 4], True, 0.0),
]
