"""The reference's own Python tests, run against this package (VERDICT r1 missing #1).

See tests/refsuite.py for how: ``rustworkx`` is aliased to ``rustworkx_b200`` and the reference's
unittest classes for the hot path run unmodified from /root/reference/tests.
"""

import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import refsuite  # noqa: E402


def test_reference_tests_pass_on_the_host_api_with_the_oracle_behind_it():
    """CPU: the reference's assertions pin the host API (dispatchers, typed functions, argument
    errors, mappings, stand-in containers, generators) and the oracle's values."""
    if not refsuite.reference_available():
        pytest.skip("/root/reference is not present (GPU box): the recording is replayed instead")
    rec = refsuite.Recorder()
    result, picked, report = refsuite.run_reference_suite(True, rec)
    assert len(picked) >= 110, picked
    assert result.wasSuccessful() and not result.skipped, report
    # the committed recording is what this run produces (the script that writes it does the same)
    import json

    with open(refsuite.GOLDEN) as f:
        golden = json.load(f)
    assert golden["graphs"] == rec.graphs
    assert golden["calls"] == json.loads(json.dumps(rec.calls))


def test_recording_covers_every_hot_path_operation():
    import json

    with open(refsuite.GOLDEN) as f:
        golden = json.load(f)
    ops = {c["op"] for c in golden["calls"]}
    assert {"betweenness", "edge_betweenness", "closeness", "distance_matrix", "distance_sums",
            "group_betweenness", "group_closeness", "newman_weighted_closeness",
            "all_pairs_dijkstra_lengths_rows"} <= ops
    assert any(g["n_bound"] != len(g["live_nodes"]) for g in golden["graphs"])  # node holes
    assert any(g["directed"] for g in golden["graphs"]) and any(not g["directed"] for g in golden["graphs"])


@pytest.mark.gpu
def test_reference_suite_recording_replayed_on_the_gpu():
    """GPU: every call the reference's tests made, through the C-ABI on the device, against the
    results those tests accepted: distance work bit-exact, betweenness within 1e-9."""
    ncalls, bad = refsuite.replay()
    assert ncalls >= 140
    assert not bad, bad[:10]


@pytest.mark.gpu
def test_reference_tests_run_directly_on_the_gpu():
    if not refsuite.reference_available():
        pytest.skip("/root/reference is not present on this box: see the replay test")
    result, picked, report = refsuite.run_reference_suite(False)
    assert result.wasSuccessful() and not result.skipped, report
