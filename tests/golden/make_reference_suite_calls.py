#!/usr/bin/env python
"""Generates tests/golden/reference_suite_calls.json (run in the container that has /root/reference).

Runs the reference's own unittest classes for the SURVEY 8 path (tests/refsuite.py: SELECTION)
against rustworkx_b200 with the CPU oracle behind the device layer, requires every one of them to
pass, and records each call that reached the device layer: graph snapshot, operation, arguments and
the result the reference's assertions accepted.  The GPU box has no /root/reference; there
tests/test_reference_suite.py replays this file through the CUDA path.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import refsuite  # noqa: E402

rec = refsuite.Recorder()
result, picked, report = refsuite.run_reference_suite(True, rec)
assert result.wasSuccessful() and not result.skipped, report
rec.dump(refsuite.GOLDEN)
print(f"{len(picked)} reference tests passed; {len(rec.calls)} calls on {len(rec.graphs)} graphs -> {refsuite.GOLDEN}")
