"""CPU oracle for the constrained-beam-search retrieval path.

TEST INFRASTRUCTURE ONLY. Nothing under ``ripor_b200/`` imports this package; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs do, and there only
as the checker or the timed CPU baseline. See the headers of beam.py, t5_math.py and ref_literal.py for
the reference file:line each function follows and for how the restatement is pinned.
"""
