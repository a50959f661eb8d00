"""CPU oracle for the OpenProvence scoring-and-pruning hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``open_provence_b200/`` may import
this package; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do, and there only
as the checker / the timed CPU baseline -- never as the product path.

Parity status: the reference's own tests hold **no** numeric golden vector for
the ModernBERT forward, so the oracle is pinned against
outputs of the reference itself run in the build container
(``tests/golden/make_golden.py`` imports the unmodified
``/root/reference/open_provence/modeling_open_provence_standalone.py`` and the
installed ``transformers`` 5.5.0 ModernBERT, and writes the fixtures under
``tests/golden/``).  ``tests/test_oracle_golden.py`` checks the numpy
restatement against those fixtures.
"""
