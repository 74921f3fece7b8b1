"""CPU oracle for the pywfo wavefunction-overlap hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker or as the timed CPU baseline -- never on the path that is shipped.

Parity status: PINNED.  ``oracle.pywfo_oracle`` is checked (tests/test_oracle.py)
against
  * the two numeric goldens the reference's own tests hold
    (pywfo/tests/test_pywfo.py:94 and :122-124, RNG recipes in
    tests/golden/make_golden.py), and
  * outputs of the unmodified reference (pywfo.main / pywfo.dets imported from
    /root/reference in the build container) on the bundled H2O2 fixture and on
    seeded random cases, committed as tests/golden/*.npz by
    tests/golden/make_golden.py.
"""
