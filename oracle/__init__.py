"""CPU oracle for the merfish3d-analysis PixelDecoder hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package
(``merfish3d-analysis_b200/``) imports this directory; only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may.  See ``oracle/decode_oracle.py`` for the pin status.
"""
