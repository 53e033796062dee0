"""CPU oracle for the LiDARCrafter denoiser hot path.

TEST INFRASTRUCTURE ONLY -- nothing under ``lidarcrafter_b200/`` imports this package.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs
may use it, and only as the checker / the timed CPU baseline.

Parity pinning: the reference ships no tests or golden vectors, so the oracle
is pinned against outputs of the UNMODIFIED reference imported in the build container
(``oracle/ref_import.py``); the generated vectors live in ``tests/golden/`` together with the script
that made them (``tests/golden/make_golden.py``).
"""
