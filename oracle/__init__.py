"""oracle/ -- TEST INFRASTRUCTURE, not product code.

CPU restatements (numpy / torch-CPU) of the reference algorithms on the hot path named by
BASELINE.json's north_star, each function citing the reference file:line it follows.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the checker / the timed
CPU baseline.  ``deeprl_b200`` never imports it; the product path fails loudly when the
CUDA library is missing instead of falling back to anything here.

Pinning: the reference ships no tests or golden vectors, so the oracle is
pinned against OUTPUTS OF THE REFERENCE ITSELF, imported in the build container through
``oracle/ref_shim.py`` (in-memory import of /root/reference, no source copy).  The script
``tests/golden/make_golden.py`` writes those outputs to ``tests/golden/*.npz`` and
``tests/test_oracle_golden.py`` checks every oracle function against them.
"""
