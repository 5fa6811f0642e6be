"""CPU oracle for the emcee red-blue walker-update path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``emcee_b200/`` may import this
package: only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and only as the checker or the
timed CPU arm -- never as the product path.

Parity status: PINNED.  The restatement in ``oracle/redblue.py`` is checked
bit-for-bit against the unmodified reference (dfm/emcee @ 8ab6c0f, imported
from /root/reference in the authoring container) driven by the same
counter-based Philox draws (``oracle/philox.PhiloxRandom`` injected as
``sampler._random``); the resulting vectors are committed under
``tests/golden/`` together with the generator ``oracle/gen_golden.py``.
"""
