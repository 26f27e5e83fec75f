"""CPU oracle for the vlite retrieval hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU (numpy + a small C file), the algorithm of the
reference's retrieval path so that the CUDA product in ``vlite_b200/`` can be
checked against it.  It is never imported by the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may use it, and there only as the checker (or as the timed CPU
baseline), never as the thing shipped.

Parity status: PINNED.  Every function here is checked in ``tests/test_oracle.py``
against golden vectors produced by executing the reference's own code
(``/root/reference/vlite/model.py``, ``main.py``, ``ctx.py``) in the build
container -- see ``tests/golden/make_golden.py`` (the generating script) and the
``tests/golden/*.npz|json|ctx`` fixtures it wrote.  The one exception is the
rescore step, which has NO reference implementation:
``oracle.search.rescore_*`` is "parity unpinned" and says so in its docstring.
"""
from . import search, synth, ctxfile  # noqa: F401
