"""CPU oracle for the neuronika dense forward/backward hot path.

TEST INFRASTRUCTURE ONLY.  This package is a numpy restatement of the
reference's ndarray algorithms (``/root/reference/neuronika-variable/src/node``,
``neuronika-nn``, ``neuronika-optim``).  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it, and only as the checker or the CPU
baseline -- never as the product path.  The product (``neuronika_b200``)
fails loudly when the CUDA library is missing; it has no CPU fallback.

Parity pin: the reference is Rust and cannot be compiled in this environment
(no rustc/cargo, crates not vendored), so the oracle is pinned against every
golden vector the reference's own tests hold for this path, transcribed
mechanically from the reference's ``test.rs`` files into ``tests/golden/*.json``
by ``tests/golden/make_goldens.py`` (see ``tests/test_oracle_goldens.py``).
Where the reference code is defective for non-uniform data (SURVEY.md section
8-c: ``utils::accumulate`` and ``convolution_backward_input``), the oracle
implements the intended mathematics, which coincides with every enabled
reference test and with torch CPU autograd.
"""
from .nodes import *  # noqa: F401,F403
from .nodes_next import *  # noqa: F401,F403
