"""CPU oracle for the partial_schur hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker or as the CPU arm that is timed beside the GPU path.  The product
(``arnoldi-py_b200/``) never imports this package and has no CPU fallback.

Parity status: PINNED.  ``tests/golden/*.npz`` were produced by importing the
unmodified reference from ``/root/reference/src`` (script:
``oracle/make_golden.py``) and ``tests/test_oracle_golden.py`` checks this
restatement against every one of them.
"""

from .krylov import (  # noqa: F401
    History,
    arg_largest_magnitude,
    arg_largest_real,
    arnoldi_expand,
    cgs_dgks,
    explicit_restarts_with_deflation,
    mgs_dgks,
    partial_schur,
    rand_unit_vector,
    restart_update,
    sorted_schur,
    spmv,
)
