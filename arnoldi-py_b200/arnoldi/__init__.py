"""Import-name alias: ``from arnoldi import partial_schur`` resolves to the B200 path.

The reference package is called ``arnoldi`` (src/arnoldi/__init__.py:3); putting
``<repo>/arnoldi-py_b200`` on ``sys.path`` instead of the reference makes existing user code
(``from arnoldi import partial_schur``, ``from arnoldi.matrices import mark``,
``from arnoldi.utils import arg_largest_real``, README.md:20-31 of the reference) run on the
device without edits.  Everything here re-exports ``arnoldi_b200``.
"""
from arnoldi_b200 import History, __version__, partial_schur  # noqa: F401
