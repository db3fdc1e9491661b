"""B200-native Krylov-Schur hot path: drop-in for arnoldi-py's ``partial_schur``."""
from .history import History  # noqa: F401
from .krylov_schur import partial_schur  # noqa: F401

__version__ = "0.1.0"
