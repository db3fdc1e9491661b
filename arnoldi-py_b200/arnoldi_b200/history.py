"""Per-eigenpair convergence record (reference: explicit_restarts.py:13-28)."""
from __future__ import annotations

import dataclasses

import numpy as np


@dataclasses.dataclass
class History:
    """``matvecs[k]`` / ``restarts[k]``: the counters at which pair k was last seen converged."""

    matvecs: np.ndarray
    restarts: np.ndarray

    @classmethod
    def from_k(cls, k):
        return cls(np.zeros(k, np.int32), np.zeros(k, np.int32))

    @property
    def k(self):
        return self.matvecs.shape[0]

    @property
    def total_matvecs(self):
        return self.matvecs.sum()
