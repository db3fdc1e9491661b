from arnoldi_b200.history import History  # noqa: F401
from arnoldi_b200.explicit_restarts import explicit_restarts_with_deflation  # noqa: F401
