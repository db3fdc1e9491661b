from arnoldi_b200.history import History  # noqa: F401
