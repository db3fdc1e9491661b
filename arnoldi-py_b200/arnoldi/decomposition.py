from arnoldi_b200.decomposition import arnoldi_decomposition  # noqa: F401
