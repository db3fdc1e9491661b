from arnoldi_b200.krylov_schur import partial_schur  # noqa: F401
