from arnoldi_b200.matrices import laplace, laplace_eigen, mark  # noqa: F401
