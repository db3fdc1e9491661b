from arnoldi_b200.ortho import dgks_gs, dgks_mgs  # noqa: F401
