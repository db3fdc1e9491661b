from arnoldi_b200.utils import (arg_largest_magnitude, arg_largest_real, ordered_schur,  # noqa: F401
                                rand_normalized_vector)
