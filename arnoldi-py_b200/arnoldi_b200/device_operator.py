"""Device operators: applying an operator WITHOUT stored entries inside the device loop.

SURVEY.md section 8f-2 / README.md:119 of the reference ("LinearOperator support").  The
reference only needs ``A.shape``, ``A.dtype`` and ``A @ x`` (decomposition.py:44,58); on the
device the equivalent is an object with

    shape, dtype
    device_apply(x_ptr, y_ptr, n, is_real, stream_ptr)

which ENQUEUES ``y = A x`` on the CUDA stream ``stream_ptr`` (raw ``cudaStream_t``) and returns
without synchronising.  ``x_ptr`` / ``y_ptr`` are device addresses of n entries: float64 when
``is_real`` (real operator, real start vector: the basis is held as float64), else interleaved
complex128.  The solver calls it once per Arnoldi step from ``ab200_expand``
(``ab200_set_operator``); no n-length data visits the host.

``TorchOperator`` adapts any function of a torch CUDA tensor (PyTorch here is plumbing: it
wraps the solver's buffers and stream; the orthogonalisation, restart and control logic are
this library's kernels as for a CSR operator).
"""
from __future__ import annotations

import numpy as np


class _DevicePointer:
    """``__cuda_array_interface__`` view of a raw device pointer."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {
            "shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 3,
            "strides": None}


class TorchOperator:
    """``fn(x) -> y`` on torch CUDA tensors as a device operator.

    ``fn`` receives a 1-D tensor (float64 or complex128, length n) that aliases the solver's
    memory -- it must not keep a reference to it -- and returns a tensor of the same shape and
    dtype.  ``dtype`` declares whether the operator is real (float64) or complex (complex128).
    """

    def __init__(self, fn, n, dtype=np.float64):
        self.fn = fn
        self.shape = (int(n), int(n))
        self.dtype = np.dtype(dtype)
        self.calls = 0

    def device_apply(self, x_ptr, y_ptr, n, is_real, stream_ptr):
        import torch
        typestr = "<f8" if is_real else "<c16"
        stream = torch.cuda.ExternalStream(stream_ptr) if stream_ptr else torch.cuda.current_stream()
        with torch.cuda.stream(stream):
            x = torch.as_tensor(_DevicePointer(x_ptr, n, typestr), device="cuda")
            y = torch.as_tensor(_DevicePointer(y_ptr, n, typestr), device="cuda")
            y.copy_(self.fn(x))
        self.calls += 1
