"""innr::backend (src/backend.rs:18-67) with the new `Cuda` variant. `Backend` is #[non_exhaustive] in the
reference, so adding a variant is non-breaking; its Display strings are a stability contract (:114-120)."""
from __future__ import annotations

import ctypes as C
import enum

from . import _lib as L


class Backend(enum.Enum):
    Avx512 = "avx512"
    Avx2Fma = "avx2+fma"
    Neon = "neon"
    Portable = "portable"
    Cuda = "cuda"

    def __str__(self):
        return self.value


def dense_backend(len_: int) -> Backend:
    """Backend the device-resident batch kernels select for `len_`-dimensional vectors: always Cuda (the
    CPU answers of the reference's dense_backend are unchanged and stay in the reference)."""
    v = C.c_int(0)
    L.call("innr_cuda_dense_backend", len_, C.byref(v))
    assert v.value == 1 and L.backend_name() == Backend.Cuda.value
    return Backend.Cuda
