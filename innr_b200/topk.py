"""Device analogue of innr::TopK (src/topk.rs:47-187).

The reference type is a streaming tracker fed one (id, distance) at a time by batch_knn (src/batch.rs:401-404).
On the device the N inserts collapse into one fused selection; `topk_from_distances` is that selection exposed on
its own: the k smallest by (total_cmp(distance), id), returned ascending like TopK::into_sorted().
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L


def topk_from_distances(distances, k: int):
    d = np.ascontiguousarray(distances, dtype=np.float32).reshape(-1)
    if k == 0 or d.size == 0:
        return []
    kk = min(k, d.size)
    ids = np.zeros(kk, np.uint32)
    ds = np.zeros(kk, np.float32)
    cnt = C.c_size_t(0)
    L.call("innr_cuda_topk_from_distances", d.ctypes.data_as(L.f32p), d.size, k, ids.ctypes.data_as(L.u32p),
           ds.ctypes.data_as(L.f32p), C.byref(cnt))
    return [(int(ids[j]), float(ds[j])) for j in range(cnt.value)]
