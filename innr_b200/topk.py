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


class TopK:  # src/topk.rs:47-187
    """innr::TopK with the selection on the device. The reference keeps a sorted buffer and pays one binary search per
    insert; here inserts are buffered on the host and settled by ONE fused selection when a result is needed
    (`threshold`, `into_sorted`) -- the N-inserts-then-read pattern of batch_knn (src/batch.rs:401-404) costs one launch.
    Which of several exactly tied candidates survives is unspecified in the reference; here the
    earlier insert wins, which is also what the reference's strict `Less` test at the boundary does (:101)."""

    _SETTLE_AT = 1 << 20

    def __init__(self, k: int):
        assert k > 0, "innr::TopK: k must be >= 1"  # :65
        self.k = int(k)
        self._ids = np.zeros(0, np.uint32)       # settled: ascending (distance, insertion order), at most k
        self._ds = np.zeros(0, np.float32)
        self._pend_ids: list[int] = []
        self._pend_ds: list[float] = []
        self._count = 0

    def insert(self, id: int, distance: float) -> None:  # :93-110
        self._pend_ids.append(int(id))
        self._pend_ds.append(float(distance))
        self._count += 1
        if len(self._pend_ids) >= self._SETTLE_AT:
            self._settle()

    def _settle(self) -> None:
        if not self._pend_ids:
            return
        ids = np.concatenate([self._ids, np.asarray(self._pend_ids, dtype=np.uint32)])
        ds = np.concatenate([self._ds, np.asarray(self._pend_ds, dtype=np.float32)])
        self._pend_ids, self._pend_ds = [], []
        pos = [p for p, _ in topk_from_distances(ds, self.k)]  # positions: settled entries first, so they win ties
        self._ids, self._ds = ids[pos], ds[pos]

    def threshold(self) -> float:  # :118-124: the k-th best distance, +inf until k candidates were offered
        if self._count < self.k:
            return float("inf")
        self._settle()
        return float(self._ds[-1])

    def __len__(self) -> int:  # :127-129
        return min(self._count, self.k)

    def is_empty(self) -> bool:  # :132-134
        return self._count == 0

    def into_sorted(self):  # :140-145: ascending (id, distance) pairs
        self._settle()
        return [(int(i), float(d)) for i, d in zip(self._ids, self._ds)]
