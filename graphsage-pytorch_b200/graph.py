"""Adjacency containers: the reference's dict-of-sets -> CSR -> device CSR.

The reference keeps the graph as `defaultdict(set)` (src/dataCenter.py:33,77) and indexes it
per node in Python (src/models.py:279).  The device path wants CSR: `rowptr` int64 [N+1],
`col` int32 [nnz], rows sorted ascending.  `AdjCSR` additionally behaves like the mapping the
reference expects (`adj[node] -> set`), so a graph too large for a dict-of-sets can still be
handed to code written against the reference's interface.
"""
from __future__ import annotations

from collections.abc import Mapping
from typing import Optional, Tuple

import numpy as np
import torch


class AdjCSR(Mapping):
    """Host CSR adjacency with `dict[int -> set[int]]` read semantics (missing key -> empty set,
    as `defaultdict(set)` gives at src/models.py:279)."""

    def __init__(self, rowptr: np.ndarray, col: np.ndarray):
        self.rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
        self.col = np.ascontiguousarray(col, dtype=np.int32)
        if self.rowptr.ndim != 1 or self.rowptr[0] != 0 or self.rowptr[-1] != len(self.col):
            raise ValueError("malformed CSR")

    @property
    def num_nodes(self) -> int:
        return len(self.rowptr) - 1

    def __getitem__(self, node) -> set:
        node = int(node)
        if 0 <= node < self.num_nodes:
            return set(self.col[self.rowptr[node]:self.rowptr[node + 1]].tolist())
        return set()

    def __iter__(self):
        return iter(range(self.num_nodes))

    def __len__(self) -> int:
        return self.num_nodes

    def degree(self) -> np.ndarray:
        return np.diff(self.rowptr)


def adj_to_csr(adj_lists, num_nodes: int) -> Tuple[np.ndarray, np.ndarray]:
    """dict-of-sets (or AdjCSR) -> (rowptr, col), rows ascending.  Keys absent from the dict
    are empty rows; ids >= num_nodes are rejected (the feature table has no such row)."""
    if isinstance(adj_lists, AdjCSR):
        if adj_lists.num_nodes != num_nodes:
            raise ValueError(f"adjacency has {adj_lists.num_nodes} rows, feature table {num_nodes}")
        return adj_lists.rowptr, adj_lists.col
    deg = np.zeros(num_nodes, dtype=np.int64)
    for node, nbrs in adj_lists.items():
        node = int(node)
        if not 0 <= node < num_nodes:
            raise ValueError(f"adjacency key {node} outside the feature table (0..{num_nodes - 1})")
        deg[node] = len(nbrs)
    rowptr = np.zeros(num_nodes + 1, dtype=np.int64)
    np.cumsum(deg, out=rowptr[1:])
    col = np.empty(int(rowptr[-1]), dtype=np.int32)
    for node, nbrs in adj_lists.items():
        if nbrs:
            node = int(node)
            row = np.fromiter(nbrs, dtype=np.int64, count=len(nbrs))
            row.sort()
            col[rowptr[node]:rowptr[node + 1]] = row
    if len(col) and (col.min() < 0 or col.max() >= num_nodes):
        raise ValueError("neighbour id outside the feature table")
    return rowptr, col


class DeviceCSR:
    """CSR resident in HBM (int64 rowptr, int32 col).  Built once per model (the reference
    re-walks the Python dict every step)."""

    def __init__(self, rowptr: np.ndarray, col: np.ndarray, device):
        self.num_nodes = len(rowptr) - 1
        self.nnz = int(len(col))
        self.rowptr = torch.from_numpy(np.ascontiguousarray(rowptr, dtype=np.int64)).to(device)
        self.col = torch.from_numpy(np.ascontiguousarray(col, dtype=np.int32)).to(device)
        if self.nnz == 0:   # keep a valid pointer for the kernels
            self.col = torch.zeros(1, dtype=torch.int32, device=device)
        self.id_bits = max(1, int(self.num_nodes).bit_length())
        self.device = self.rowptr.device

    @classmethod
    def from_device(cls, rowptr: torch.Tensor, col: torch.Tensor) -> "DeviceCSR":
        """Wrap a CSR that already lives in HBM (graphs generated or loaded on the device: 1.6B
        entries never pass through a host dict-of-sets).  Pass the result as `adj_lists`."""
        if rowptr.dtype != torch.int64 or col.dtype != torch.int32 or not rowptr.is_cuda or col.device != rowptr.device:
            raise ValueError("rowptr int64 / col int32 on the same CUDA device")
        self = cls.__new__(cls)
        self.num_nodes = int(rowptr.shape[0]) - 1
        self.nnz = int(col.shape[0])
        self.rowptr, self.col = rowptr.contiguous(), col.contiguous()
        self.id_bits = max(1, int(self.num_nodes).bit_length())
        self.device = rowptr.device
        return self

    @classmethod
    def from_adj(cls, adj_lists, num_nodes: int, device) -> "DeviceCSR":
        rowptr, col = adj_to_csr(adj_lists, num_nodes)
        return cls(rowptr, col, device)
