"""Device-resident CSR graph: what ``adjacency`` is in the reference
(/root/reference/main.py:134-136, a scipy ``csr_matrix`` on the host) becomes a pair of HBM arrays.

Layout in HBM: ``indptr`` int64 [N+1] (papers100M-shape has nnz = 3.2e9), ``indices`` int32 [nnz],
rows sorted, duplicates collapsed, self-loops kept -- exactly scipy's canonical CSR of the bool matrix.
"""
from __future__ import annotations

from typing import Optional

import torch

from ._lib import lib, ptr, GrapesError
import ctypes


def _require_cuda(device) -> torch.device:
    device = torch.device(device)
    if device.type != "cuda" or not torch.cuda.is_available():
        raise GrapesError("grapes_b200 has no CPU fallback: a CUDA (sm_100a) device is required")
    if device.index is None:                         # torch.device('cuda') means the CURRENT device, not device 0
        device = torch.device("cuda", torch.cuda.current_device())
    return device


def csr_from_edge_index(edge_index: torch.Tensor, num_nodes: int, device) -> "tuple[torch.Tensor, torch.Tensor]":
    """``edge_index`` int64 [2, E] (host or device) -> (indptr int64 [N+1], indices int32 [nnz]) in HBM: scipy's
    canonical CSR of main.py:134-136, through the C ABI (``grapes_csr_from_edges``)."""
    device = _require_cuda(device)
    N = int(num_nodes)
    ei = edge_index.to(device=device, dtype=torch.int64)
    if ei.dim() != 2 or ei.shape[0] != 2:
        raise ValueError("edge_index must have shape [2, E]")
    src, dst = ei[0].contiguous(), ei[1].contiguous()
    E = int(src.numel())
    L = lib()
    with torch.cuda.device(device):
        ws_bytes = int(L.cdll.grapes_csr_workspace_bytes(N, E))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=device)
        indptr = torch.empty(N + 1, dtype=torch.int64, device=device)
        cap = torch.empty(max(E, 1), dtype=torch.int32, device=device)
        meta = torch.zeros(2, dtype=torch.int64, device=device)          # [0] nnz, [1] error flag (low int32)
        L.grapes_csr_from_edges(ptr(src), ptr(dst), E, N, ptr(indptr), ptr(cap), meta.data_ptr(), meta.data_ptr() + 8,
                                ptr(ws), ws_bytes, torch.cuda.current_stream(device).cuda_stream)
        nnz, err = (int(v) for v in meta.cpu().tolist())
    if err & 0xffffffff:
        raise ValueError("edge_index holds node ids outside [0, num_nodes)")    # scipy: index exceeds matrix dimensions
    del ws
    return indptr, (cap if nnz == cap.numel() else cap[:nnz].clone())


class DeviceGraph:
    """CSR adjacency + the per-device library context.

    Accepted wherever the reference passes its scipy ``adjacency`` (get_neighborhoods,
    slice_adjacency).  Built from the same ``(edge_index, num_nodes)`` the reference feeds scipy.
    """

    def __init__(self, indptr: torch.Tensor, indices: torch.Tensor, num_nodes: int,
                 max_frontier: Optional[int] = None, partials_bytes: int = 64 << 20):
        device = _require_cuda(indptr.device)
        assert indptr.dtype == torch.int64 and indices.dtype == torch.int32
        assert indptr.numel() == num_nodes + 1
        self.device = device
        self.num_nodes = int(num_nodes)
        self.indptr = indptr.contiguous()
        self.indices = indices.contiguous()
        self.nnz = int(indices.numel())
        self.num_words = (self.num_nodes + 31) // 32
        if max_frontier is None:
            max_frontier = max(self.num_nodes, 1 << 22)
        self.max_frontier = int(min(max_frontier, (1 << 31) - 1))
        self._ctx = ctypes.c_void_p()
        self._extra = []
        L = lib()
        rc = L.cdll.grapes_ctx_create(device.index, self.num_nodes, self.max_frontier, partials_bytes,
                                      ctypes.byref(self._ctx))
        if rc != 0:
            raise GrapesError(f"grapes_ctx_create failed ({rc}): {L.last_error()}")

    @property
    def ctx(self):
        return self._ctx

    def new_ctx(self, partials_bytes: int = 64 << 20):
        """Another library context for the same graph: one per CUDA stream that issues calls concurrently
        (a ctx owns the scan / split-K scratch and is not re-entrant)."""
        c = ctypes.c_void_p()
        L = lib()
        rc = L.cdll.grapes_ctx_create(self.device.index, self.num_nodes, self.max_frontier, partials_bytes,
                                      ctypes.byref(c))
        if rc != 0:
            raise GrapesError(f"grapes_ctx_create failed ({rc}): {L.last_error()}")
        self._extra.append(c)
        return c

    def __del__(self):
        try:
            for c in getattr(self, "_extra", []):
                lib().cdll.grapes_ctx_destroy(c)
            self._extra = []
            if getattr(self, "_ctx", None):
                lib().cdll.grapes_ctx_destroy(self._ctx)
                self._ctx = None
        except Exception:
            pass

    @property
    def shape(self):
        return (self.num_nodes, self.num_nodes)

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_edge_index(cls, edge_index: torch.Tensor, num_nodes: int, device="cuda", **kw) -> "DeviceGraph":
        """main.py:134-136: ``csr_matrix((ones bool, edge_index), (N, N))`` -- duplicates collapse, columns sorted,
        self-loops kept.  Built on the device by ``grapes_csr_from_edges`` (counting sort by row + per-row sort/unique,
        csrc/csr_build.cu); one-off, outside the step.  Ids outside [0, N) raise ValueError like scipy."""
        device = _require_cuda(device)
        indptr, indices = csr_from_edge_index(edge_index, num_nodes, device)
        return cls(indptr, indices, num_nodes, **kw)

    @classmethod
    def from_scipy(cls, adjacency, device="cuda", **kw) -> "DeviceGraph":
        adjacency = adjacency.tocsr()
        adjacency.sum_duplicates()
        adjacency.sort_indices()
        device = _require_cuda(device)
        indptr = torch.from_numpy(adjacency.indptr.astype("int64")).to(device)
        indices = torch.from_numpy(adjacency.indices.astype("int32")).to(device)
        return cls(indptr, indices, adjacency.shape[0], **kw)

    def to_scipy(self):
        import numpy as np
        import scipy.sparse as sp
        indptr = self.indptr.cpu().numpy()
        indices = self.indices.cpu().numpy()
        return sp.csr_matrix((np.ones(indices.shape[0], dtype=bool), indices, indptr), shape=self.shape)
