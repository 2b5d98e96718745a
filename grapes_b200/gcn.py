"""Drop-in mirror of the reference's ``modules/gcn.py`` ``GCN`` (/root/reference/modules/gcn.py:9-42)
with ``torch_geometric.nn.GCNConv`` (PyG 2.5.2 defaults) replaced by the C-ABI CUDA kernels.

Parameter names follow PyG (``gcn_layers.{i}.lin.weight`` [out, in], ``gcn_layers.{i}.bias`` [out]) so a
reference ``state_dict`` loads unchanged.  ``forward`` returns ``(logits, memory_alloc_MB)`` like the
reference (gcn.py:40-42).
"""
from __future__ import annotations

import math
from typing import List, Union

import torch
import torch.nn as nn
import torch.nn.functional as F

from ._lib import lib, ptr, GrapesError
from .utils import _any_ctx, _stream


class NormAdj:
    """gcn_norm structure of one edge_index on n nodes: dst-sorted CSR (forward), src-sorted CSR
    (backward), deg^-1/2 with the added self-loops."""

    def __init__(self, edge_index: torch.Tensor, n: int):
        if not edge_index.is_cuda:
            raise GrapesError("grapes_b200 has no CPU fallback: edge_index must live on a CUDA device")
        dev = edge_index.device
        self.holder = _any_ctx(dev)
        L, ctx = lib(), self.holder.ctx
        E = int(edge_index.shape[1])
        self.n, self.E = int(n), E
        src = edge_index[0].to(torch.int32).contiguous()
        dst = edge_index[1].to(torch.int32).contiguous()
        capE, capn = max(E, 1), max(self.n, 1)
        if max(capE, capn) > self.holder.max_frontier:
            raise GrapesError("edge list larger than the library context's max_frontier")
        self.cnt = torch.tensor([E, self.n, 0, 0, 0, 0], dtype=torch.int32, device=dev)
        scratch = torch.zeros(capn, dtype=torch.int32, device=dev)
        tmp = torch.empty(capE, dtype=torch.int32, device=dev)
        ovf = torch.zeros(1, dtype=torch.int32, device=dev)
        self.in_off = torch.zeros(capn + 1, dtype=torch.int32, device=dev)
        self.in_src = torch.empty(capE, dtype=torch.int32, device=dev)
        self.out_off = torch.zeros(capn + 1, dtype=torch.int32, device=dev)
        self.out_dst = torch.empty(capE, dtype=torch.int32, device=dev)
        self.dinv = torch.empty(capn, dtype=torch.float32, device=dev)
        c = self.cnt.data_ptr()
        L.grapes_build_csr(ctx, ptr(dst), ptr(src), c, capE, c + 4, capn, ptr(scratch), 0, ptr(self.in_off),
                           ptr(self.in_src), ptr(tmp), ptr(self.dinv), c + 8, ptr(ovf), _stream())
        L.grapes_build_csr(ctx, ptr(src), ptr(dst), c, capE, c + 4, capn, ptr(scratch), 0, ptr(self.out_off),
                           ptr(self.out_dst), ptr(tmp), None, c + 12, ptr(ovf), _stream())

    def aggregate(self, x: torch.Tensor, transpose: bool = False, bias=None) -> torch.Tensor:
        L, ctx = lib(), self.holder.ctx
        x = x.contiguous()
        Fdim = x.shape[1]
        out = torch.empty_like(x)
        off, idx = (self.out_off, self.out_dst) if transpose else (self.in_off, self.in_src)
        L.grapes_aggregate(ctx, ptr(x), Fdim, Fdim, None, self.cnt.data_ptr() + 4, max(self.n, 1), ptr(off), ptr(idx),
                           ptr(self.dinv), None, 0, ptr(bias), 0, ptr(out), Fdim, None, None, -1, _stream())
        return out


class GraphNorm(NormAdj):
    """gcn_norm structure of the WHOLE graph, built once from a :class:`DeviceGraph` (full-batch evaluation,
    /root/reference/eval.py:47-56: ``gcn_c(x, edge_index)`` over every edge).  PyG semantics:
    stored self-loops are dropped, one self-loop per node is added, deg = in-degree + 1; the in-neighbour lists are
    the CSR of the transposed adjacency (edge_index[0] = source row, edge_index[1] = destination column),
    ascending source inside a row.  Two builders of the same arrays (see ``__init__``): the default sorts the
    (destination, source) keys with ``torch.sort`` (one-off set-up, GPU-verified); the opt-in "lib" builder uses the
    library's own kernels -- ``grapes_row_offsets`` + ``grapes_expand_rows`` turn the CSR back into its edge list,
    ``grapes_build_csr`` (counting sort by destination, per-row sort, self-loops dropped, deg^-1/2) is the routine that
    builds every hop's structure.  One-off, cached on the graph; the aggregation itself is ``grapes_aggregate``
    (TMA-staged SpMM) straight on these arrays.  nnz must fit int32 offsets (papers100M-shape needs sharding)."""

    def __init__(self, graph, edge_index=None, builder=None):
        """``builder``: "torch" (default; GPU-verified in round 1 and 2) sorts the (destination, source) keys with
        ``torch.sort``; "lib" (``GRAPES_GRAPHNORM_BUILDER=lib``) builds the same arrays with the library's own kernels
        (``grapes_row_offsets`` / ``grapes_expand_rows`` / ``grapes_build_csr``).  The "lib" builder was written after this
        round's GPU budget was spent: it is exercised by ``bench.py --full-eval`` (a process of its own) and by
        the opt-in test ``tests/test_gpu_eval.py::test_full_graph_forward_tensor_core_path`` (GRAPES_TEST_UNVERIFIED=1)."""
        import os
        builder = builder or os.environ.get("GRAPES_GRAPHNORM_BUILDER", "torch")
        self._own_ctx = None
        self.graph = graph
        if builder == "lib":
            self._build_lib(graph, edge_index)
        else:
            self._build_torch(graph, edge_index)

    def _build_torch(self, graph, edge_index=None):
        """``edge_index`` given: the structure of THAT edge list (duplicates kept and counted, exactly what
        ``gcn_c(x, data.edge_index)`` sees in eval.py:50); otherwise the graph's canonical CSR (duplicates collapsed
        by main.py:134)."""
        dev = graph.device
        N = graph.num_nodes
        self.holder = graph
        if edge_index is not None:
            ei = edge_index.to(device=dev, dtype=torch.int64)
            src, dst = ei[0], ei[1]
        else:
            counts = graph.indptr[1:] - graph.indptr[:-1]
            src = torch.repeat_interleave(torch.arange(N, device=dev, dtype=torch.int64), counts)
            dst = graph.indices.to(torch.int64)
        if src.numel() >= (1 << 31) - 1:
            raise GrapesError("full-graph aggregation needs nnz < 2^31 per device")
        keep = src != dst
        key = dst[keep] * N + src[keep]
        del src, dst, keep
        self.n, self.E = N, int(key.numel())
        key = torch.sort(key).values
        d = torch.div(key, N, rounding_mode="floor")
        self.in_src = (key - d * N).to(torch.int32)
        del key
        indeg = torch.bincount(d, minlength=N)
        del d
        self.in_off = torch.zeros(N + 1, dtype=torch.int32, device=dev)
        self.in_off[1:] = torch.cumsum(indeg, 0).to(torch.int32)
        self.dinv = (1.0 / torch.sqrt((indeg + 1).to(torch.float32))).contiguous()
        self.cnt = torch.tensor([int(self.in_src.numel()), N, 0, 0, 0, 0], dtype=torch.int32, device=dev)
        self.out_off = self.out_dst = None

    def _build_lib(self, graph, edge_index=None):
        """``edge_index`` given: the structure of THAT edge list (duplicates kept and counted, exactly what
        ``gcn_c(x, data.edge_index)`` sees in eval.py:50); otherwise the graph's canonical CSR (duplicates collapsed
        by main.py:134)."""
        import ctypes
        dev = graph.device
        N = graph.num_nodes
        L = lib()
        st = _stream()
        E = int(edge_index.shape[1]) if edge_index is not None else graph.nnz
        if E >= (1 << 31) - 1:
            raise GrapesError("full-graph aggregation needs nnz < 2^31 per device")
        capE, capn = max(E, 1), max(N, 1)
        # a context whose scan scratch covers E entries (the graph's own is sized for frontiers)
        self._own_ctx = ctypes.c_void_p()
        rc = L.cdll.grapes_ctx_create(dev.index, N, max(capE, capn), 1 << 20, ctypes.byref(self._own_ctx))
        if rc != 0:
            raise GrapesError(f"grapes_ctx_create failed ({rc}): {L.last_error()}")
        ctx = self._own_ctx
        self.holder = self                       # NormAdj.aggregate reads holder.ctx
        self.graph = graph
        self.n, self.E = N, E
        cnt = torch.tensor([E, N, 0, 0, 0, 0], dtype=torch.int32, device=dev)
        c = cnt.data_ptr()
        if edge_index is not None:
            ei = edge_index.to(device=dev)
            src = ei[0].to(torch.int32).contiguous()
            dst = ei[1].to(torch.int32).contiguous()
        else:
            # the CSR back as an edge list: position of a row in `rows` == its id, so e_row is the source id
            rows = torch.arange(N, dtype=torch.int32, device=dev)
            row_off = torch.empty(N + 1, dtype=torch.int32, device=dev)
            src = torch.empty(capE, dtype=torch.int32, device=dev)
            dst = torch.empty(capE, dtype=torch.int32, device=dev)
            cntP = torch.tensor([N, 0, 0, 0], dtype=torch.int32, device=dev)
            L.grapes_row_offsets(ctx, ptr(graph.indptr), ptr(rows), cntP.data_ptr(), capn, ptr(row_off), cntP.data_ptr() + 4,
                                 capE, None, None, cntP.data_ptr() + 8, st)
            L.grapes_expand_rows(ctx, ptr(graph.indptr), ptr(graph.indices), ptr(rows), cntP.data_ptr(), capn, ptr(row_off),
                                 cntP.data_ptr() + 4, capE, ptr(src), ptr(dst), None, st)
            del rows, row_off
        scratch = torch.zeros(capn, dtype=torch.int32, device=dev)
        tmp = torch.empty(capE, dtype=torch.int32, device=dev)
        ovf = torch.zeros(1, dtype=torch.int32, device=dev)
        self.in_off = torch.zeros(capn + 1, dtype=torch.int32, device=dev)
        in_src = torch.empty(capE, dtype=torch.int32, device=dev)
        self.dinv = torch.empty(capn, dtype=torch.float32, device=dev)
        L.grapes_build_csr(ctx, ptr(dst), ptr(src), c, capE, c + 4, capn, ptr(scratch), 0, ptr(self.in_off), ptr(in_src),
                           ptr(tmp), ptr(self.dinv), c + 8, ptr(ovf), st)
        nnz = int(cnt[2].item())                 # stored entries without the dropped self-loops (one-off read-back)
        if int(ovf.item()) != 0:                 # e.g. more hub rows than the context's worklist holds: never a silent result
            raise GrapesError(f"GraphNorm (library builder): grapes_build_csr flagged overflow bits {int(ovf.item())}; "
                              "use builder='torch'")
        self.in_src = in_src[:nnz].clone() if nnz < capE else in_src
        del src, dst, tmp, scratch, in_src
        self.cnt = torch.tensor([nnz, N, 0, 0, 0, 0], dtype=torch.int32, device=dev)
        self.out_off = self.out_dst = None

    @property
    def ctx(self):
        return self._own_ctx if self._own_ctx is not None else self.graph.ctx

    def __del__(self):
        try:
            if getattr(self, "_own_ctx", None):
                lib().cdll.grapes_ctx_destroy(self._own_ctx)
                self._own_ctx = None
        except Exception:
            pass

    def aggregate(self, x, transpose=False, bias=None):
        if transpose:
            raise GrapesError("GraphNorm is forward-only (evaluation); training uses the sampled blocks")
        return super().aggregate(x, False, bias)


def dense_bias_relu_tc(Y: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """relu(Y W^T + b) for a tall Y [n, K] on the tcgen05 kernels (3xTF32, fp32-accurate; ``grapes_gemm_bias_relu_tc``): the
    hidden layer of the whole-graph evaluation forward.  Requires out_features % 128 == 0 (<= 512)."""
    holder = _any_ctx(Y.device)
    L, ctx, st = lib(), holder.ctx, _stream()
    n, K = Y.shape
    D = weight.shape[0]
    ldy = Y.stride(0)
    assert Y.stride(1) == 1 and ldy % 4 == 0 and Y.data_ptr() % 16 == 0
    ldw = (K + 3) // 4 * 4
    w = weight.detach().float().contiguous()
    Wh = torch.empty((D, ldw), dtype=torch.float32, device=Y.device)
    Wl = torch.empty((D, ldw), dtype=torch.float32, device=Y.device)
    L.grapes_split_tf32(ctx, ptr(w), K, D, K, ptr(Wh), ptr(Wl), ldw, st)
    H = torch.empty((n, D), dtype=torch.float32, device=Y.device)
    cnt = torch.tensor([n], dtype=torch.int32, device=Y.device)
    b = bias.detach().float().contiguous()
    L.grapes_gemm_bias_relu_tc(ctx, ptr(Y), ldy, cnt.data_ptr(), n, K, ptr(Wh), ptr(Wl), ldw, D, ptr(b), ptr(H), D, st)
    return H


@torch.no_grad()
def full_graph_forward(gcn: "GCN", x: torch.Tensor, gn: GraphNorm) -> torch.Tensor:
    """``gcn_c(x, edge_index)`` over the whole graph (eval.py:50) for the two-layer classifier, forward only, every product on
    this library's kernels: Y = A_hat X (TMA-staged SpMM at the narrower width), H = relu(Y W1^T + b1) on the tensor cores,
    Z = H W2^T, logits = A_hat Z + b2.  Falls back to the module's own forward for other depths / hidden widths."""
    layers = list(gcn.gcn_layers)
    D = layers[0].lin.weight.shape[0] if layers else 0
    if len(layers) != 2 or D % 128 != 0 or D > 512 or x.shape[0] < 4096 or gcn.training and gcn.dropout > 0:
        return gcn(x, gn)[0]
    n, F = x.shape
    xp = x.float()
    if F % 4 != 0:                                    # 16-byte rows for the TMA-staged aggregation and the A tiles
        xp = torch.zeros((n, (F + 3) // 4 * 4), dtype=torch.float32, device=x.device)
        xp[:, :F].copy_(x)
        Y = gn.aggregate(xp)
    else:
        Y = gn.aggregate(xp.contiguous())
    H = dense_bias_relu_tc(Y[:, :F] if Y.shape[1] != F else Y, layers[0].lin.weight, layers[0].bias)
    w2 = layers[1].lin.weight.detach().float().contiguous()
    C = w2.shape[0]
    C4 = (C + 3) // 4 * 4
    Z = _gemm(3, H, D, w2, D, n, C, D, ldc=C4) if C4 != C else _gemm(3, H, D, w2, D, n, C, D)
    b2 = layers[1].bias.detach().float()
    if C4 != C:
        b2 = F_pad(b2, C4 - C)
    return gn.aggregate(Z, bias=b2.contiguous())[:, :C]


def F_pad(t: torch.Tensor, k: int) -> torch.Tensor:
    return F.pad(t, (0, k))


def _gemm(layout, A, lda, B, ldb, M, N, K, bias=None, relu=0, ldc=None):
    holder = _any_ctx(A.device)
    ldc = N if ldc is None else ldc
    C = torch.empty((M, N), dtype=torch.float32, device=A.device) if ldc == N else \
        torch.zeros((M, ldc), dtype=torch.float32, device=A.device)
    if M > 0:
        lib().grapes_gemm(holder.ctx, layout, ptr(A), lda, ptr(B), ldb, ptr(C), ldc, None, M, N, K, ptr(bias), relu,
                          None, 0, _stream())
    return C


def _gemm_tn(A, B, R, M, N):
    holder = _any_ctx(A.device)
    out = torch.zeros((M, N), dtype=torch.float32, device=A.device)
    if R > 0:
        lib().grapes_gemm_tn(holder.ctx, ptr(A), M, ptr(B), N, None, R, M, N, 1.0, 0, ptr(out), _stream())
    return out


class _GCNConvFn(torch.autograd.Function):
    """out = A_hat (x W^T) + b, computed at the narrower width: (A_hat x) W^T when in <= out."""

    @staticmethod
    def forward(ctx_, x, weight, bias, adj: NormAdj):
        x = x.contiguous().float()
        w = weight.contiguous().float()
        n, I = x.shape
        O = w.shape[0]
        ctx_.adj, ctx_.agg_first = adj, I <= O
        if ctx_.agg_first:
            y = adj.aggregate(x)
            out = _gemm(3, y, I, w, I, n, O, I, bias=bias)
            ctx_.save_for_backward(y, w)
        else:
            O4 = (O + 3) // 4 * 4
            if O4 != O and n >= 4096:                                 # float4 / TMA aggregation path: pad the width
                h = _gemm(3, x, I, w, I, n, O, I, ldc=O4)
                b4 = None if bias is None else F.pad(bias.detach().float(), (0, O4 - O))
                out = adj.aggregate(h, bias=b4)[:, :O]
            else:
                h = _gemm(3, x, I, w, I, n, O, I)
                out = adj.aggregate(h, bias=bias)
            ctx_.save_for_backward(x, w)
        return out

    @staticmethod
    def backward(ctx_, dout):
        saved, w = ctx_.saved_tensors
        adj = ctx_.adj
        dout = dout.contiguous().float()
        n, O = dout.shape
        I = w.shape[1]
        need_dx = ctx_.needs_input_grad[0]
        db = dout.sum(0) if ctx_.needs_input_grad[2] else None
        dx = None
        if ctx_.agg_first:
            dW = _gemm_tn(dout, saved, n, O, I)                       # dout^T (A_hat x)
            if need_dx:
                dy = _gemm(1, dout, O, w, I, n, I, O)                # dout W
                dx = adj.aggregate(dy, transpose=True)
        else:
            dh = adj.aggregate(dout, transpose=True)                  # A_hat^T dout
            dW = _gemm_tn(dh, saved, n, O, I)
            if need_dx:
                dx = _gemm(1, dh, O, w, I, n, I, O)
        return dx, dW, db, None


class _Lin(nn.Module):
    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels))


class GCNConv(nn.Module):
    """PyG 2.5.2 ``GCNConv(in, out)`` with default arguments (improved=False, cached=False,
    add_self_loops=True, normalize=True, bias=True); glorot weight, zero bias."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin = _Lin(in_channels, out_channels)
        self.bias = nn.Parameter(torch.zeros(out_channels))
        self.reset_parameters()

    def reset_parameters(self):
        a = math.sqrt(6.0 / (self.in_channels + self.out_channels))
        with torch.no_grad():
            self.lin.weight.uniform_(-a, a)
            self.bias.zero_()

    def forward(self, x, edge_index):
        adj = edge_index if isinstance(edge_index, NormAdj) else NormAdj(edge_index, x.shape[0])
        return _GCNConvFn.apply(x, self.lin.weight, self.bias, adj)


class GCN(nn.Module):
    def __init__(self, in_features: int, hidden_dims: List[int], dropout: float = 0.):
        super(GCN, self).__init__()
        self.dropout = dropout
        dims = [in_features] + hidden_dims
        gcn_layers = []
        for i in range(len(hidden_dims) - 1):
            gcn_layers.append(GCNConv(in_channels=dims[i], out_channels=dims[i + 1]))
        gcn_layers.append(GCNConv(in_channels=dims[-2], out_channels=dims[-1]))
        self.gcn_layers = nn.ModuleList(gcn_layers)

    def forward(self, x: torch.Tensor, edge_index: Union[torch.Tensor, List[torch.Tensor]]):
        layerwise_adjacency = type(edge_index) == list
        for i, layer in enumerate(self.gcn_layers[:-1], start=1):
            edges = edge_index[-i] if layerwise_adjacency else edge_index
            x = torch.relu(layer(x, edges))
            x = F.dropout(x, p=self.dropout, training=self.training)
        edges = edge_index[0] if layerwise_adjacency else edge_index
        logits = self.gcn_layers[-1](x, edges)
        logits = F.dropout(logits, p=self.dropout, training=self.training)
        memory_alloc = torch.cuda.memory_allocated() / (1024 * 1024)
        return logits, memory_alloc
