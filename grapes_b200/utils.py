"""Drop-in mirror of the reference's ``modules/utils.py`` hot-path callables
(/root/reference/modules/utils.py:13-134), backed by the C-ABI CUDA library.

Same names, argument meaning, return ordering and error behaviour; ``adjacency`` is a
:class:`grapes_b200.graph.DeviceGraph` instead of a scipy CSR.  The work runs on the GPU; id tensors (int64 like the
reference) are returned on the device of the id tensors passed in -- the reference's loop keeps its masks and id lists
on the host (main.py:138-140,161-163) and indexes them with these results, so host ids in give host ids out -- and
log-probs / logits stay on the GPU.  These wrappers read the data-dependent sizes back (one sync per call) --
they exist for drop-in use and for parity tests; the training loop uses
:class:`grapes_b200.engine.GrapesEngine`, which never leaves the device.
"""
from __future__ import annotations

import logging
from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

from ._lib import lib, ptr, GrapesError
from .graph import DeviceGraph

NOISE_PHILOX, NOISE_GUMBEL, NOISE_UNIFORM, NOISE_KEYS, NOISE_TOPK_PROBS = 0, 1, 2, 3, 4


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _i32(t: Tensor, device) -> Tensor:
    return t.to(device=device, dtype=torch.int32).contiguous()


def _expand(adjacency: DeviceGraph, rows32: Tensor):
    """rows -> (e_row, e_col, m) with one size read-back."""
    L, ctx, dev = lib(), adjacency.ctx, adjacency.device
    P = rows32.numel()
    cnt = torch.zeros(4, dtype=torch.int32, device=dev)
    if P == 0:
        empty = torch.empty(0, dtype=torch.int32, device=dev)
        return empty, empty, 0, cnt, torch.zeros(1, dtype=torch.int32, device=dev)
    cnt[0] = P
    row_off = torch.empty(P + 1, dtype=torch.int32, device=dev)
    cap_m = (1 << 31) - 1
    L.grapes_row_offsets(ctx, ptr(adjacency.indptr), ptr(rows32), cnt.data_ptr(), max(P, 1), ptr(row_off),
                         cnt.data_ptr() + 4, cap_m, None, None, cnt.data_ptr() + 8, _stream())
    m = int(cnt[1].item())
    e_row = torch.empty(max(m, 1), dtype=torch.int32, device=dev)
    e_col = torch.empty(max(m, 1), dtype=torch.int32, device=dev)
    if m > 0:
        if m > adjacency.max_frontier:
            raise GrapesError("neighbourhood larger than the graph context's max_frontier")
        L.grapes_expand_rows(ctx, ptr(adjacency.indptr), ptr(adjacency.indices), ptr(rows32), cnt.data_ptr(),
                             max(P, 1), ptr(row_off), cnt.data_ptr() + 4, m, ptr(e_row), ptr(e_col), None, _stream())
    return e_row[:m], e_col[:m], m, cnt, row_off


def get_neighborhoods(nodes: Tensor, adjacency: DeviceGraph) -> Tensor:
    """Returns the neighbors of a set of nodes from a given adjacency matrix (utils.py:74-82):
    int64 [2, m] = (node id repeated, neighbour id), row-major by position, neighbours ascending."""
    rows32 = _i32(nodes, adjacency.device)
    e_row, e_col, m, _, _ = _expand(adjacency, rows32)
    src = rows32.to(torch.int64)[e_row.to(torch.int64)]
    return torch.stack([src, e_col.to(torch.int64)], dim=0).to(nodes.device)


def slice_adjacency(adjacency: DeviceGraph, rows: Tensor, cols: Tensor) -> Tensor:
    """Selects a block from a sparse adjacency matrix, given the row and column indices; the result is
    returned as an edge index in GLOBAL ids, [0] = rows, [1] = cols (utils.py:85-95)."""
    L, ctx, dev = lib(), adjacency.ctx, adjacency.device
    rows32, cols32 = _i32(rows, dev), _i32(cols, dev)
    e_row, e_col, m, cnt, _ = _expand(adjacency, rows32)
    if m == 0 or cols32.numel() == 0:
        return torch.empty((2, 0), dtype=torch.int64, device=rows.device)
    bm = torch.zeros(adjacency.num_words, dtype=torch.int32, device=dev)
    cnt[2] = cols32.numel()
    L.grapes_bitmap_set(ctx, ptr(cols32), cnt.data_ptr() + 8, max(cols32.numel(), 1), ptr(bm), _stream())
    out_src = torch.empty(max(m, 1), dtype=torch.int32, device=dev)
    out_dst = torch.empty(max(m, 1), dtype=torch.int32, device=dev)
    ovf = torch.zeros(1, dtype=torch.int32, device=dev)
    L.grapes_slice_block(ctx, ptr(rows32), ptr(e_row), ptr(e_col), cnt.data_ptr() + 4, max(m, 1), ptr(bm),
                         ptr(out_src), ptr(out_dst), max(m, 1), cnt.data_ptr() + 12, ptr(ovf), _stream())
    e = int(cnt[3].item())
    return torch.stack([out_src[:e].to(torch.int64), out_dst[:e].to(torch.int64)], dim=0).to(rows.device)


class TensorMap:
    """A class used to quickly map integers in a tensor to an interval of integers from 0 to
    len(tensor) - 1 (utils.py:98-120).

    Example:
        >>> nodes = torch.tensor([22, 32, 42, 52])
        >>> node_map = TensorMap(size=nodes.max() + 1)
        >>> node_map.update(nodes)
        >>> node_map.map(torch.tensor([52, 42, 32, 22, 22]))
        tensor([3, 2, 1, 0, 0])
    """

    def __init__(self, size, device="cpu"):
        self.map_tensor = torch.empty(int(size), dtype=torch.long, device=device)
        self.values = torch.arange(int(size), device=device)

    def update(self, keys: Tensor):
        keys = keys.to(self.map_tensor.device)
        self.map_tensor[keys] = self.values[:len(keys)]

    def map(self, keys):
        return self.map_tensor[keys.to(self.map_tensor.device)]


class _SelectFn(torch.autograd.Function):
    """log_prob_i depends on logit_i only: d log_prob_i / d logit_i = mask_i - sigmoid(logit_i)."""

    @staticmethod
    def forward(ctx_, logits_flat, dl, log_prob):
        ctx_.save_for_backward(dl)
        return log_prob

    @staticmethod
    def backward(ctx_, grad_out):
        (dl,) = ctx_.saved_tensors
        return grad_out * dl, None, None


_CTX_CACHE = {}


def _any_ctx(device):
    """selection needs a library context but no graph: a 1-node context per device."""
    key = str(device)
    if key not in _CTX_CACHE:
        dev = torch.device(device)
        _CTX_CACHE[key] = DeviceGraph(torch.zeros(2, dtype=torch.int64, device=dev),
                                      torch.zeros(0, dtype=torch.int32, device=dev), 1, max_frontier=1 << 20,
                                      partials_bytes=64 << 20)
    return _CTX_CACHE[key]


def sample_neighborhoods_from_probs(logits: Tensor, neighbor_nodes: Tensor, num_samples: int = -1,
                                    gumbel_noise: Optional[Tensor] = None, noise_mode: Optional[int] = None,
                                    rng_state: Optional[Tensor] = None,
                                    ) -> Tuple[Tensor, Tensor, Dict[str, Tensor]]:
    """Gumbel-top-k node selection + Bernoulli log-probability of the chosen mask (utils.py:13-71).

    ``gumbel_noise`` (optional, length n) injects the reference's noise; otherwise uniforms are drawn
    on the device (Philox).  Returns (sampled ids ascending, log_prob [n] carrying grad, stats)."""
    if not logits.is_cuda:
        raise GrapesError("grapes_b200 has no CPU fallback: logits must live on a CUDA device")
    dev = logits.device
    k = int(num_samples)
    n = neighbor_nodes.shape[0]
    if k < n:
        assert k > 0                                             # utils.py:35
    elif k <= 0:
        k = max(n, 1)
    holder = _any_ctx(dev)
    L, ctx = lib(), holder.ctx
    lf = logits.detach().reshape(-1).to(torch.float32).contiguous()
    assert lf.numel() == n
    nb32 = _i32(neighbor_nodes, dev)
    cnt = torch.zeros(4, dtype=torch.int32, device=dev)
    cnt[0] = n
    cap = max(n, 1)
    ukeys = torch.empty(cap, dtype=torch.int32, device=dev)
    work = torch.zeros(int(L.cdll.grapes_select_work_floats(ctx, cap)), dtype=torch.float32, device=dev)
    sampled = torch.empty(cap, dtype=torch.int32, device=dev)
    log_prob = torch.empty(cap, dtype=torch.float32, device=dev)
    dl = torch.empty(cap, dtype=torch.float32, device=dev)
    stats = torch.zeros(4, dtype=torch.float32, device=dev)
    acc = torch.zeros(2, dtype=torch.float32, device=dev)
    if gumbel_noise is not None:
        mode = NOISE_GUMBEL if noise_mode is None else noise_mode
        noise = gumbel_noise.to(device=dev, dtype=torch.float32).contiguous()
    else:
        mode = NOISE_PHILOX if noise_mode is None else noise_mode
        noise = None
    if rng_state is None:
        rng_state = torch.tensor([int(torch.initial_seed()) & 0x7fffffffffffffff,
                                  int(torch.randint(0, 1 << 62, (1,)).item())], dtype=torch.int64, device=dev)
    L.grapes_select_topk(ctx, ptr(lf), None, ptr(nb32), cnt.data_ptr(), cap, k, mode, ptr(noise), ptr(rng_state),
                         ptr(ukeys), ptr(work), None, ptr(sampled), 0, cnt.data_ptr() + 4, None, None, ptr(log_prob),
                         acc.data_ptr(), ptr(stats), ptr(dl), acc.data_ptr() + 4, None, _stream())
    s = int(cnt[1].item())
    out_nodes = sampled[:s].to(torch.int64).to(neighbor_nodes.device)
    lp = _SelectFn.apply(logits.reshape(-1), dl[:n], log_prob[:n])
    if k >= n:
        return out_nodes, lp, {}                                 # utils.py:31-33
    stats_dict = {"min_prob": stats[0], "max_prob": stats[1], "mean_entropy": stats[2], "std_entropy": stats[3]}
    return out_nodes, lp, stats_dict


def get_logger():
    """Get a default logger that includes a timestamp (utils.py:123-134)."""
    logger = logging.getLogger('')
    logger.handlers = []
    ch = logging.StreamHandler()
    formatter = logging.Formatter('%(asctime)s - %(levelname)s - %(name)s - %(message)s', datefmt='%H:%M:%S')
    ch.setFormatter(formatter)
    logger.addHandler(ch)
    logger.setLevel('INFO')
    return logger
