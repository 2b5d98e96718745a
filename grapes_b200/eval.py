"""Drop-in mirror of the reference's ``eval.py`` (/root/reference/eval.py:11-165): same ``evaluate`` signature,
same two modes, same return ``(accuracy, f1)``.

* ``full_batch=True`` (the default of every reference config, eval.py:47-70): ONE forward of the classifier over the
  whole graph.  The reference moves the model to the CPU for this (``eval_on_cpu=True``, main.py:43); here it stays
  on the B200 -- the whole-graph aggregation is ``grapes_aggregate`` (TMA-staged SpMM, csrc/spmm_tma.cu) on the
  gcn_norm structure of the full CSR (:class:`grapes_b200.gcn.GraphNorm`, built once and cached on the graph).
* ``full_batch=False`` (eval.py:71-163): every batch runs on the training engine's kernels, forward only
  (``GrapesEngine.predict``): frontier expansion, bitmap dedup / relabel, fused gather + aggregation, the sampler GCN's
  tcgen05 forward, DETERMINISTIC top-k on the probabilities (eval.py:126-130, ``GRAPES_NOISE_NONE_TOPK_PROBS``), the
  evaluation's block direction ``slice_adjacency(rows=previous_nodes, cols=batch_nodes)`` (eval.py:140-142, a filter of
  the hop's own row expansion), classifier forward, argmax of the target rows.  Batches are enqueued back to back
  (CUDA-graph replays); the host synchronises once, after the last batch.

``eval_on_cpu`` is accepted for signature compatibility and ignored: there is no CPU fallback (north_star).
sklearn's accuracy / micro-F1 on single-label predictions are both the fraction of correct predictions, computed
on the device.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch

from ._lib import GrapesError
from .gcn import GCN, GraphNorm, full_graph_forward
from .graph import DeviceGraph
from .utils import TensorMap, get_logger


def _graph_norm(adjacency: DeviceGraph, data) -> GraphNorm:
    """gcn_norm structure of ``data.edge_index`` (what eval.py:50 feeds GCNConv), cached on the graph object."""
    ei = getattr(data, "edge_index", None)
    key = None if ei is None else (ei.data_ptr(), tuple(ei.shape))
    cached = getattr(adjacency, "_graph_norm", None)
    if cached is None or cached[0] != key:
        cached = (key, GraphNorm(adjacency, edge_index=ei))
        adjacency._graph_norm = cached
    return cached[1]


def _scores(logits_masked: torch.Tensor, y_masked: torch.Tensor) -> Tuple[float, float]:
    if y_masked.dim() == 1:                                           # eval.py:51-55
        acc = (torch.argmax(logits_masked, dim=1) == y_masked).float().mean().item()
        return acc, acc                                               # micro-F1 == accuracy (single label)
    y_pred = logits_masked > 0                                        # eval.py:57-70
    y_true = y_masked > 0.5
    tp = int((y_true & y_pred).sum()); fp = int((~y_true & y_pred).sum()); fn = int((y_true & ~y_pred).sum())
    try:
        precision, recall = tp / (tp + fp), tp / (tp + fn)
        f1 = 2 * (precision * recall) / (precision + recall)
    except ZeroDivisionError:
        f1 = 0.
    return f1, f1


@torch.inference_mode()
def evaluate(gcn_c: GCN,
             gcn_gf: Optional[GCN],
             data,
             args,
             adjacency: DeviceGraph,
             node_map: Optional[TensorMap],
             num_indicators: int,
             device: torch.device,
             mask: torch.Tensor = None,
             eval_on_cpu: bool = True,
             loader=None,
             full_batch: bool = False,
             return_predictions: bool = False,
             engine=None,
             ):
    get_logger().info('Evaluating')
    device = torch.device(device)
    if device.type != "cuda":
        raise GrapesError("grapes_b200 has no CPU fallback: evaluation runs on the CUDA device (eval_on_cpu is ignored)")
    x = data.x.to(device)
    y = data.y.to(device)
    gcn_c = gcn_c.to(device).eval()
    mask_d = mask.to(device) if mask is not None else torch.ones(data.num_nodes, dtype=torch.bool, device=device)

    if full_batch:
        # eval.py:50.  GRAPES_EVAL_TC=1: the fused whole-graph forward with its hidden layer on the tensor cores
        # (gcn.full_graph_forward; written after this round's GPU budget was spent, measured by bench.py's full_graph_eval leg)
        if os.environ.get("GRAPES_EVAL_TC", "0") == "1":
            logits_total = full_graph_forward(gcn_c, x, _graph_norm(adjacency, data))
        else:
            logits_total, _ = gcn_c(x, _graph_norm(adjacency, data))
        res = _scores(logits_total[mask_d], y[mask_d])
        return (res + (logits_total,)) if return_predictions else res

    # ---- mini-batch message passing (eval.py:71-163) on the engine's device path ----
    assert loader is not None, 'loader must be provided if full_batch is False'
    batches = [b[0] if isinstance(b, (tuple, list)) else b for b in loader]
    eng = engine if engine is not None else _eval_engine(adjacency, data, args, gcn_c, gcn_gf, num_indicators, device,
                                                         max([int(b.numel()) for b in batches] + [1]))
    total = sum(int(b.numel()) for b in batches)
    preds = torch.zeros(max(total, 1), dtype=torch.int32, device=device)
    off = 0
    for b in batches:                                                  # every batch is enqueued; nothing is read back
        n = int(b.numel())
        eng.predict(b.to(device), preds[off:off + n])
        off += n
    eng.check_overflow()                                               # the one synchronisation of the evaluation
    all_predictions = preds[:total].to(torch.int64)
    targets = y[mask_d]
    acc = (all_predictions == targets).float().mean().item()          # accuracy_score == micro f1_score here
    return (acc, acc, all_predictions) if return_predictions else (acc, acc)


def _eval_engine(adjacency: DeviceGraph, data, args, gcn_c: GCN, gcn_gf: GCN, num_indicators: int, device, batch_size: int):
    """Engine for a stand-alone ``evaluate(..., full_batch=False)`` call (the training loop passes its own): built once
    per (graph, features, shape) and cached on the graph object; the modules' weights are loaded on every call."""
    from .engine import GrapesEngine
    D = int(gcn_c.gcn_layers[0].lin.weight.shape[0])
    C = int(gcn_c.gcn_layers[-1].lin.weight.shape[0])
    x = data.x.to(device).contiguous()
    key = (x.data_ptr(), tuple(x.shape), D, C, int(batch_size), int(args.num_samples), int(args.sampling_hops),
           bool(num_indicators))
    cached = getattr(adjacency, "_eval_engine", None)
    if cached is None or cached[0] != key:
        eng = GrapesEngine(adjacency, x, data.y.to(device), num_classes=C, batch_size=int(batch_size),
                           num_samples=int(args.num_samples), sampling_hops=int(args.sampling_hops),
                           use_indicators=bool(num_indicators), hidden_dim=D)
        cached = (key, eng)
        adjacency._eval_engine = cached
    eng = cached[1]
    eng.load_state_dicts(gcn_c=gcn_c.state_dict(), gcn_gf=gcn_gf.state_dict())
    return eng
