"""Drop-in mirror of the reference's ``eval.py`` (/root/reference/eval.py:11-165): same ``evaluate`` signature,
same two modes, same return ``(accuracy, f1)``.

* ``full_batch=True`` (the default of every reference config, eval.py:47-70): ONE forward of the classifier over the
  whole graph.  The reference moves the model to the CPU for this (``eval_on_cpu=True``, main.py:43); here it stays
  on the B200 -- the whole-graph aggregation is ``grapes_aggregate`` (TMA-staged SpMM, csrc/spmm_tma.cu) on the
  gcn_norm structure of the full CSR (:class:`grapes_b200.gcn.GraphNorm`, built once and cached on the graph).
* ``full_batch=False`` (eval.py:71-163): the hop loop with DETERMINISTIC top-k on the probabilities
  (eval.py:126-130) and -- differently from training -- ``slice_adjacency(rows=previous_nodes, cols=batch_nodes)``
  (eval.py:140-142).  Built from the same drop-in pieces as the training path (``get_neighborhoods``,
  ``slice_adjacency``, ``TensorMap``, ``GCN``, the selection kernel in ``GRAPES_NOISE_NONE_TOPK_PROBS`` mode).

``eval_on_cpu`` is accepted for signature compatibility and ignored: there is no CPU fallback (north_star).
sklearn's accuracy / micro-F1 on single-label predictions are both the fraction of correct predictions, computed
on the device.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from ._lib import GrapesError
from .gcn import GCN, GraphNorm
from .graph import DeviceGraph
from .utils import (NOISE_TOPK_PROBS, TensorMap, get_logger, get_neighborhoods, sample_neighborhoods_from_probs,
                    slice_adjacency)


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
        logits_total, _ = gcn_c(x, _graph_norm(adjacency, data))            # eval.py:50
        res = _scores(logits_total[mask_d], y[mask_d])
        return (res + (logits_total,)) if return_predictions else res

    # ---- mini-batch message passing (eval.py:71-163) ----
    assert loader is not None, 'loader must be provided if full_batch is False'
    gcn_gf = gcn_gf.to(device).eval()
    N = data.num_nodes
    if node_map is None:
        node_map = TensorMap(size=N, device=device)
    prev_nodes_mask = torch.zeros(N, dtype=torch.bool, device=device)
    batch_nodes_mask = torch.zeros(N, dtype=torch.bool, device=device)
    indicator_features = torch.zeros((N, num_indicators), device=device)
    values = torch.arange(N, device=device)
    all_predictions = []
    for batch_id, batch in enumerate(loader):
        target_nodes = batch[0].to(device)
        previous_nodes = target_nodes.clone()
        all_nodes_mask = torch.zeros_like(prev_nodes_mask)
        all_nodes_mask[target_nodes] = True
        indicator_features.zero_()
        if num_indicators:
            indicator_features[target_nodes, -1] = 1.0
        global_edge_indices = []
        for hop in range(args.sampling_hops):
            neighborhoods = get_neighborhoods(previous_nodes, adjacency)                 # eval.py:94
            prev_nodes_mask.zero_()
            batch_nodes_mask.zero_()
            prev_nodes_mask[previous_nodes] = True
            batch_nodes_mask[neighborhoods.view(-1)] = True
            neighbor_nodes_mask = batch_nodes_mask & ~prev_nodes_mask
            batch_nodes = values[batch_nodes_mask]
            neighbor_nodes = values[neighbor_nodes_mask]
            if num_indicators:
                indicator_features[neighbor_nodes, hop] = 1.0
            node_map.update(batch_nodes)
            local_neighborhoods = node_map.map(neighborhoods)
            if args.use_indicators:
                xb = torch.cat([x[batch_nodes], indicator_features[batch_nodes]], dim=1)
            else:
                xb = x[batch_nodes]
            if neighbor_nodes.numel() > 0:
                node_logits, _ = gcn_gf(xb, local_neighborhoods)                         # eval.py:121
                node_logits = node_logits[node_map.map(neighbor_nodes)]
                # torch.topk(Bernoulli(logits).probs, k=min(c, num_samples)) -> ids ascending (eval.py:126-130)
                k = min(int(neighbor_nodes.size(0)), int(args.num_samples))
                sampled_neighboring_nodes, _, _ = sample_neighborhoods_from_probs(
                    node_logits, neighbor_nodes, k, noise_mode=NOISE_TOPK_PROBS)
            else:
                sampled_neighboring_nodes = neighbor_nodes
            all_nodes_mask[sampled_neighboring_nodes] = True
            batch_nodes = torch.cat([target_nodes, sampled_neighboring_nodes], dim=0)
            k_hop_edges = slice_adjacency(adjacency, rows=previous_nodes, cols=batch_nodes)   # eval.py:140-142
            global_edge_indices.append(k_hop_edges)
            previous_nodes = batch_nodes.clone()
        all_nodes = values[all_nodes_mask]
        node_map.update(all_nodes)
        edge_indices = [node_map.map(e) for e in global_edge_indices]
        logits_total, _ = gcn_c(x[all_nodes], edge_indices)
        predictions = torch.argmax(logits_total, dim=1)
        predictions = predictions[node_map.map(target_nodes)]
        all_predictions.append(predictions)
    all_predictions = torch.cat(all_predictions) if all_predictions else torch.zeros(0, dtype=torch.long, device=device)
    targets = y[mask_d]
    acc = (all_predictions == targets).float().mean().item()          # accuracy_score == micro f1_score here
    return (acc, acc, all_predictions) if return_predictions else (acc, acc)
