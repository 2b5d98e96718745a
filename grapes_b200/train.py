"""``train(args)`` with the reference's contract (/root/reference/main.py:57-364): same ``Arguments``,
same epoch / batch order (sequential, un-shuffled batches of ``train_idx``, last batch partial,
main.py:125-126), same return tuple ``(test_f1, mem_point1, mem_point2, mem_point3)``; the batch body runs
on :class:`grapes_b200.engine.GrapesEngine` instead of scipy + PyG.
"""
from __future__ import annotations

import os
from typing import Optional

import torch

from ._lib import GrapesError
from .args import Arguments
from .data import get_data
from .engine import GrapesEngine
from .gcn import GCN
from .graph import DeviceGraph
from .utils import get_logger


def evaluate(engine: GrapesEngine, data, mask: torch.Tensor, full_batch: bool = True, args=None, graph=None):
    """The reference's two evaluation modes (/root/reference/eval.py:47-70 full-batch, :71-163 mini-batch) on the
    device, with the engine's current weights loaded into drop-in ``GCN`` modules (grapes_b200/eval.py)."""
    from .eval import evaluate as _evaluate
    dev = engine.device
    sd = engine.state_dicts()
    gcn_c = GCN(engine.F, [engine.D, engine.C]).to(dev)
    gcn_c.load_state_dict(sd["gcn_c"])
    gcn_gf = GCN(engine.Fp, [engine.D, 1]).to(dev)
    gcn_gf.load_state_dict(sd["gcn_gf"])
    graph = graph or engine.g
    loader = None
    if not full_batch:
        idx = mask.nonzero().squeeze(1)
        loader = [(b,) for b in torch.split(idx, engine.bsz)]       # DataLoader(TensorDataset(idx), batch_size) main.py:127-132
    ns = args if args is not None else type("A", (), dict(sampling_hops=engine.H, num_samples=engine.k,
                                                          use_indicators=engine.use_ind))()
    return _evaluate(gcn_c, gcn_gf, data, ns, graph, None, engine.num_ind, dev, mask=mask, eval_on_cpu=False,
                     loader=loader, full_batch=full_batch)


def train(args: Arguments, data=None, device: Optional[torch.device] = None, use_cuda_graph: bool = True,
          max_batches: Optional[int] = None):
    logger = get_logger()
    if not torch.cuda.is_available():
        raise GrapesError("grapes_b200 has no CPU fallback: a CUDA (sm_100a) device is required")
    device = device or torch.device("cuda", torch.cuda.current_device())
    if data is None:
        path = os.path.join(os.getcwd(), 'data', args.dataset)
        data, num_features, num_classes = get_data(root=path, name=args.dataset, seed=args.seed, split_id=args.split_id)
    else:
        num_features, num_classes = data.num_features, data.num_classes
    if args.embed_nodes or data.x is None:
        if not args.embed_nodes:
            raise ValueError('Dataset does not contain node features, and embed_nodes is False. '
                             'Did you mean to run with --embed_nodes=True?')          # main.py:91-94
        # main.py:95-100: embeddings = nn.Parameter(FloatTensor(N, node_emb_dim).normal_()), appended to optimizer_c.
        # Here the table lives in HBM and the engine's optimiser launch updates it in place (grapes_adam_embed).
        logger.info('Using learned node embeddings for features')
        gen = torch.Generator().manual_seed(0 if args.seed is None else args.seed)
        data.x = torch.empty(data.num_nodes, args.node_emb_dim).normal_(generator=gen).to(device)
        data.num_features = num_features = args.node_emb_dim
    if args.model_type != 'gcn':
        raise ValueError("only model_type='gcn' is wired in the reference's train() (main.py:109)")
    if args.dropout != 0.:
        raise NotImplementedError("dropout > 0 is not used by any reference config; use grapes_b200.gcn.GCN directly")

    graph = DeviceGraph.from_edge_index(data.edge_index, data.num_nodes, device=device)
    x_dev = data.x.to(device).contiguous()
    if args.embed_nodes:
        data.x = x_dev                       # evaluation reads the table the optimiser updates
    engine = GrapesEngine(graph, x_dev, data.y.to(device), num_classes=num_classes, embed_nodes=args.embed_nodes,
                          batch_size=args.batch_size, num_samples=args.num_samples, sampling_hops=args.sampling_hops,
                          use_indicators=args.use_indicators, hidden_dim=args.hidden_dim, lr_gc=args.lr_gc,
                          lr_gf=args.lr_gf, loss_coef=args.loss_coef, log_z_init=args.log_z_init,
                          reg_param=args.reg_param, random_sampling=args.random_sampling,
                          reinforce_baseline=args.reinforce_baseline, seed=0 if args.seed is None else args.seed)
    train_idx = data.train_mask.nonzero().squeeze(1).to(device)
    batches = [b.to(torch.int32).contiguous() for b in torch.split(train_idx, args.batch_size)]   # DataLoader(TensorDataset(train_idx), batch_size)
    if max_batches is not None:
        batches = batches[:max_batches]

    mem1, mem2, mem3 = [], [], []
    logger.info('Training')
    test_f1 = 0.0
    for epoch in range(1, args.max_epochs + 1):
        acc_c = torch.zeros((), device=device)
        acc_gfn = torch.zeros((), device=device)
        for bi, batch in enumerate(batches):
            # the next batch's reset + hop-0 front end is enqueued next to this step's classifier tail
            engine.step(batch, use_graph=use_cuda_graph, next_targets=batches[bi + 1] if bi + 1 < len(batches) else None)
            acc_c += engine.scal[0] / len(batches)               # deferred: no .item() inside the loop
            acc_gfn += engine.scal[4] / len(batches)
            mb = torch.cuda.memory_allocated() / (1024 * 1024)
            mem1.append(torch.cuda.max_memory_allocated() / (1024 * 1024)); mem2.append(mb); mem3.append(mb)
        engine.check_overflow()
        if (epoch + 1) % args.eval_frequency == 0:                # main.py:319
            accuracy, f1 = evaluate(engine, data, data.val_mask, full_batch=args.eval_full_batch)
            logger.info(f'loss_gfn={acc_gfn.item():.6f}, loss_c={acc_c.item():.6f}, '
                        f'valid_accuracy={accuracy:.3f}, valid_f1={f1:.3f}')
    test_accuracy, test_f1 = evaluate(engine, data, data.test_mask, full_batch=args.eval_full_batch)
    logger.info(f'test_accuracy={test_accuracy:.3f}, test_f1={test_f1:.3f}')
    train.last_engine = engine
    return test_f1, mem1, mem2, mem3
