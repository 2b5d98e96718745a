"""Learned node features (``--embed_nodes``; SURVEY.md section 8 row f4, /root/reference/main.py:89-100,116).

The table is an ``nn.Parameter`` inside ``optimizer_c``: its gradient is d loss_c / d x (the sampler nets' gradient is
zeroed by the next ``optimizer_c.zero_grad()`` before it is ever applied, main.py:263,287), non-zero only on the batch's
``all_nodes`` rows, and torch's Adam is dense (rows touched earlier keep moving).  Parity against the oracle with the
same table, weights and injected Gumbel noise: integer contracts bit-exact, the row gradient to 1e-5 (or 2x the
reference's own fp32 deviation from float64), the table after optimiser steps within Adam's sign-flip bound and 1e-4."""
import numpy as np
import pytest
import torch

import test_gpu_engine as E
from parity_utils import _grad_ok
from oracle import reference_port as rp

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,seed", [("tiny", 0), ("small", 1)])
def test_embedding_gradient_matches_oracle(cuda_device, name, seed):
    d, st, eng, train_idx, B = E._setup(name, seed, cuda_device, embed_nodes=True)
    x0 = eng.x.clone()
    rec, ref = E._check_step(st, eng, train_idx[:B], cuda_device, apply_optim=False)
    ref32 = rp.reference_step(st.fp32, train_idx[:B], gumbel_noise=[h["noise"] for h in ref["hops"]], apply_optim=False)
    an = ref["all_nodes"]
    outside = torch.ones(d.num_nodes, dtype=torch.bool)
    outside[an] = False
    assert float(ref["grad_x"][outside].abs().max()) == 0.0       # the gradient is sparse: all_nodes rows only
    assert _grad_ok(rec["grad_x_rows"], ref["grad_x"][an], ref32["grad_x"][an], ())
    assert torch.equal(eng.x, x0)                                  # apply_optim=False leaves the table alone


def test_embedding_table_tracks_oracle_over_adam_steps(cuda_device):
    d, st, eng, train_idx, B = E._setup("small", 2, cuda_device, embed_nodes=True)
    x0 = eng.x.clone()
    nsteps, lr = 3, 1e-3
    touched = torch.zeros(d.num_nodes, dtype=torch.bool)
    for i in range(nsteps):
        rec, ref = E._check_step(st, eng, train_idx[i * B:(i + 1) * B], cuda_device, apply_optim=True, post_optim=(i > 0))
        touched[ref["all_nodes"]] = True
    got, want = eng.x.double().cpu(), st.x.detach().double()
    err = (got - want).abs()
    assert err.max() <= 2.0 * lr * nsteps * 1.01                   # Adam sign-flip bound (see test_three_steps_with_adam)
    scale = want.abs().max()
    frac = (err[touched] > 1e-4 * scale).double().mean().item()
    frac32 = ((st.fp32.x.detach().double() - want).abs()[touched] > 1e-4 * scale).double().mean().item()
    assert frac <= max(0.01, 3.0 * frac32), f"{frac:.4f} of the touched entries off by > 1e-4 (reference fp32: {frac32:.4f})"
    assert torch.equal(eng.x.cpu()[~touched], x0.cpu()[~touched])  # never-touched rows: update is exactly zero
    assert (got[touched] - x0.double().cpu()[touched]).abs().max() > 0.5 * lr
    assert float(eng.adam_steps[0]) == nsteps


def _bitmap(ids: torch.Tensor, N: int, dev):
    W = (N + 31) // 32
    bits = np.zeros(W * 32, dtype=np.uint8)
    bits[ids.numpy()] = 1
    words = np.packbits(bits.reshape(W, 32), axis=1, bitorder="little").view(np.uint32).reshape(W)
    pop = np.array([bin(int(w)).count("1") for w in words], dtype=np.int64)
    pref = np.concatenate([[0], np.cumsum(pop)]).astype(np.int32)
    return torch.from_numpy(words.view(np.int32)).to(dev), torch.from_numpy(pref).to(dev)


def test_adam_embed_kernel_matches_dense_torch_adam(cuda_device):
    """grapes_adam_embed against torch.optim.Adam fed the DENSE gradient (zero rows outside the batch), 4 steps with
    different row sets: rows touched once keep moving afterwards, untouched rows stay bit-identical."""
    from grapes_b200.graph import DeviceGraph
    from grapes_b200._lib import lib, ptr
    dev = cuda_device
    N, F, lr = 1000, 64, 1e-2
    gen = torch.Generator().manual_seed(5)
    g = DeviceGraph.from_edge_index(torch.randint(0, N, (2, 4000), generator=gen), N, device=dev)
    x = torch.randn(N, F, generator=gen)
    ref = torch.nn.Parameter(x.clone().double())
    opt = torch.optim.Adam([ref], lr=lr)
    tab, m, v = x.to(dev), torch.zeros(N, F, device=dev), torch.zeros(N, F, device=dev)
    steps = torch.zeros(2, device=dev)
    L = lib()
    for it in range(4):
        ids = torch.sort(torch.randperm(N, generator=gen)[:100 + 50 * it]).values
        rows = torch.randn(ids.numel(), F, generator=gen) * (10.0 ** (-it))
        bm, pref = _bitmap(ids, N, dev)
        rows_dev = rows.to(dev)
        L.grapes_adam_embed(g.ctx, ptr(tab), ptr(m), ptr(v), N, F, ptr(bm), ptr(pref), ptr(rows_dev), F, lr, 0.9, 0.999,
                            1e-8, ptr(steps), torch.cuda.current_stream().cuda_stream)
        steps[0] += 1
        opt.zero_grad()
        ref.grad = torch.zeros(N, F, dtype=torch.float64)
        ref.grad[ids] = rows.double()
        opt.step()
        assert (tab.double().cpu() - ref.detach()).abs().max() < 2e-6 * (it + 1)
    assert (tab.cpu() != x).any(dim=1).sum() < N


def test_embed_graph_replay_matches_eager(cuda_device):
    """The captured step (table update included) replays to the same table as eager launches, bit for bit."""
    tabs = []
    for use_graph in (False, True):
        d, st, eng, train_idx, B = E._setup("small", 3, cuda_device, embed_nodes=True)
        for i in range(4):
            eng.step(train_idx[i * B:(i + 1) * B].to(cuda_device), use_graph=use_graph,
                     next_targets=train_idx[(i + 1) * B:(i + 2) * B].to(cuda_device))   # ignored: no prefetch with a live table
        eng.check_overflow()
        torch.cuda.synchronize()
        tabs.append(eng.x.clone())
    assert torch.equal(tabs[0], tabs[1])
    assert not torch.equal(tabs[0].cpu(), d.x)


def test_train_with_embed_nodes_runs(cuda_device):
    """train(args) with a feature-less dataset: ValueError without --embed_nodes (main.py:91-94), trains with it."""
    from grapes_b200.args import Arguments
    from grapes_b200.synth import make_synth
    from grapes_b200.train import train
    a = Arguments()
    a.dataset, a.max_epochs, a.batch_size, a.num_samples, a.sampling_hops = "tiny", 2, 32, 8, 2
    a.eval_frequency, a.node_emb_dim = 1, 16
    d = make_synth("tiny", seed=0, features=False)
    d.num_features = 0
    with pytest.raises(ValueError):
        train(a, data=d, device=cuda_device)
    a.embed_nodes = True
    f1, *_ = train(a, data=d, device=cuda_device)
    eng = train.last_engine
    assert eng.embed_nodes and eng.F == 16 and d.x is eng.x and 0.0 <= f1 <= 1.0
    assert float(eng.adam_steps[0]) == 2 * 4                      # 120 train nodes / 32 per batch = 4 batches x 2 epochs
    assert float(eng.emb_exp_avg.abs().sum()) > 0
