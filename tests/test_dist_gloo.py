"""World-size-2 gloo test (CPU) of the data-parallel host logic: batch sharding and the gradient all-reduce.
Per-rank gradients come from the oracle, so the property checked is the one DESIGN.md section 8 states:
all-reduced gradient == mean of the per-rank (per-batch) gradients, and every rank ends with the same weights."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from grapes_b200.dist import allreduce_mean_, flatten_grads, ranks_agree, shard_batches


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from grapes_b200.synth import make_synth
    from oracle import reference_port as rp
    d = make_synth("tiny", seed=0)
    st = rp.OracleState(d, sampling_hops=2, num_samples=8, seed=7)
    tr = d.train_mask.nonzero().squeeze(1)
    B = 32
    mine = shard_batches(tr.numel() // B, rank, world)
    g = torch.Generator().manual_seed(100 + rank)
    b = mine[0]
    rec = rp.reference_step(st, tr[b * B:(b + 1) * B], apply_optim=False)
    grads = {}
    for key in ("grads_c", "grads_gf", "grads_z"):
        for n, t in rec[key].items():
            grads[f"{key}.{n}"] = t
    local = flatten_grads(grads).clone()
    flat = allreduce_mean_(local.clone())
    # apply the averaged gradient through Adam on every rank
    params = [p for net in (st.gcn_c, st.gcn_gf, st.gcn_z) for _, p in sorted(net.named_parameters())]
    # the agreement the engine uses before it picks the gradient exchange: one failing rank -> every rank sees False
    agree = (ranks_agree(True, "cpu"), ranks_agree(rank != 1, "cpu"), ranks_agree(False, "cpu"))
    out[rank] = dict(batches=mine, local=local, reduced=flat, agree=agree,
                     w0=torch.cat([p.detach().reshape(-1) for p in params]).clone())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_batches_partition():
    for nb, world in ((10, 2), (11, 4), (192, 8), (3, 4)):
        shards = [shard_batches(nb, r, world) for r in range(world)]
        assert len({len(s) for s in shards}) == 1                       # same number of steps on every rank
        flat = sorted(i for s in shards for i in s)
        assert flat == list(range((nb // world) * world))               # disjoint, in loader order
        for r, s in enumerate(shards):
            assert all(i % world == r for i in s)


def test_gradient_allreduce_is_mean_of_per_rank_gradients():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert out[0]["batches"][0] == 0 and out[1]["batches"][0] == 1
    mean = (out[0]["local"] + out[1]["local"]) / 2
    assert not torch.equal(out[0]["local"], out[1]["local"])             # different batches -> different gradients
    for r in range(world):
        assert torch.allclose(out[r]["reduced"], mean, rtol=1e-6, atol=0)
    assert torch.equal(out[0]["reduced"], out[1]["reduced"])             # identical on every rank -> identical Adam steps
    assert torch.equal(out[0]["w0"], out[1]["w0"])
    # enable_data_parallel's peer -> NCCL fall-back is a GROUP decision (grapes_b200.dist.ranks_agree)
    assert out[0]["agree"] == (True, False, False) and out[1]["agree"] == (True, False, False)
