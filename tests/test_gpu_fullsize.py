"""BASELINE.json's full sizes (products-shaped: 2.45 M nodes / 123.7 M nnz, 3 hops, B 1024, k 256; arxiv-shaped)
checked through size-independent properties -- the CPU oracle needs ~0.35 s per step there and the scipy CSR does not
travel, so the integer contracts are restated with plain torch index ops on the device and must hold bit-exactly:

  * batch_nodes = sorted unique (prev rows that have neighbours  U  their neighbours)          main.py:183-190
  * neighbor_nodes = batch_nodes \\ prev, ascending                                              main.py:187-190
  * sampled: exactly k ids, ascending, subset of neighbor_nodes                                 utils.py:44-60
  * prev_{h+1} = [targets | sampled_h]                                                          main.py:236-247
  * block_h = A[rows = prev_{h+1}][:, cols = prev_h] in row-major / ascending-column order      utils.py:85-95
  * all_nodes = sorted unique (targets U sampled_0 U ... )                                      main.py:252
  * losses finite; a replay of the same batch with the same noise is bit-identical (no fp atomics anywhere).
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _neighbours(indptr, indices, rows):
    """(row position, neighbour id) pairs of CSR rows `rows`, row-major, ascending neighbour (= get_neighborhoods)."""
    beg, end = indptr[rows], indptr[rows + 1]
    cnt = end - beg
    pos = torch.repeat_interleave(torch.arange(rows.numel(), device=rows.device), cnt)
    start = torch.cumsum(cnt, 0) - cnt
    off = torch.arange(int(cnt.sum()), device=rows.device) - start[pos]
    return pos, indices[beg[pos] + off].long()


@pytest.mark.parametrize("name", ["arxiv", "products"])
def test_full_size_step_properties(cuda_device, name):
    from bench import build_workload
    from grapes_b200.engine import GrapesEngine
    from grapes_b200.graph import DeviceGraph
    dev = cuda_device
    cfg, indptr, indices, x, y, train_idx = build_workload(name, 0, dev)
    N, B, k, H = cfg["N"], cfg["batch_size"], cfg["num_samples"], cfg["sampling_hops"]
    g = DeviceGraph(indptr, indices, N)
    eng = GrapesEngine(g, x, y, num_classes=cfg["C"], batch_size=B, num_samples=k, sampling_hops=H, hidden_dim=256, seed=0)
    targets = train_idx[3 * B:4 * B]
    gen = torch.Generator(device=dev).manual_seed(5)
    cap = eng.cap_n
    u = torch.rand(H, cap, generator=gen, device=dev).clamp_(1e-7, 1 - 1e-7)
    noise = [(-torch.log(-torch.log(u[h]))).contiguous() for h in range(H)]
    rec = eng.step(targets, gumbel_noise=noise, apply_optim=False, record=True)
    eng.check_overflow()
    prev = targets.long()
    union = [targets.long()]
    for h, a in enumerate(rec["hops"]):
        assert torch.equal(a["prev"].long(), prev), f"hop {h}: prev"
        pos, nb = _neighbours(indptr, indices, prev)
        batch = torch.unique(torch.cat([prev[pos], nb]))                      # sorted
        assert torch.equal(a["batch_nodes"].long(), batch), f"hop {h}: batch_nodes"
        is_prev = torch.zeros(N, dtype=torch.bool, device=dev)
        is_prev[prev] = True
        neigh = batch[~is_prev[batch]]
        assert torch.equal(a["neighbor_nodes"].long(), neigh), f"hop {h}: neighbor_nodes"
        # local relabel of the expanded edges (TensorMap.map == rank in batch_nodes)
        loc_src = torch.searchsorted(batch, prev[pos])
        loc_dst = torch.searchsorted(batch, nb)
        assert torch.equal(a["e_src"].long(), loc_src) and torch.equal(a["e_dst"].long(), loc_dst), f"hop {h}: local edges"
        s = a["sampled"].long()
        assert s.numel() == min(k, neigh.numel()), f"hop {h}: |sampled|"
        assert bool((s[1:] > s[:-1]).all()), f"hop {h}: sampled ascending"
        is_nb = torch.zeros(N, dtype=torch.bool, device=dev)
        is_nb[neigh] = True
        assert bool(is_nb[s].all()), f"hop {h}: sampled within neighbours"
        # the sampled set is the top-k of key = log(sigmoid(logit)) + noise over the candidates (utils.py:37-44)
        lg = a["logits_all"].double()[torch.searchsorted(batch, neigh)]
        keys_ref = torch.log(torch.sigmoid(lg)) + noise[h][: neigh.numel()].double()
        keys = a["keys"].double()
        assert float((keys - keys_ref).abs().max()) < 1e-4 * max(1.0, float(keys_ref.abs().max())), f"hop {h}: keys"
        if s.numel() < neigh.numel():
            kth = torch.sort(keys, descending=True).values
            assert kth[k - 1] > kth[k], "tie at the top-k boundary: pick another noise seed"
            assert torch.equal(torch.sort(neigh[torch.topk(keys, k).indices]).values, s), f"hop {h}: top-k set"
        nxt = torch.cat([targets.long(), s])
        # block: rows = nxt (newer set), cols = prev, row-major by position in rows, ascending column id
        rpos, rnb = _neighbours(indptr, indices, nxt)
        keep = is_prev[rnb]
        blk = torch.stack([nxt[rpos][keep], rnb[keep]])
        assert torch.equal(a["block_edges"].long(), blk), f"hop {h}: block"
        union.append(s)
        prev = nxt
    all_nodes = torch.unique(torch.cat(union))
    assert torch.equal(rec["all_nodes"].long(), all_nodes)
    assert torch.equal(rec["target_local"].long(), torch.searchsorted(all_nodes, targets.long()))
    sc = rec["scalars"]
    assert all(map(lambda v: v == v and abs(v) < 1e30, (sc["loss_c"], sc["loss_gfn"])))
    g1 = eng.grads.clone()
    rec2 = eng.step(targets, gumbel_noise=noise, apply_optim=False, record=True)
    assert torch.equal(eng.grads, g1), "replay of the same batch is not bit-identical"
    assert rec2["scalars"]["loss_c"] == sc["loss_c"] and rec2["scalars"]["loss_gfn"] == sc["loss_gfn"]
