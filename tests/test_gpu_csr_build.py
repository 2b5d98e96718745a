"""grapes_csr_from_edges against scipy's canonical CSR of main.py:134-136, bit-exact.
Covers the three per-row sort tiers (warp <= 32 entries, shared memory <= 32768, in-HBM hub rows), duplicate edges,
self-loops, isolated nodes, the empty edge list and out-of-range ids (scipy raises ValueError)."""
import numpy as np
import pytest
import torch

from oracle import reference_port as rp

pytestmark = pytest.mark.gpu


def _check(ei: torch.Tensor, N: int, dev):
    from grapes_b200.graph import csr_from_edge_index
    adj = rp.build_adjacency(ei, N)
    indptr, indices = csr_from_edge_index(ei, N, dev)
    assert indptr.dtype == torch.int64 and indices.dtype == torch.int32
    assert np.array_equal(indptr.cpu().numpy(), adj.indptr.astype(np.int64))
    assert int(indices.numel()) == adj.nnz
    assert np.array_equal(indices.cpu().numpy(), adj.indices.astype(np.int32))
    return adj


@pytest.mark.parametrize("N,E,seed", [(1, 5, 0), (7, 40, 1), (1000, 20000, 2), (50000, 400000, 3), (300, 90000, 4)])
def test_random_multigraph_matches_scipy(cuda_device, N, E, seed):
    g = torch.Generator().manual_seed(seed)
    ei = torch.randint(0, N, (2, E), generator=g)                  # duplicates and self-loops at random
    ei = torch.cat([ei, ei[:, : E // 3], torch.arange(N).repeat(2, 1)[:, ::2]], dim=1)
    adj = _check(ei, N, cuda_device)
    assert adj.nnz < ei.shape[1]


def test_empty_and_isolated(cuda_device):
    _check(torch.zeros((2, 0), dtype=torch.int64), 5, cuda_device)
    _check(torch.tensor([[3], [3]]), 9, cuda_device)               # one self-loop, eight isolated nodes
    _check(torch.tensor([[0, 8, 8, 0], [8, 0, 0, 8]]), 9, cuda_device)


@pytest.mark.parametrize("hub_deg", [33, 64, 65, 128, 129, 256, 257, 512, 513, 1024, 1025, 2048, 2049, 32768, 40000])
def test_hub_rows_every_sort_tier(cuda_device, hub_deg):
    """row 5 gets `hub_deg` entries before dedup (shuffled, with repeats): 33..256 -> warp shared-memory tier, 257..2048 -> medium CTA tier,
    2049..32768 -> shared-memory long tier, beyond -> in-HBM bitonic."""
    N = 60000
    g = torch.Generator().manual_seed(hub_deg)
    cols = torch.randint(0, N, (hub_deg,), generator=g)
    cols[: hub_deg // 4] = cols[hub_deg // 4: 2 * (hub_deg // 4)]  # a quarter of the entries are duplicates
    hub = torch.stack([torch.full((hub_deg,), 5), cols])
    rest = torch.randint(0, N, (2, 100000), generator=g)
    ei = torch.cat([hub, rest, hub.flip(0)], dim=1)[:, torch.randperm(2 * hub_deg + 100000, generator=g)]
    _check(ei, N, cuda_device)


def test_reddit_like_dense_rows(cuda_device):
    """every row in the medium tier (Reddit-shape: ~500 entries per row)"""
    N, deg = 3000, 500
    g = torch.Generator().manual_seed(7)
    ei = torch.stack([torch.arange(N).repeat_interleave(deg), torch.randint(0, N, (N * deg,), generator=g)])
    _check(ei[:, torch.randperm(N * deg, generator=g)], N, cuda_device)


def test_out_of_range_ids_raise(cuda_device):
    from grapes_b200.graph import csr_from_edge_index
    with pytest.raises(ValueError):
        csr_from_edge_index(torch.tensor([[0, 4], [1, 2]]), 4, cuda_device)
    with pytest.raises(ValueError):
        csr_from_edge_index(torch.tensor([[0, 1], [-1, 2]]), 4, cuda_device)


def test_products_shape_properties(cuda_device):
    """BASELINE-size graph (2.45 M nodes, 123.7 M directed pairs): size-independent properties instead of scipy --
    rows strictly ascending (sorted + deduplicated), indptr monotone with indptr[N] = nnz, the edge set is the unique
    set of input pairs (compared through a 64-bit key checksum), symmetric input -> symmetric CSR."""
    from grapes_b200.graph import csr_from_edge_index
    dev = cuda_device
    N, pairs = 2_449_029, 61_859_140
    g = torch.Generator(device=dev).manual_seed(0)
    a = torch.randint(0, N, (pairs,), generator=g, device=dev)
    b = torch.randint(0, N, (pairs,), generator=g, device=dev)
    ei = torch.stack([torch.cat([a, b]), torch.cat([b, a])])
    del a, b
    indptr, indices = csr_from_edge_index(ei, N, dev)
    nnz = int(indices.numel())
    assert int(indptr[0]) == 0 and int(indptr[-1]) == nnz
    deg = indptr[1:] - indptr[:-1]
    assert int(deg.min()) >= 0
    rows = torch.repeat_interleave(torch.arange(N, device=dev), deg)
    key = rows * N + indices.long()
    assert bool((key[1:] > key[:-1]).all())                        # globally strictly ascending = sorted + unique
    ref = torch.unique(ei[0] * N + ei[1])
    assert ref.numel() == nnz
    assert torch.equal(ref, key)
    tkey = indices.long() * N + rows                               # transpose is the same edge set
    assert int(tkey.sum()) == int(key.sum()) and int((tkey ^ (tkey >> 17)).sum()) == int((key ^ (key >> 17)).sum())
