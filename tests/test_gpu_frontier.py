"""Bit-exact parity of the CUDA graph ops with the CPU oracle (integer work: no tolerance).
Covers SURVEY.md section 8 rows a2 (get_neighborhoods), a3 (mask dedup), a4 (TensorMap), a10 (slice)."""
import numpy as np
import pytest
import torch

from grapes_b200.synth import make_synth
from oracle import reference_port as rp

pytestmark = pytest.mark.gpu


def _graph(name, seed, device, **kw):
    from grapes_b200.graph import DeviceGraph
    d = make_synth(name, seed=seed, **kw)
    adj = rp.build_adjacency(d.edge_index, d.num_nodes)
    g = DeviceGraph.from_edge_index(d.edge_index, d.num_nodes, device=device)
    return d, adj, g


@pytest.mark.parametrize("name,seed", [("tiny", 0), ("cora", 0), ("small", 1)])
def test_csr_matches_scipy(cuda_device, name, seed):
    d, adj, g = _graph(name, seed, cuda_device)
    assert g.nnz == adj.nnz
    assert np.array_equal(g.indptr.cpu().numpy(), adj.indptr.astype(np.int64))
    assert np.array_equal(g.indices.cpu().numpy(), adj.indices.astype(np.int32))


@pytest.mark.parametrize("name,seed,P", [("tiny", 0, 17), ("cora", 0, 512), ("cora", 3, 1), ("small", 1, 700)])
def test_get_neighborhoods_bit_exact(cuda_device, name, seed, P):
    from grapes_b200.utils import get_neighborhoods
    d, adj, g = _graph(name, seed, cuda_device)
    gen = torch.Generator().manual_seed(seed)
    nodes = torch.randperm(d.num_nodes, generator=gen)[:P]
    ref = rp.get_neighborhoods(nodes, adj)
    got = get_neighborhoods(nodes, g)
    assert got.dtype == torch.int64 and got.shape == ref.shape
    assert torch.equal(got.cpu(), ref)


def test_get_neighborhoods_empty_and_isolated(cuda_device):
    from grapes_b200.graph import DeviceGraph
    from grapes_b200.utils import get_neighborhoods
    # node 3 and 5 have no out-edges; node 0 has a self-loop
    ei = torch.tensor([[0, 0, 1, 2, 4, 4], [0, 2, 0, 1, 1, 2]])
    adj = rp.build_adjacency(ei, 6)
    g = DeviceGraph.from_edge_index(ei, 6, device=cuda_device)
    for nodes in (torch.tensor([3, 5]), torch.tensor([0, 3, 4]), torch.tensor([], dtype=torch.long), torch.tensor([4, 0, 1, 2])):
        ref = rp.get_neighborhoods(nodes, adj)
        got = get_neighborhoods(nodes, g).cpu()
        assert got.shape == ref.shape and torch.equal(got, ref)


def test_hub_rows_power_law(cuda_device):
    """Zipf-like degree law: a few rows hold most edges (SURVEY.md 7.2 'hub rows')."""
    from grapes_b200.utils import get_neighborhoods, slice_adjacency
    d, adj, g = _graph("small", 2, cuda_device, power_law=1.5)
    deg = np.diff(adj.indptr)
    assert deg.max() > 20 * max(deg.mean(), 1)
    hubs = torch.from_numpy(np.argsort(-deg)[:40].copy()).long()
    ref = rp.get_neighborhoods(hubs, adj)
    assert torch.equal(get_neighborhoods(hubs, g).cpu(), ref)
    cols = torch.randperm(d.num_nodes, generator=torch.Generator().manual_seed(0))[:900]
    assert torch.equal(slice_adjacency(g, hubs, cols).cpu(), rp.slice_adjacency(adj, hubs, cols))


@pytest.mark.parametrize("name,seed,R,Cn", [("tiny", 0, 20, 30), ("cora", 0, 528, 512), ("small", 1, 160, 128)])
def test_slice_adjacency_bit_exact(cuda_device, name, seed, R, Cn):
    from grapes_b200.utils import slice_adjacency
    d, adj, g = _graph(name, seed, cuda_device)
    gen = torch.Generator().manual_seed(seed + 10)
    perm = torch.randperm(d.num_nodes, generator=gen)
    cols = perm[:Cn]
    rows = torch.cat([cols[: R // 2], perm[Cn:Cn + R - R // 2]])       # overlap like T u S vs prev
    ref = rp.slice_adjacency(adj, rows, cols)
    got = slice_adjacency(g, rows, cols).cpu()
    assert got.shape == ref.shape and torch.equal(got, ref)


def test_tensormap_docstring_vector(cuda_device):
    """The only known-answer vector in the reference: utils.py:104-108."""
    from grapes_b200.utils import TensorMap
    nodes = torch.tensor([22, 32, 42, 52], device=cuda_device)
    node_map = TensorMap(size=nodes.max() + 1, device=cuda_device)
    node_map.update(nodes)
    assert node_map.map(torch.tensor([52, 42, 32, 22, 22])).cpu().tolist() == [3, 2, 1, 0, 0]
