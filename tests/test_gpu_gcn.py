"""GCNConv / GCN parity: logits and gradients within
1e-5 relative (to the tensor's scale) of the fp64 restatement of PyG 2.5.2 GCNConv."""
import pytest
import torch

from oracle import reference_port as rp

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _close(got, ref, tol=TOL):
    ref = ref.double()
    scale = ref.abs().max().clamp_min(1e-30)
    err = (got.double().cpu() - ref).abs().max() / scale
    assert err < tol, f"relative error {err:.3e} >= {tol}"


def _rand_graph(n, E, gen):
    ei = torch.randint(0, n, (2, E), generator=gen)
    return ei


@pytest.mark.parametrize("n,E,dims", [(50, 200, [12, 32, 5]), (1661, 2071, [104, 256, 1]),
                                     (544, 700, [100, 256, 47]), (300, 5000, [33, 17, 3])])
def test_gcn_forward_backward(cuda_device, n, E, dims):
    from grapes_b200.gcn import GCN
    gen = torch.Generator().manual_seed(n)
    ei = _rand_graph(n, E, gen)
    ei[:, :5] = ei[0, :5]                    # a few explicit self-loops (must be replaced, not doubled)
    x = torch.randn(n, dims[0], generator=gen)
    ref = rp.OracleGCN(dims[0], dims[1:], generator=gen).double()
    net = GCN(dims[0], dims[1:]).to(cuda_device)
    net.load_state_dict({k: v.float() for k, v in ref.state_dict().items()})
    out_ref, _ = ref(x.double(), ei)
    w = torch.randn(out_ref.shape, generator=gen).double()
    (out_ref * w).sum().backward()
    out, mem = net(x.to(cuda_device), ei.to(cuda_device))
    assert isinstance(mem, float)
    (out * w.float().to(cuda_device)).sum().backward()
    _close(out, out_ref.detach())
    for (name, p), (_, pr) in zip(net.named_parameters(), ref.named_parameters()):
        _close(p.grad, pr.grad)


def test_gcn_layerwise_edge_list_quirk(cuda_device):
    """With a list of 3 blocks the hidden layer uses edges[-1] and the last layer edges[0]; edges[1] is
    never consumed (gcn.py:30-36, SURVEY.md section 3.1 'layer-index quirk')."""
    from grapes_b200.gcn import GCN
    gen = torch.Generator().manual_seed(0)
    n = 200
    edges = [_rand_graph(n, 300, gen) for _ in range(3)]
    x = torch.randn(n, 20, generator=gen)
    ref = rp.OracleGCN(20, [64, 7], generator=gen).double()
    net = GCN(20, [64, 7]).to(cuda_device)
    net.load_state_dict({k: v.float() for k, v in ref.state_dict().items()})
    out_ref, _ = ref(x.double(), edges)
    out, _ = net(x.to(cuda_device), [e.to(cuda_device) for e in edges])
    _close(out, out_ref.detach())
    edges2 = [edges[0], _rand_graph(n, 10, gen), edges[2]]
    out2, _ = net(x.to(cuda_device), [e.to(cuda_device) for e in edges2])
    assert torch.equal(out, out2)            # deterministic kernels: bitwise identical


def test_gcn_input_gradient(cuda_device):
    """dX path (needed by embed_nodes)."""
    from grapes_b200.gcn import GCN
    gen = torch.Generator().manual_seed(3)
    n = 120
    ei = _rand_graph(n, 500, gen)
    x = torch.randn(n, 40, generator=gen)
    ref = rp.OracleGCN(40, [16, 4], generator=gen).double()
    net = GCN(40, [16, 4]).to(cuda_device)
    net.load_state_dict({k: v.float() for k, v in ref.state_dict().items()})
    xr = x.double().requires_grad_(True)
    ref(xr, ei)[0].square().sum().backward()
    xg = x.to(cuda_device).requires_grad_(True)
    net(xg, ei.to(cuda_device))[0].square().sum().backward()
    _close(xg.grad, xr.grad)
