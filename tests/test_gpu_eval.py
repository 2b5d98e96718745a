"""Evaluation parity: /root/reference/eval.py restated in
oracle/reference_port.py::reference_evaluate vs grapes_b200.eval.evaluate on the device.
Integer outputs (sampled sets, blocks, predictions) bit-exact; logits within 1e-5 of the fp64 oracle."""
import types

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


def _setup(name, seed, dev, **over):
    from grapes_b200.gcn import GCN
    from grapes_b200.graph import DeviceGraph
    from grapes_b200.synth import make_synth, SHAPES
    cfg = dict(SHAPES[name]); cfg.update(over)
    d = make_synth(name, seed=seed, **over)
    st = rp.OracleState(d, sampling_hops=cfg["sampling_hops"], num_samples=cfg["num_samples"], seed=seed + 7,
                        dtype=torch.float64)
    g = DeviceGraph.from_edge_index(d.edge_index, d.num_nodes, device=dev)
    gcn_c = GCN(d.num_features, [256, d.num_classes]).to(dev)
    gcn_c.load_state_dict({k: v.float() for k, v in st.gcn_c.state_dict().items()})
    gcn_gf = GCN(d.num_features + st.num_indicators, [256, 1]).to(dev)
    gcn_gf.load_state_dict({k: v.float() for k, v in st.gcn_gf.state_dict().items()})
    args = types.SimpleNamespace(sampling_hops=cfg["sampling_hops"], num_samples=cfg["num_samples"], use_indicators=True)
    return cfg, d, st, g, gcn_c, gcn_gf, args


@pytest.mark.parametrize("name,seed", [("tiny", 0), ("cora", 0), ("small", 1)])
def test_full_batch_eval_matches_oracle(cuda_device, name, seed):
    from grapes_b200.eval import evaluate
    cfg, d, st, g, gcn_c, gcn_gf, args = _setup(name, seed, cuda_device)
    ref = rp.reference_evaluate(st, d.test_mask, full_batch=True)
    acc, f1, logits = evaluate(gcn_c, gcn_gf, d, args, g, None, st.num_indicators, cuda_device, mask=d.test_mask,
                               full_batch=True, return_predictions=True)
    _close(logits, ref["logits"])
    pred = torch.argmax(logits, dim=1).cpu()[d.test_mask]
    # argmax may only differ where the oracle's top two logits are closer than the float tolerance
    top2 = ref["logits"][d.test_mask].topk(2, dim=1).values
    safe = (top2[:, 0] - top2[:, 1]) > 1e-4 * ref["logits"].abs().max()
    assert torch.equal(pred[safe], ref["predictions"][safe])
    assert abs(acc - ref["accuracy"]) <= (~safe).sum().item() / max(int(d.test_mask.sum()), 1) + 1e-6
    assert acc == f1


def test_full_batch_eval_multilabel(cuda_device):
    from grapes_b200.eval import evaluate
    cfg, d, st, g, gcn_c, gcn_gf, args = _setup("tiny", 3, cuda_device, multilabel=True)
    ref = rp.reference_evaluate(st, d.val_mask, full_batch=True)
    acc, f1 = evaluate(gcn_c, gcn_gf, d, args, g, None, st.num_indicators, cuda_device, mask=d.val_mask, full_batch=True)
    assert abs(f1 - ref["f1"]) < 1e-6 and acc == f1


def test_graph_norm_matches_edge_list_structure(cuda_device):
    """GraphNorm (whole CSR, TMA-staged kernel at n >= 4096) == NormAdj(edge_index) bit for bit, and both equal the
    register-staged kernel (grapes_agg_variant 1)."""
    from grapes_b200._lib import lib
    from grapes_b200.gcn import GraphNorm, NormAdj
    from grapes_b200.graph import DeviceGraph
    from grapes_b200.synth import make_synth
    d = make_synth("small", seed=2, power_law=1.0)
    ei = d.edge_index.clone()
    ei[:, :7] = ei[0, :7]                                      # stored self-loops: dropped, then one added per node
    g = DeviceGraph.from_edge_index(ei, d.num_nodes, device=cuda_device)
    x = d.x.to(cuda_device)
    gn = GraphNorm(g)
    uniq = torch.unique(ei[0] * d.num_nodes + ei[1])
    ei_u = torch.stack([uniq // d.num_nodes, uniq % d.num_nodes]).to(cuda_device)
    na = NormAdj(ei_u, d.num_nodes)
    assert torch.equal(gn.in_off.cpu()[: d.num_nodes + 1], na.in_off.cpu()[: d.num_nodes + 1])
    assert torch.equal(gn.in_src.cpu(), na.in_src.cpu()[: gn.in_src.numel()])
    assert torch.equal(gn.dinv.cpu(), na.dinv.cpu()[: d.num_nodes])
    L = lib()
    try:
        y_tma = gn.aggregate(x)
        L.cdll.grapes_agg_variant(1)
        y_reg = gn.aggregate(x)
        y_na = na.aggregate(x)
    finally:
        L.cdll.grapes_agg_variant(0)
    assert torch.equal(y_tma, y_reg) and torch.equal(y_tma, y_na)
    _close(y_tma, rp.gcn_conv(d.x.double(), ei_u.cpu(), torch.eye(d.num_features, dtype=torch.float64), None))


@pytest.mark.parametrize("name,seed,bs", [("tiny", 0, 32), ("tiny", 1, 50), ("cora", 0, 512)])
def test_mini_batch_eval_matches_oracle(cuda_device, name, seed, bs):
    from grapes_b200.eval import evaluate
    cfg, d, st, g, gcn_c, gcn_gf, args = _setup(name, seed, cuda_device)
    mask = d.val_mask
    ref = rp.reference_evaluate(st, mask, full_batch=False, batch_size=bs)
    idx = mask.nonzero().squeeze(1)
    loader = [(b,) for b in torch.split(idx, bs)]
    acc, f1, pred = evaluate(gcn_c, gcn_gf, d, args, g, None, st.num_indicators, cuda_device, mask=mask,
                             loader=loader, full_batch=False, return_predictions=True)
    assert torch.equal(pred.cpu(), ref["predictions"])
    assert abs(acc - ref["accuracy"]) < 1e-6 and acc == f1


@pytest.mark.parametrize("name,seed,bs,tc", [("tiny", 0, 32, False), ("small", 1, 128, True), ("small", 2, 100, True)])
def test_mini_batch_eval_engine_path_bit_exact_per_hop(cuda_device, name, seed, bs, tc):
    """The evaluator's batch body on the engine (GrapesEngine.predict, eval.py:84-153): per hop the candidate list, the
    deterministic top-k set and the block in the EVALUATION direction (rows = previous_nodes, cols = targets u sampled,
    eval.py:140-142) bit-exact against the oracle; classifier logits at 1e-5; eager launches == graph replay."""
    from grapes_b200.engine import GrapesEngine
    cfg, d, st, g, gcn_c, gcn_gf, args = _setup(name, seed, cuda_device)
    dev = cuda_device
    mask = d.val_mask
    ref = rp.reference_evaluate(st, mask, full_batch=False, batch_size=bs)
    eng = GrapesEngine(g, d.x.to(dev), d.y.to(dev), num_classes=d.num_classes, batch_size=bs,
                       num_samples=cfg["num_samples"], sampling_hops=cfg["sampling_hops"], use_tensor_cores=tc)
    eng.load_state_dicts(gcn_c=st.gcn_c.state_dict(), gcn_gf=st.gcn_gf.state_dict())
    idx = mask.nonzero().squeeze(1)
    out = torch.zeros(bs, dtype=torch.int32, device=dev)
    out_g = torch.zeros(bs, dtype=torch.int32, device=dev)
    for bi, t in enumerate(torch.split(idx, bs)):
        b = int(t.numel())
        eng.predict(t.to(dev), out, use_graph=False)
        torch.cuda.synchronize()
        eng.check_overflow()
        rb = ref["batches"][bi]
        sizes = eng.hop_sizes()
        for h, (sz, rh) in enumerate(zip(sizes, rb["hops"])):
            hw = eng.hops[h]
            assert torch.equal(hw.nb_nodes[:sz["c"]].cpu().long(), rh["neighbor_nodes"]), f"batch {bi} hop {h}: candidates"
            assert torch.equal(eng.prev[h + 1][b:b + sz["s"]].cpu().long(), rh["sampled"]), f"batch {bi} hop {h}: top-k set"
            blk = torch.stack([hw.blk_src[:sz["blk"]], hw.blk_dst[:sz["blk"]]]).cpu().long()
            assert torch.equal(blk, rh["block_edges"]), f"batch {bi} hop {h}: block (rows = previous_nodes)"
        A = eng.count("A")
        assert torch.equal(eng.all_nodes[:A].cpu().long(), rb["all_nodes"])
        _close(eng.logits_c[:A], rb["logits"])
        eng.predict(t.to(dev), out_g, use_graph=True)
        torch.cuda.synchronize()
        assert torch.equal(out[:b], out_g[:b]), "graph replay differs from eager launches"


def test_train_mini_batch_eval_uses_configured_batch_size(cuda_device):
    """main.py:127-132 builds val_loader / test_loader with args.batch_size -- NOT the size of the last (partial) training
    batch.  After training on a split that is not a multiple of batch_size, the mini-batch evaluation of train() must
    equal the oracle's evaluation of the SAME weights with batch_size = args.batch_size."""
    from grapes_b200.args import Arguments
    from grapes_b200.synth import make_synth
    from grapes_b200 import train as T
    d = make_synth("tiny", seed=5)
    assert int(d.train_mask.sum()) % 32 != 0
    args = Arguments.parse_args(["--dataset", "tiny", "--batch_size", "32", "--num_samples", "8", "--sampling_hops", "2",
                                   "--max_epochs", "1", "--eval_frequency", "5", "--eval_full_batch", "False",
                                   "--eval_on_cpu", "False"])
    f1, *_ = T.train(args, data=d, device=cuda_device)
    eng = T.train.last_engine
    assert eng.bsz != 32 and eng.B == 32                      # the last training batch was partial
    st = rp.OracleState(d, sampling_hops=2, num_samples=8, seed=0, dtype=torch.float64)
    sd = {k: {n: t.detach().cpu().double() for n, t in v.items()} for k, v in eng.state_dicts().items()}
    st.gcn_c.load_state_dict(sd["gcn_c"]); st.gcn_gf.load_state_dict(sd["gcn_gf"])
    ref = rp.reference_evaluate(st, d.test_mask, full_batch=False, batch_size=32)
    assert abs(f1 - ref["f1"]) < 1e-6


@pytest.mark.skipif(__import__("os").environ.get("GRAPES_TEST_UNVERIFIED", "0") != "1",
                    reason="written after the round's GPU budget was spent; opt in with GRAPES_TEST_UNVERIFIED=1 "
                           "(the same path is measured and cross-checked by scripts/bench_full_eval.py in its own process)")
@pytest.mark.parametrize("name,seed", [("small", 1), ("small", 4)])
def test_full_graph_forward_tensor_core_path(cuda_device, name, seed):
    """gcn.full_graph_forward (GraphNorm built by the library's kernels, hidden layer on grapes_gemm_bias_relu_tc) against
    the float64 oracle at 1e-5 and the structure against the torch-sort builder bit for bit."""
    from grapes_b200.gcn import GraphNorm, full_graph_forward
    cfg, d, st, g, gcn_c, gcn_gf, args = _setup(name, seed, cuda_device)
    gl, gt = GraphNorm(g, builder="lib"), GraphNorm(g, builder="torch")
    N = d.num_nodes
    assert torch.equal(gl.in_off[:N + 1], gt.in_off[:N + 1]) and torch.equal(gl.in_src, gt.in_src)
    assert torch.equal(gl.dinv[:N], gt.dinv[:N])
    ge = GraphNorm(g, edge_index=d.edge_index.to(cuda_device), builder="lib")       # duplicates kept, self-loops dropped
    ref = rp.reference_evaluate(st, d.test_mask, full_batch=True)
    logits = full_graph_forward(gcn_c.eval(), d.x.to(cuda_device), ge)
    _close(logits, ref["logits"])
