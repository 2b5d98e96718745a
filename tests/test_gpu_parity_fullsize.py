"""Float parity WHERE THE NUMBER IS QUOTED: one whole training step of the DEFAULT engine (tcgen05 3xTF32 GEMMs, TMA
aggregation, multi-stream step) on the BASELINE.json graphs bench.py runs -- products-shaped (2.45 M nodes / 123.7 M nnz,
B 1024, 3 hops: the headline config), arxiv-shaped, Reddit-shaped (F = 602, ~490-entry rows) -- against the float64 CPU
oracle (oracle/reference_port.py, restatement of /root/reference/main.py:161-291) run on the SAME graph: the scipy CSR is
rebuilt from the device indptr / indices exactly as bench.py's cpu_baseline leg does.

Checked by tests/parity_utils.py::check_step: every integer contract bit-exact per hop (frontier, dedup, relabel, sampled
set, blocks, all_nodes), aggregated features Y, sampler logits, log-probs, classifier logits, loss_c, log_z, loss_gfn and
all three gradient sets at 1e-5 of the tensor's scale (gradients: or 2x the reference's own fp32 deviation), with relu-kink
flips of the tensor-core path verified one by one against the oracle's float64 pre-activations.

The last case is the papers100M-shaped STORAGE format (bf16 feature table, int64 indptr past 2^31) on a graph small enough
to build in seconds: targets are rows whose CSR offsets lie beyond 2^31; the oracle runs on the order-preserving compaction
of the touched sub-problem (ids relabelled monotonically, so every ascending-id contract maps back one to one)."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from oracle import reference_port as rp
from parity_utils import check_step, well_separated

pytestmark = pytest.mark.gpu


def _oracle_pair(data, adj, cfg, seed):
    kw = dict(sampling_hops=cfg["sampling_hops"], num_samples=cfg["num_samples"], hidden_dim=256, seed=seed,
              adjacency=adj)
    st = rp.OracleState(data, dtype=torch.float64, **kw)
    st.fp32 = rp.OracleState(data, dtype=torch.float32, **kw)
    return st


@pytest.mark.parametrize("name,batch", [("arxiv", 5), ("products", 3), ("reddit", 2)])
def test_full_size_float_parity_default_path(cuda_device, name, batch):
    from bench import build_workload
    from grapes_b200.engine import GrapesEngine
    from grapes_b200.graph import DeviceGraph
    from grapes_b200.synth import SynthData
    dev = cuda_device
    torch.manual_seed(4242)                     # the oracle draws its Gumbel noise from the global CPU generator
    cfg, indptr, indices, x, y, train_idx = build_workload(name, 0, dev, native_csr=True)
    cfg.pop("csr_build", None)
    N, B = cfg["N"], cfg["batch_size"]
    ip, ix = indptr.cpu().numpy(), indices.cpu().numpy()
    adj = sp.csr_matrix((np.ones(ix.shape[0], dtype=bool), ix, ip), shape=(N, N))       # as bench.py cpu_reference_run
    data = SynthData(x=x.cpu(), y=y.cpu(), edge_index=torch.zeros(2, 0, dtype=torch.long), train_mask=None,
                     val_mask=None, test_mask=None, num_nodes=N, num_features=cfg["F"], num_classes=cfg["C"])
    st = _oracle_pair(data, adj, cfg, seed=100)
    eng = GrapesEngine(DeviceGraph(indptr, indices, N), x, y, num_classes=cfg["C"], batch_size=B,
                       num_samples=cfg["num_samples"], sampling_hops=cfg["sampling_hops"], hidden_dim=256, seed=0)
    assert eng.use_tc and eng.use_tc_bwd and eng.multi_stream, "this test pins the DEFAULT path"
    eng.load_state_dicts(gcn_c=st.gcn_c.state_dict(), gcn_gf=st.gcn_gf.state_dict(), gcn_z=st.gcn_z.state_dict())
    # a draw whose top-k boundary is not within fp32 rounding of a tie (else the bit-exact set claim is void): the
    # oracle alone decides, before the device is asked anything
    for attempt in range(6):
        targets = train_idx[(batch + attempt) * B:(batch + attempt + 1) * B].cpu()
        pre = rp.reference_step(st, targets, apply_optim=False)
        noise = [h["noise"] for h in pre["hops"]]
        if well_separated(pre, rp.reference_step(st.fp32, targets, gumbel_noise=noise, apply_optim=False)):
            break
    else:
        pytest.fail("no well-separated draw in 6 batches")
    rec, ref = check_step(st, eng, targets, dev, apply_optim=False, gumbel_noise=noise)
    sizes = [h["n"] for h in rec["hops"]]
    assert min(sizes) > 10 * B, sizes                                    # a real frontier, not a toy
    print(f"{name}: frontier rows per hop {sizes}, loss_c {rec['scalars']['loss_c']:.6f} "
          f"(oracle {ref['loss_c'].item():.6f}), loss_gfn {rec['scalars']['loss_gfn']:.6e} "
          f"(oracle {ref['loss_gfn'].item():.6e}), verified relu-kink units {rec.get('relu_flips')}")


def _regular_big_csr(N, deg, dev, chunk=1 << 22):
    """Directed graph with `deg` distinct ascending neighbours per row, N * deg > 2^31 stored entries, built with
    arithmetic only: row i -> base_i + j * S (j < deg), base_i a hash of i.  Canonical CSR by construction."""
    S = 479_001
    span = N - (deg - 1) * S
    assert span > 0
    indices = torch.empty(N * deg, dtype=torch.int32, device=dev)
    j = torch.arange(deg, device=dev, dtype=torch.int64) * S
    for r0 in range(0, N, chunk):
        r1 = min(N, r0 + chunk)
        i = torch.arange(r0, r1, device=dev, dtype=torch.int64)
        base = (i * 2654435761 + 12345) % span
        indices[r0 * deg:r1 * deg] = (base.unsqueeze(1) + j.unsqueeze(0)).reshape(-1).to(torch.int32)
    indptr = torch.arange(N + 1, device=dev, dtype=torch.int64) * deg
    return indptr, indices


def test_bf16_table_int64_offsets_past_2_31(cuda_device):
    from grapes_b200.engine import GrapesEngine
    from grapes_b200.graph import DeviceGraph
    from grapes_b200.synth import SynthData
    dev = cuda_device
    torch.manual_seed(777)
    N, deg, F, C, B, k, H = 1 << 25, 66, 128, 19, 512, 256, 2
    indptr, indices = _regular_big_csr(N, deg, dev)
    assert int(indptr[-1]) > (1 << 31)
    g = torch.Generator(device=dev).manual_seed(3)
    x = torch.empty(N, F, dtype=torch.bfloat16, device=dev)
    for r0 in range(0, N, 1 << 22):
        r1 = min(N, r0 + (1 << 22))
        x[r0:r1] = torch.randn(r1 - r0, F, generator=g, device=dev).to(torch.bfloat16)
    y = torch.randint(0, C, (N,), generator=g, device=dev)
    # targets: rows whose CSR offsets lie beyond 2^31
    first_high = (1 << 31) // deg + 1
    targets = torch.sort(torch.randperm(N - first_high, generator=g, device=dev)[:B] + first_high).values
    assert int(indptr[targets[0]]) > (1 << 31)
    eng = GrapesEngine(DeviceGraph(indptr, indices, N), x, y, num_classes=C, batch_size=B, num_samples=k,
                       sampling_hops=H, hidden_dim=256, seed=0)
    assert eng.x_bf16 and eng.use_tc
    cfg = dict(sampling_hops=H, num_samples=k)
    # the oracle's weights depend on the seed and the widths only: take them from a one-node problem first
    one = SynthData(x=torch.zeros(1, F), y=torch.zeros(1, dtype=torch.long), edge_index=torch.zeros(2, 0, dtype=torch.long),
                    train_mask=None, val_mask=None, test_mask=None, num_nodes=1, num_features=F, num_classes=C)
    w0 = _oracle_pair(one, sp.csr_matrix((1, 1), dtype=bool), cfg, seed=100)
    eng.load_state_dicts(gcn_c=w0.gcn_c.state_dict(), gcn_gf=w0.gcn_gf.state_dict(), gcn_z=w0.gcn_z.state_dict())
    for attempt in range(6):
        # device pass with injected noise: which nodes does the step touch?
        u = torch.rand(H, eng.cap_n, generator=g, device=dev).clamp_(1e-7, 1 - 1e-7)
        noise = [(-torch.log(-torch.log(u[h]))).contiguous() for h in range(H)]
        rec = eng.step(targets, gumbel_noise=noise, apply_optim=False, record=True)
        eng.check_overflow()
        # every row the step expands (hop rows + the last block's rows) with its TRUE neighbour list from the device CSR
        rows = torch.unique(torch.cat([h["prev"].long() for h in rec["hops"]] +
                                      [torch.cat([targets, rec["hops"][-1]["sampled"].long()])]))
        beg = indptr[rows]
        cnt = indptr[rows + 1] - beg
        pos = torch.repeat_interleave(torch.arange(rows.numel(), device=dev), cnt)
        off = torch.arange(int(cnt.sum()), device=dev) - (torch.cumsum(cnt, 0) - cnt)[pos]
        nb = indices[beg[pos] + off].long()
        # order-preserving compaction: U ascending, new id = rank in U (every ascending-id contract maps back 1:1)
        U = torch.unique(torch.cat([targets, rows, nb]))
        n_sub = int(U.numel())
        adj = sp.csr_matrix((np.ones(nb.numel(), dtype=bool),
                             (torch.searchsorted(U, rows)[pos].cpu().numpy(), torch.searchsorted(U, nb).cpu().numpy())),
                            shape=(n_sub, n_sub))
        data = SynthData(x=x[U].float().cpu(), y=y[U].cpu(), edge_index=torch.zeros(2, 0, dtype=torch.long),
                         train_mask=None, val_mask=None, test_mask=None, num_nodes=n_sub, num_features=F, num_classes=C)
        st = _oracle_pair(data, adj, cfg, seed=100)
        sub_targets = torch.searchsorted(U, targets).cpu()
        # the oracle gets the device pass's noise by candidate position (identical candidate lists <=> identical order);
        # if its frontier differed from the device's, it would run into rows the sub-problem does not hold and mismatch
        cs = [h["c"] for h in rec["hops"]]
        noise_cpu = [noise[h][:cs[h]].double().cpu() for h in range(H)]
        ref = rp.reference_step(st, sub_targets, gumbel_noise=noise_cpu, apply_optim=False)
        ref32 = rp.reference_step(st.fp32, sub_targets, gumbel_noise=noise_cpu, apply_optim=False)
        if well_separated(ref, ref32):
            break
    else:
        pytest.fail("no well-separated draw in 6 attempts")
    Uc = U.cpu()
    from parity_utils import _grad_ok, _rel, relu_flips
    for h, (a, b) in enumerate(zip(rec["hops"], ref["hops"])):
        assert b["neighbor_nodes"].numel() == cs[h], f"hop {h}: candidate count"
        assert torch.equal(a["batch_nodes"].cpu().long(), Uc[b["batch_nodes"]]), f"hop {h} batch_nodes"
        assert torch.equal(a["neighbor_nodes"].cpu().long(), Uc[b["neighbor_nodes"]]), f"hop {h} neighbor_nodes"
        assert torch.equal(torch.stack([a["e_src"], a["e_dst"]]).cpu().long(), b["local_neighborhoods"]), f"hop {h} local edges"
        assert torch.equal(a["sampled"].cpu().long(), Uc[b["sampled"]]), f"hop {h} sampled set"
        assert torch.equal(a["block_edges"].cpu().long(), Uc[b["block_edges"]]), f"hop {h} block"
        assert _rel(a["logits_all"], b["logits_all"]) < 1e-5, f"hop {h} logits"
        assert _rel(a["log_prob"], b["log_prob"]) < 1e-5, f"hop {h} log_prob"
    assert torch.equal(rec["all_nodes"].cpu().long(), Uc[ref["all_nodes"]])
    assert _rel(rec["logits_c"], ref["logits_c"]) < 1e-5
    s = rec["scalars"]
    assert abs(s["loss_c"] - ref["loss_c"].item()) < 1e-5 * abs(ref["loss_c"].item())
    assert abs(s["log_z"] - ref["log_z"].item()) < 1e-5 * max(1.0, abs(ref["log_z"].item()))
    assert abs(s["loss_gfn"] - ref["loss_gfn"].item()) < 4e-5 * abs(ref["loss_gfn"].item())
    before = {k_: {n: p.detach().clone() for n, p in net.named_parameters()}
              for k_, net in (("gcn_gf", st.gcn_gf), ("gcn_z", st.gcn_z))}
    flips = relu_flips(eng, st, ref, before)
    for key, rkey in (("gcn_c", "grads_c"), ("gcn_gf", "grads_gf"), ("gcn_z", "grads_z")):
        for name, gref in ref[rkey].items():
            assert _grad_ok(rec["grads"][key][name], gref, ref32[rkey][name], flips.get(key, ())), f"grad {key} {name}"
