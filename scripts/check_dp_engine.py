"""torchrun --nproc-per-node W scripts/check_dp_engine.py
Data parallelism as the PRODUCT runs it (GrapesEngine.enable_data_parallel; the reference is single-process,
/root/reference/main.py:125-132,263-289 is the per-rank loop): every rank steps its own batch, the engine exchanges the
flat gradient and applies both Adam updates.  Checked for both exchanges ('peer': one kernel over NVLink peer memory,
'nccl': ncclAllReduce inside the step):
  * the gradient every rank holds after the exchange == mean over ranks of the float64 ORACLE gradients of the ranks'
    batches (1e-5 of the tensor's scale; fp32 SIMT path, same injected Gumbel noise as the oracle);
  * parameters are bit-identical on every rank after eager steps and after CUDA-graph replayed steps (Philox noise);
  * 'peer' and 'nccl' agree to fp32 rounding;
  * a rank whose peer never arrives raises GRAPES_OVF_PEER_TIMEOUT and its step is a no-op (parameters, moments and
    step counts untouched), and the failure is sticky."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from grapes_b200.dist import shard_batches                       # noqa: E402
from grapes_b200.engine import GrapesEngine                      # noqa: E402
from grapes_b200.graph import DeviceGraph                        # noqa: E402
from grapes_b200.synth import SHAPES, make_synth                 # noqa: E402
from oracle import reference_port as rp                          # noqa: E402  (test infrastructure)


def flat(named):
    return torch.cat([named[k].reshape(-1).double() for k in sorted(named)])


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    cfg = SHAPES["small"]
    d = make_synth("small", seed=3)
    B, k, H = cfg["batch_size"], cfg["num_samples"], cfg["sampling_hops"]
    tr = d.train_mask.nonzero().squeeze(1)
    mine = shard_batches(tr.numel() // B, rank, world)
    g = DeviceGraph.from_edge_index(d.edge_index, d.num_nodes, device=dev)

    def engine(exchange):
        st = rp.OracleState(d, sampling_hops=H, num_samples=k, seed=103, dtype=torch.float64)
        e = GrapesEngine(g, d.x.to(dev), d.y.to(dev), num_classes=d.num_classes, batch_size=B, num_samples=k,
                         sampling_hops=H, seed=5, use_tensor_cores=False)
        e.load_state_dicts(gcn_c=st.gcn_c.state_dict(), gcn_gf=st.gcn_gf.state_dict(), gcn_z=st.gcn_z.state_dict())
        e.enable_data_parallel(exchange=exchange)
        return st, e

    finals = {}
    for exchange in ("peer", "nccl"):
        st, eng = engine(exchange)
        # ---- step 0, eager, injected noise: exchanged gradient == mean of the per-rank oracle gradients ----
        torch.manual_seed(1000 + rank)
        t0 = tr[mine[0] * B:(mine[0] + 1) * B]
        ref = rp.reference_step(st, t0, apply_optim=False)
        noise = [None if h["noise"] is None else h["noise"].float().to(dev) for h in ref["hops"]]
        rec = eng.step(t0.to(dev), gumbel_noise=noise, apply_optim=True, record=True)
        eng.check_overflow()
        for a, b in zip(rec["hops"], ref["hops"]):
            assert torch.equal(a["sampled"].cpu().long(), b["sampled"]), "sampled set differs from the oracle"
        mine_ref = torch.cat([flat(ref["grads_c"]), flat(ref["grads_gf"]), flat(ref["grads_z"])]).to(dev)
        allref = [torch.empty_like(mine_ref) for _ in range(world)]
        dist.all_gather(allref, mine_ref)
        mean_ref = torch.stack(allref).mean(0)
        got = torch.cat([flat(rec["grads"]["gcn_c"]), flat(rec["grads"]["gcn_gf"]), flat(rec["grads"]["gcn_z"])])
        off = 0
        for key in ("gcn_c", "gcn_gf", "gcn_z"):
            for name in sorted(rec["grads"][key]):
                n = rec["grads"][key][name].numel()
                a_, b_ = got[off:off + n], mean_ref[off:off + n]
                err = ((a_ - b_).abs().max() / b_.abs().max().clamp_min(1e-30)).item()
                assert err < 1e-5, f"{exchange}: exchanged grad {key}.{name} vs mean of oracle grads: {err:.2e}"
                off += n
        assert not torch.equal(allref[0], allref[-1]), "ranks ran the same batch"

        def same_everywhere(tag):
            ps = [torch.empty_like(eng.params) for _ in range(world)]
            dist.all_gather(ps, eng.params)
            for r in range(world):
                assert torch.equal(ps[r], ps[0]), f"{exchange} {tag}: rank {r} parameters differ from rank 0"
            ms = [torch.empty_like(eng.exp_avg) for _ in range(world)]
            dist.all_gather(ms, eng.exp_avg)
            assert all(torch.equal(m_, ms[0]) for m_ in ms), f"{exchange} {tag}: Adam moments differ across ranks"
        same_everywhere("eager step")
        # ---- graph-replayed steps with cross-step prefetch, device-side noise ----
        bl = [tr[b * B:(b + 1) * B].to(dev).to(torch.int32).contiguous() for b in mine[1:7]]
        for j, b in enumerate(bl):
            eng.step(b, use_graph=True, next_targets=bl[j + 1] if j + 1 < len(bl) else None)
        torch.cuda.synchronize()
        eng.check_overflow()
        eng.raise_on_flags(eng.scalars()["flags"])
        assert float(eng.adam_steps[0]) == 7.0 and float(eng.adam_steps[1]) == 7.0
        same_everywhere("graph replay")
        finals[exchange] = eng.params.clone()
        if exchange == "peer":
            keep = (st, eng)
    # the two exchanges add the ranks' gradients in different orders: a last-bit difference, which Adam's first updates
    # (~ lr * sign(g)) turn into a 2*lr flip on entries whose gradient is zero up to rounding -- bounded, and rare
    diff = (finals["peer"] - finals["nccl"]).abs()
    assert diff.max().item() <= 2.0 * 1e-3 * 7 * 1.01, f"peer vs nccl parameters: {diff.max().item():.3e} beyond the sign-flip bound"
    rel = (diff > 1e-4 * finals["nccl"].abs().max()).double().mean().item()
    assert rel < 0.02, f"{rel:.4f} of the parameters differ by > 1e-4 between the peer and the NCCL exchange"

    # ---- a peer that never arrives: flagged, no-op, sticky (rank 0 steps alone) ----
    st, eng = keep
    dist.barrier()
    if rank == 0:
        before = (eng.params.clone(), eng.exp_avg.clone(), eng.exp_avg_sq.clone(), eng.adam_steps.clone())
        for _ in range(2):
            eng.step(tr[mine[0] * B:(mine[0] + 1) * B].to(dev), apply_optim=True)
            torch.cuda.synchronize()
            assert int(eng.scalars()["flags"]) & 32, "timeout not reported in the step's flag word"
            for a_, b_ in zip(before, (eng.params, eng.exp_avg, eng.exp_avg_sq, eng.adam_steps)):
                assert torch.equal(a_, b_), "a failed exchange must leave parameters, moments and step counts untouched"
        try:
            eng.check_overflow()
            raise AssertionError("check_overflow did not raise after a peer timeout")
        except RuntimeError as e:
            assert "exchange failed" in str(e)
    dist.barrier()
    if rank == 0:
        print(f"dp engine ok: world={world}; parameters off by > 1e-4 between peer and NCCL exchange: {rel:.4f}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    try:
        main()
    except BaseException:
        import traceback
        print("RANK", os.environ.get("RANK"), "FAILED:\n" + traceback.format_exc(), flush=True)
        raise
