"""torchrun --nproc-per-node W scripts/check_peer_exchange.py
grapes_allreduce_adam_peer (mean all-reduce + both Adam groups over NVLink peer memory) against NCCL all_reduce(AVG) +
grapes_adam_step2 on the same per-rank gradients: parameters agree to fp32 rounding (different summation order), and
are bit-identical ACROSS ranks with the peer kernel.  Several steps, so both slots and the sequence number are used;
once eagerly and once replayed from a CUDA graph."""
import os, sys
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from grapes_b200._lib import lib, ptr
from grapes_b200.dist import PeerGradExchange
from grapes_b200.utils import _any_ctx


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    L, ctx = lib(), _any_ctx(dev).ctx
    n, off0, n0, off1, n1 = 91_337, 0, 37_943, 37_943, 53_394
    g0 = torch.Generator(device=dev).manual_seed(7)
    params = torch.randn(n, generator=g0, device=dev)
    pe = PeerGradExchange(n, dev)
    dist.barrier()                                   # every rank's buffer is zeroed before anyone publishes
    st = torch.cuda.current_stream().cuda_stream
    gr = torch.Generator(device=dev).manual_seed(100 + rank)

    def fresh():
        return dict(p=params.clone(), m=torch.zeros(n, device=dev), v=torch.zeros(n, device=dev),
                    steps=torch.zeros(2, device=dev), ovf=torch.zeros(1, dtype=torch.int32, device=dev))
    A, B = fresh(), fresh()
    gbuf = torch.zeros(n, device=dev)

    def peer_step():
        L.grapes_allreduce_adam_peer(ctx, pe.peer_ptrs, rank, world, ptr(gbuf), n, ptr(B["p"]), ptr(gbuf), ptr(B["m"]),
                                     ptr(B["v"]), off0, n0, 1e-3, off1, n1, 1e-4, 0.9, 0.999, 1e-8, ptr(B["steps"]),
                                     ptr(pe.state), ptr(B["ovf"]), torch.cuda.current_stream().cuda_stream)

    graph = None
    worst = 0.0
    for step in range(6):
        g = torch.randn(n, generator=gr, device=dev) * (1.0 + step)
        ga = g.clone()
        dist.all_reduce(ga, op=dist.ReduceOp.AVG)
        L.grapes_adam_step2(ctx, ptr(A["p"]), ptr(ga), ptr(A["m"]), ptr(A["v"]), off0, n0, 1e-3, off1, n1, 1e-4, 0.9, 0.999,
                            1e-8, ptr(A["steps"]), st)
        gbuf.copy_(g)
        if step < 3:
            peer_step()
        else:                                   # the same two launches replayed from a CUDA graph
            if graph is None:
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                side = torch.cuda.Stream()
                with torch.cuda.graph(graph, stream=side):
                    peer_step()
            graph.replay()
        torch.cuda.synchronize()
        assert int(B["ovf"].item()) == 0, "peer timeout flag set"
        assert torch.allclose(gbuf, ga, rtol=1e-5, atol=1e-6), "mean gradient differs from NCCL AVG"
        err = ((A["p"] - B["p"]).abs().max() / A["p"].abs().max()).item()
        worst = max(worst, err)
        gathered = [torch.empty_like(B["p"]) for _ in range(world)]
        dist.all_gather(gathered, B["p"])
        for r in range(world):
            assert torch.equal(gathered[r], gathered[0]), f"step {step}: rank {r} parameters differ from rank 0"
    assert worst < 1e-5, worst
    assert float(B["steps"][0]) == 6.0 and float(B["steps"][1]) == 6.0
    # timing: peer exchange vs NCCL + adam, 200 iterations each
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier(); torch.cuda.synchronize()
    e0.record()
    for _ in range(200):
        graph.replay()
    e1.record(); torch.cuda.synchronize()
    t_peer = e0.elapsed_time(e1) / 200
    dist.barrier(); torch.cuda.synchronize()
    e0.record()
    for _ in range(200):
        dist.all_reduce(ga, op=dist.ReduceOp.AVG)
        L.grapes_adam_step2(ctx, ptr(A["p"]), ptr(ga), ptr(A["m"]), ptr(A["v"]), off0, n0, 1e-3, off1, n1, 1e-4, 0.9, 0.999,
                            1e-8, ptr(A["steps"]), st)
    e1.record(); torch.cuda.synchronize()
    t_nccl = e0.elapsed_time(e1) / 200
    if rank == 0:
        print(f"peer exchange ok: world={world} max rel param diff vs NCCL+adam {worst:.2e}; "
              f"{t_peer*1e3:.1f} us per exchange (graph replay) vs {t_nccl*1e3:.1f} us NCCL all_reduce + adam", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
