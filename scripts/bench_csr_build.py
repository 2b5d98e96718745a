"""Times grapes_csr_from_edges (main.py:134-136) on a synthetic shape: the C-ABI call alone
on caller-owned buffers, CUDA events on the launching stream, median of 5, for the shape-chosen form and both grouping
variants, next to the torch sort/unique formulation it replaced.  Prints one JSON line."""
import json
import sys

import torch

sys.path.insert(0, ".")
from grapes_b200._lib import lib, ptr                         # noqa: E402
from grapes_b200.graph import csr_from_edge_index            # noqa: E402
from grapes_b200.synth import SHAPES, synth_edge_index       # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "products"
cfg = SHAPES[name]
dev = torch.device("cuda:0")
N, E = cfg["N"], cfg["E_dir"]
ei = synth_edge_index(N, E, 0, device=dev)
src, dst = ei[0].contiguous(), ei[1].contiguous()
L = lib()
ws_bytes = int(L.cdll.grapes_csr_workspace_bytes(N, E))
ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
indptr = torch.empty(N + 1, dtype=torch.int64, device=dev)
cap = torch.empty(E, dtype=torch.int32, device=dev)
meta = torch.zeros(2, dtype=torch.int64, device=dev)
st = torch.cuda.current_stream().cuda_stream


def build():
    L.grapes_csr_from_edges(ptr(src), ptr(dst), E, N, ptr(indptr), ptr(cap), meta.data_ptr(), meta.data_ptr() + 8, ptr(ws),
                            ws_bytes, st)


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def torch_way():
    key = torch.unique(ei[0] * N + ei[1], sorted=True)
    rows = torch.div(key, N, rounding_mode="floor")
    indices = (key - rows * N).to(torch.int32)
    ip = torch.zeros(N + 1, dtype=torch.int64, device=dev)
    torch.cumsum(torch.bincount(rows, minlength=N), 0, out=ip[1:])
    return ip, indices


res = {}
for mode, tag in ((0, "csr_from_edges_ms"), (1, "direct_scatter_ms"), (2, "partitioned_scatter_ms")):
    L.cdll.grapes_csr_set_direct_scatter(mode)
    res[tag] = round(timed(build), 3)
L.cdll.grapes_csr_set_direct_scatter(0)
build()
nnz = int(meta[0])
ip2, ix2 = torch_way()
assert torch.equal(indptr, ip2) and torch.equal(cap[:nnz], ix2)
ip3, ix3 = csr_from_edge_index(ei, N, dev)
assert torch.equal(ip3, ip2) and torch.equal(ix3, ix2)
ms_t = timed(torch_way, reps=3)
alg = 16 * E + 4 * E + 16 * E + 8 * E + 8 * E + 4 * E + 8 * E + 2 * 12 * N + 8 * nnz      # phases of csrc/csr_build.cu
print(json.dumps({"workload": name, "N": N, "E": E, "nnz": nnz, **res, "torch_sort_unique_ms": round(ms_t, 3),
                  "algorithmic_GB": round(alg / 1e9, 3), "GBps": round(alg / res["csr_from_edges_ms"] / 1e6, 1),
                  "timed": "the C-ABI call on preallocated buffers, median of 5"}))
