"""Times grapes_csr_from_edges (main.py:134-136) on a synthetic shape, CUDA events on the
launching stream, next to the torch sort/unique formulation it replaced.  Prints one JSON line."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from grapes_b200.graph import csr_from_edge_index          # noqa: E402
from grapes_b200.synth import SHAPES, synth_edge_index       # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "products"
cfg = SHAPES[name]
dev = torch.device("cuda:0")
N, E = cfg["N"], cfg["E_dir"]
ei = synth_edge_index(N, E, 0, device=dev)
torch.cuda.synchronize()


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def torch_way():
    key = torch.unique(ei[0] * N + ei[1], sorted=True)
    rows = torch.div(key, N, rounding_mode="floor")
    indices = (key - rows * N).to(torch.int32)
    indptr = torch.zeros(N + 1, dtype=torch.int64, device=dev)
    torch.cumsum(torch.bincount(rows, minlength=N), 0, out=indptr[1:])
    return indptr, indices


ms, (indptr, indices) = timed(lambda: csr_from_edge_index(ei, N, dev))
ms_t, (ip2, ix2) = timed(torch_way)
assert torch.equal(indptr, ip2) and torch.equal(indices, ix2)
nnz = int(indices.numel())
alg = 16 * E + 4 * E + 16 * E + 4 * E + 8 * E + 2 * 12 * N + 8 * nnz      # phases of csrc/csr_build.cu
print(json.dumps({"workload": name, "N": N, "E": E, "nnz": nnz, "csr_from_edges_ms": round(ms, 3),
                  "torch_sort_unique_ms": round(ms_t, 3), "algorithmic_GB": round(alg / 1e9, 3),
                  "GBps": round(alg / ms / 1e6, 1), "includes": "workspace allocation + nnz readback (one sync)"}))
