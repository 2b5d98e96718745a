"""Synthetic graphs with the shapes BASELINE.json names.

The reference has no generator (its ``get_synth`` only loads files,
/root/reference/modules/data.py:88-111); the real datasets need network access.
SURVEY.md section 8(d) defines the generator used here: ``E_dir/2`` undirected
pairs with i.i.d. uniform endpoints, symmetrised, then fed to the same CSR
constructor main.py uses (/root/reference/main.py:134-136: duplicates collapse,
self-loops are kept).  Features ~ N(0,1) fp32, labels ~ U{0..C-1}, the train
split is the first ``n_train`` ids of a seeded permutation kept in ascending id
order (the reference's DataLoader does not shuffle, main.py:126).
"""
from __future__ import annotations

import dataclasses
from typing import Optional

import torch

#: name -> (N, E_dir, F, C, n_train, batch, k, hops)   
SHAPES = {
    "cora":     dict(N=2_708,       E_dir=10_556,        F=1_433, C=7,   n_train=1_208,   batch_size=512,  num_samples=16,  sampling_hops=2),
    "arxiv":    dict(N=169_343,     E_dir=2 * 1_166_243, F=128,   C=40,  n_train=90_941,  batch_size=256,  num_samples=256, sampling_hops=2),
    "reddit":   dict(N=232_965,     E_dir=114_615_892,   F=602,   C=41,  n_train=153_431, batch_size=256,  num_samples=256, sampling_hops=2),
    "products": dict(N=2_449_029,   E_dir=2 * 61_859_140, F=100,  C=47,  n_train=196_615, batch_size=1024, num_samples=256, sampling_hops=3),
    "papers":   dict(N=111_059_956, E_dir=2 * 1_615_685_872, F=128, C=172, n_train=1_207_179, batch_size=1024, num_samples=256, sampling_hops=2),
    # small shapes used by the tests
    "tiny":     dict(N=300,         E_dir=1_800,         F=12,    C=5,   n_train=120,     batch_size=32,   num_samples=8,   sampling_hops=2),
    "small":    dict(N=5_000,       E_dir=60_000,        F=36,    C=6,   n_train=2_000,   batch_size=128,  num_samples=32,  sampling_hops=3),
}


@dataclasses.dataclass
class SynthData:
    """Duck-types the fields of ``torch_geometric.data.Data`` the hot path reads
    (/root/reference/main.py:65-68,125-136): x, y, edge_index, masks."""
    x: Optional[torch.Tensor]
    y: torch.Tensor
    edge_index: torch.Tensor
    train_mask: torch.Tensor
    val_mask: torch.Tensor
    test_mask: torch.Tensor
    num_nodes: int
    num_features: int
    num_classes: int

    @property
    def num_edges(self) -> int:
        return int(self.edge_index.shape[1])


def synth_edge_index(N: int, E_dir: int, seed: int, device="cpu", power_law: float = 0.0) -> torch.Tensor:
    """``E_dir/2`` random undirected pairs, both directions emitted (int64 [2, E_dir]).

    ``power_law > 0`` draws endpoints from a Zipf-like law (hub rows), the stress
    variant of SURVEY.md section 8(d)."""
    g = torch.Generator(device=device).manual_seed(seed)
    half = E_dir // 2
    if power_law > 0.0:
        u = torch.rand(2, half, generator=g, device=device, dtype=torch.float64)
        pairs = (N * u.pow(1.0 + power_law)).long().clamp_(max=N - 1)
    else:
        pairs = torch.randint(0, N, (2, half), generator=g, device=device, dtype=torch.int64)
    return torch.cat([pairs, pairs.flip(0)], dim=1)


def make_synth(name: str = "cora", seed: int = 0, device="cpu", multilabel: bool = False,
               power_law: float = 0.0, features: bool = True, **overrides) -> SynthData:
    cfg = dict(SHAPES[name])
    cfg.update(overrides)
    N, E_dir, F, C, n_train = cfg["N"], cfg["E_dir"], cfg["F"], cfg["C"], cfg["n_train"]
    g = torch.Generator(device=device).manual_seed(seed + 1)
    edge_index = synth_edge_index(N, E_dir, seed, device=device, power_law=power_law)
    x = torch.randn(N, F, generator=g, device=device, dtype=torch.float32) if features else None
    if multilabel:
        y = (torch.rand(N, C, generator=g, device=device) < 0.2).float()
    else:
        y = torch.randint(0, C, (N,), generator=g, device=device, dtype=torch.int64)
    perm = torch.randperm(N, generator=g, device=device)
    n_val = min((N - n_train) // 2, n_train)
    train_mask = torch.zeros(N, dtype=torch.bool, device=device)
    val_mask = torch.zeros(N, dtype=torch.bool, device=device)
    test_mask = torch.zeros(N, dtype=torch.bool, device=device)
    train_mask[perm[:n_train]] = True
    val_mask[perm[n_train:n_train + n_val]] = True
    test_mask[perm[n_train + n_val:]] = True
    return SynthData(x=x, y=y, edge_index=edge_index, train_mask=train_mask, val_mask=val_mask,
                     test_mask=test_mask, num_nodes=N, num_features=F, num_classes=C)
