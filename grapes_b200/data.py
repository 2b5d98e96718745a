"""``get_data(root, name, seed, split_id) -> (data, num_features, num_classes)`` with the reference's
signature (/root/reference/modules/data.py:252-292).

The reference downloads Planetoid / OGB / Reddit2 / ... through PyG and ogb; neither the packages nor a
network exist here, so dataset NAMES map to synthetic graphs of the same SHAPE.
A real dataset exported as ``<root>/<name>.pt`` (a dict with x, y, edge_index, train/val/test masks) is
loaded instead when present.
"""
from __future__ import annotations

import os

import torch

from .synth import SHAPES, SynthData, make_synth

_ALIASES = {"cora": "cora", "ogbn-arxiv": "arxiv", "arxiv": "arxiv", "reddit": "reddit", "reddit2": "reddit",
            "products": "products", "ogbn-products": "products", "papers": "papers", "ogbn-papers100m": "papers",
            "tiny": "tiny", "small": "small"}


def get_data(root: str, name: str, seed: int = None, split_id: int = 0):
    path = os.path.join(root or ".", f"{name}.pt")
    if os.path.isfile(path):
        d = torch.load(path)
        x = d.get("x")
        data = SynthData(x=x, y=d["y"], edge_index=d["edge_index"], train_mask=d["train_mask"],
                         val_mask=d["val_mask"], test_mask=d["test_mask"], num_nodes=int(d["y"].shape[0]),
                         num_features=0 if x is None else int(x.shape[1]),
                         num_classes=int(d["y"].max()) + 1 if d["y"].dim() == 1 else int(d["y"].shape[1]))
        return data, data.num_features, data.num_classes
    key = name.lower()
    if key.startswith("synth:"):
        key = key.split(":", 1)[1]
    shape = _ALIASES.get(key)
    if shape is None or shape not in SHAPES:
        raise ValueError(f"Dataset {name} is not available offline (no {path}); synthetic shapes: {sorted(SHAPES)}")
    if not name.lower().startswith("synth:"):
        from .utils import get_logger
        get_logger().warning(f"dataset '{name}' is not available offline (no {path}): substituting a SYNTHETIC graph of the "
                             f"same shape ('{shape}') with random features and labels -- scores are not the dataset's")
    data = make_synth(shape, seed=0 if seed is None else seed)
    data.synthetic = True
    return data, data.num_features, data.num_classes
