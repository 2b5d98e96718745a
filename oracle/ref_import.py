"""TEST INFRASTRUCTURE ONLY.  Imports the reference's own ``modules/utils.py`` from
/root/reference (present in the build container, absent on the GPU box) so the
restatement in oracle/reference_port.py can be validated against it and golden vectors
can be generated from it.  Nothing under tests -m gpu, smoke() or bench.py may use this.

The reference pins scipy 1.13.1 (environment.yml:95); on scipy 1.18 indexing a CSR matrix
with a ``torch.Tensor`` raises, so :class:`TensorIndexCSR` converts tensor indices to numpy
first -- an index-type shim, not a semantic change.
"""
from __future__ import annotations

import importlib.util
import os
import sys

import numpy as np
import scipy.sparse as sp
import torch

REFERENCE_ROOT = os.environ.get("GRAPES_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "modules", "utils.py"))


def load_reference_utils():
    """Import /root/reference/modules/utils.py as a stand-alone module (it only needs
    logging/os/numpy/psutil/scipy/torch, utils.py:1-10)."""
    path = os.path.join(REFERENCE_ROOT, "modules", "utils.py")
    spec = importlib.util.spec_from_file_location("grapes_reference_utils", path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules["grapes_reference_utils"] = mod
    spec.loader.exec_module(mod)
    return mod


def _conv(i):
    if isinstance(i, torch.Tensor):
        return i.cpu().numpy()
    if isinstance(i, tuple):
        return tuple(_conv(j) for j in i)
    return i


class TensorIndexCSR:
    """Wraps a scipy CSR so ``adj[tensor]`` / ``adj[:, tensor]`` work like they did on scipy 1.13."""

    def __init__(self, m: sp.csr_matrix):
        self.m = m

    def __getitem__(self, idx):
        return TensorIndexCSR(self.m[_conv(idx)])

    def tocoo(self):
        return self.m.tocoo()

    @property
    def shape(self):
        return self.m.shape


def reference_adjacency(edge_index: torch.Tensor, num_nodes: int) -> TensorIndexCSR:
    """main.py:134-136 verbatim call (bool ones, edge_index as (row, col))."""
    ei = edge_index.numpy()
    adj = sp.csr_matrix((np.ones(ei.shape[1], dtype=bool), ei), shape=(num_nodes, num_nodes))
    return TensorIndexCSR(adj)


def load_reference_eval():
    """Import /root/reference/eval.py (``evaluate``, eval.py:11-165) as a stand-alone module.  It imports
    ``torch_geometric`` for ONE type annotation (eval.py:4,14: absent here) and ``modules.utils`` (the reference's own,
    loaded by :func:`load_reference_utils`); both names are bound for the duration of the import only.  Everything the
    function then executes is the reference's own code."""
    import types
    ru = load_reference_utils()
    path = os.path.join(REFERENCE_ROOT, "eval.py")
    tg = types.ModuleType("torch_geometric")
    tg.data = types.ModuleType("torch_geometric.data")
    tg.data.Data = object
    pkg = types.ModuleType("modules")
    pkg.utils = ru
    names = {"torch_geometric": tg, "torch_geometric.data": tg.data, "modules": pkg, "modules.utils": ru}
    saved = {k: sys.modules.get(k) for k in names}
    sys.modules.update(names)
    try:
        spec = importlib.util.spec_from_file_location("grapes_reference_eval", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod, ru


class RecordingModule(torch.nn.Module):
    """Wraps a model handed to the reference's ``evaluate`` and keeps what it was called with and what it returned (the
    reference returns only (accuracy, f1); the per-batch blocks and logits are read off its ``gcn_c`` calls)."""

    def __init__(self, inner):
        super().__init__()
        self.inner = inner
        self.calls = []

    def forward(self, x, edge_index):
        out = self.inner(x, edge_index)
        ei = [e.clone() for e in edge_index] if isinstance(edge_index, list) else edge_index.clone()
        self.calls.append({"x": x.clone(), "edge_index": ei, "logits": out[0].clone()})
        return out
