"""The steps either side of the sampler in ``Sampler_*.sample()``, batched:

* ``FlagSampler`` -- ``init_flags`` (ccsd/src/utils/cc_utils.py:883-914).  The reference rebuilds the padded
  adjacency tensor of the WHOLE training set on every call (``graphs_to_tensor`` / ``ccs_to_tensors``) to draw
  ``batch_size`` rows of it; here the per-object node flags are computed once, kept on the device, and a call is
  one ``np.random.randint`` (the reference's own draw, so the same numpy seed gives the same flags) + one gather.
* ``cc_cells`` / ``ccs_from_incidence`` -- ``cc_from_incidence`` (cc_utils.py:156-265): the reference walks the N
  nodes, N(N-1)/2 pairs and K candidate cells of EVERY sample in Python with ``.item()`` synchronisations; here the
  presence / label reductions of the whole batch run on the device (``ccsd_cc_cells`` for the rank-2 columns) and
  the host only touches the cells that exist.
* ``quantize`` / ``mol_onehot`` live in ``ccsd_b200.solver``.
"""
from __future__ import annotations

import itertools
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _native as nat


def node_flags(adj: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """graph_utils.py:62-77."""
    flags = torch.abs(adj).sum(-1).gt(eps).to(dtype=torch.float32)
    if flags.dim() == 3:
        flags = flags[:, 0, :]
    return flags


def graphs_to_tensor(graph_list: Sequence[Any], max_node_num: int) -> torch.Tensor:
    """graph_utils.py:324-356 (node order = ``g.nodes`` order, zero padding to max_node_num)."""
    import networkx as nx

    out = np.zeros((len(graph_list), max_node_num, max_node_num), np.float32)
    for n, g in enumerate(graph_list):
        nodes = [v for v, _ in g.nodes.data("feature")]
        a = nx.to_numpy_array(g, nodelist=nodes)
        if a.shape[-1] > max_node_num:
            raise ValueError(f"Original number of nodes {a.shape[-1]} is greater (>) that the desired number of nodes "
                             f"after padding {max_node_num}")
        out[n, : a.shape[0], : a.shape[1]] = a
    return torch.from_numpy(out)


class FlagSampler:
    """``init_flags`` with the dataset-side work done once.  ``objs``: a list of networkx graphs, or the padded
    adjacency tensor [n_objects, N, N] the caller already has (for combinatorial complexes: ``ccs_to_tensors(...)[0]``)."""

    def __init__(self, objs, max_node_num: int, device="cuda") -> None:
        adjs = objs if isinstance(objs, torch.Tensor) else graphs_to_tensor(objs, max_node_num)
        if adjs.dim() != 3 or adjs.shape[1] != max_node_num or adjs.shape[2] != max_node_num:
            raise ValueError("FlagSampler: expected adjacency tensors [n_objects, max_node_num, max_node_num]")
        self.n = adjs.shape[0]
        self.table = node_flags(adjs.to(device))          # [n_objects, N] on the sampler's device

    def sample(self, batch_size: int) -> torch.Tensor:
        idx = np.random.randint(0, self.n, batch_size)    # the reference's draw (cc_utils.py:904, 911)
        return self.table.index_select(0, torch.from_numpy(idx).to(self.table.device))


def cc_cells(rank2: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """(present uint8 [B,K], row int32 [B,K], label float32 [B,K]) of ``rank2`` [B,E,K] -- the per-column part of
    cc_from_incidence (cc_utils.py:243-247) for the whole batch, one device pass."""
    lib = nat.load()
    r2 = rank2.contiguous().to(torch.float32)
    if r2.dim() != 3:
        raise ValueError("cc_cells: rank2 must be [B,E,K]")
    if r2.device.type != "cuda" and not nat.is_emulation():
        raise RuntimeError("ccsd_b200 needs a CUDA device; there is no CPU fallback")
    B, E, K = r2.shape
    present = torch.empty(B, K, dtype=torch.uint8, device=r2.device)
    row = torch.empty(B, K, dtype=torch.int32, device=r2.device)
    label = torch.empty(B, K, dtype=torch.float32, device=r2.device)
    if B * K:
        stream = torch.cuda.current_stream(r2.device).cuda_stream if r2.device.type == "cuda" else None
        with (torch.cuda.device(r2.device) if r2.device.type == "cuda" else _null()):
            nat.check(lib.ccsd_cc_cells(r2.data_ptr(), present.data_ptr(), row.data_ptr(), label.data_ptr(), B, E, K, stream))
    return present, row, label


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def cell_table(N: int, d_min: int, d_max: int) -> List[Tuple[int, ...]]:
    """Candidate rank-2 cells in the reference's order (cc_utils.py:72-76)."""
    out: List[Tuple[int, ...]] = []
    for d in range(d_min, d_max + 1):
        out.extend(itertools.combinations(range(N), d))
    return out


def ccs_from_incidence(x: torch.Tensor, adj: torch.Tensor, rank2: Optional[torch.Tensor], d_min: int, d_max: int,
                       cc_factory=None) -> List[Any]:
    """Batched ``cc_from_incidence`` for non-molecule complexes with scalar edge / cell labels (what the CC samplers
    produce: sampler.py:531-560).  Returns one object per sample: ``cc_factory()`` receiving the reference's
    ``add_cell(cell, rank=..., **attr)`` calls in the reference's order (pass ``toponetx``'s CombinatorialComplex), or,
    when ``cc_factory`` is None, a plain list of ``(cell, rank, attr)`` tuples."""
    B, N = x.shape[0], x.shape[1]
    node_on = (x != 0).any(-1)                                   # :199
    iu = torch.triu_indices(N, N, 1, device=adj.device)
    edge_val = adj[:, iu[0], iu[1]]                              # :218 pairs i < j in row-major order
    node_on_h, x_h = node_on.cpu().numpy(), x.detach().cpu().numpy()
    edge_h = edge_val.detach().cpu().numpy()
    pairs = list(zip(iu[0].tolist(), iu[1].tolist()))
    cells = None
    if rank2 is not None:
        present, _, label = cc_cells(rank2)
        pres_h, lab_h = present.cpu().numpy(), label.cpu().numpy()
        cells = cell_table(N, d_min, d_max)
    out = []
    for b in range(B):
        rec = _Recorder() if cc_factory is None else cc_factory()
        for i in np.nonzero(node_on_h[b])[0]:
            rec.add_cell((int(i),), rank=0, **{f"label_{j}": x_h[b, i, j].item() for j in range(x_h.shape[2])})
        for e in np.nonzero(edge_h[b])[0]:
            rec.add_cell(pairs[e], rank=1, label=edge_h[b, e].item())
        if cells is not None:
            for k in np.nonzero(pres_h[b])[0]:
                rec.add_cell(cells[k], 2, label=lab_h[b, k].item())
        out.append(rec.calls if cc_factory is None else rec)
    return out


class _Recorder:
    def __init__(self) -> None:
        self.calls: List[Tuple[Tuple[int, ...], int, Dict[str, float]]] = []

    def add_cell(self, cell, rank, **attr):
        self.calls.append((tuple(cell), int(rank), dict(attr)))
