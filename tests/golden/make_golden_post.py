"""Golden vectors for the post-processing row from the UNMODIFIED reference in /root/reference:
``cc_from_incidence`` (ccsd/src/utils/cc_utils.py:156-265) and ``init_flags`` (:883-914) on seeded inputs.

    python tests/golden/make_golden_post.py        (build container only)  ->  tests/golden/post_reference.json

toponetx is not installed here; its CombinatorialComplex is replaced by a class that RECORDS the add_cell calls the
reference makes (cell, rank, attributes, in order) -- which is exactly what the function computes."""
from __future__ import annotations

import json
import pickle
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import refstubs  # noqa: E402

refstubs.install()


class RecordingCC:
    def __init__(self, *a, **k):
        self.calls = []

    def add_cell(self, cell, rank=None, **attr):
        self.calls.append([list(cell), int(rank), {k: float(v) for k, v in attr.items()}])


sys.modules["toponetx.classes.combinatorial_complex"].CombinatorialComplex = RecordingCC
from ccsd.src.utils import cc_utils as rcc  # noqa: E402
from ccsd.src.utils import graph_utils as rgu  # noqa: E402
from easydict import EasyDict  # noqa: E402

rcc.CombinatorialComplex = RecordingCC


def make_batch(seed, B, N, F, d_min, d_max):
    """Quantised sampler-like outputs: masked node features, symmetric 0/1 adjacency, sparse signed rank-2 incidence."""
    g = torch.Generator().manual_seed(seed)
    E, K = N * (N - 1) // 2, rcc.get_rank2_dim(N, d_min, d_max)[1]
    n = torch.randint(max(2, N // 2), N + 1, (B,), generator=g)
    flags = (torch.arange(N)[None, :] < n[:, None]).float()
    x = (torch.rand(B, N, F, generator=g) > 0.4).float() * flags[:, :, None]
    a = torch.triu((torch.rand(B, N, N, generator=g) > 0.6).float(), 1)
    adj = (a + a.transpose(1, 2)) * flags[:, :, None] * flags[:, None, :]
    r2 = torch.randn(B, E, K, generator=g)
    r2 = torch.where(torch.rand(B, E, K, generator=g) > 0.93, r2, torch.zeros(()))
    r2[:, :, ::3] = 0           # whole candidate cells absent
    r2[0, 1, 1] = -2.5          # a negative entry with the largest magnitude of its column
    r2[0, 2, 1] = 2.5           # tie in |.|: the first maximum wins
    return x, adj, r2, E, K


def main():
    out = {"cc_from_incidence": [], "init_flags": []}
    for seed, B, N, F, d_min, d_max in [(0, 3, 6, 2, 3, 3), (1, 2, 7, 3, 3, 4), (2, 2, 5, 1, 2, 4)]:
        x, adj, r2, E, K = make_batch(seed, B, N, F, d_min, d_max)
        calls = []
        for b in range(B):
            cc = rcc.cc_from_incidence([x[b], adj[b], r2[b]], d_min, d_max, is_molecule=False)
            calls.append(cc.calls)
        out["cc_from_incidence"].append({"seed": seed, "B": B, "N": N, "F": F, "d_min": d_min, "d_max": d_max, "calls": calls})
    # init_flags on the shipped community_small graphs (plain networkx pickles)
    with open("/root/reference/data/community_small.pkl", "rb") as f:
        graphs = pickle.load(f)
    cfg = EasyDict({"data": {"batch_size": 16, "max_node_num": 20}})
    for seed, bs in [(0, None), (7, 33)]:
        np.random.seed(seed)
        fl = rcc.init_flags(graphs, cfg, bs)
        out["init_flags"].append({"seed": seed, "batch_size": bs, "flags": fl.numpy().astype(int).tolist()})
    # the graph set itself cannot travel: ship its padded adjacency tensor (what graphs_to_tensor returns), bit-packed
    adjs = rgu.graphs_to_tensor(graphs, 20).numpy().astype(np.uint8)
    out["community_small_adjs_packbits"] = np.packbits(adjs).tolist()
    out["community_small_adjs_shape"] = list(adjs.shape)
    (Path(__file__).resolve().parent / "post_reference.json").write_text(json.dumps(out))
    print("written", sum(len(c["calls"]) for c in out["cc_from_incidence"]), "complexes")


if __name__ == "__main__":
    main()
