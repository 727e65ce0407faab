"""Graph statistics of ccsd/src/evaluation (test infrastructure, numpy / networkx only): the MMD estimator of
mmd.py:230-257 (compute_mmd: mean k(x, x') + mean k(y, y') - 2 mean k(x, y), histograms normalised to PMFs) with the
total-variation Gaussian kernel gaussian_tv (mmd.py:111-131; the default EMD kernel needs pyemd), degree histograms
(stats.py:40-57, sigma = 1) and clustering-coefficient histograms (stats.py:206-220, 100 bins, sigma = 0.1)."""
from __future__ import annotations

from typing import List, Sequence

import networkx as nx
import numpy as np


def graphs_from_adjs(adjs: np.ndarray) -> List[nx.Graph]:
    """Quantised adjacency matrices [B,N,N] -> graphs without isolated nodes / self loops (graph_utils.adjs_to_graphs)."""
    out = []
    for a in adjs:
        g = nx.from_numpy_array(np.asarray(a, dtype=np.int64))
        g.remove_edges_from(list(nx.selfloop_edges(g)))
        g.remove_nodes_from(list(nx.isolates(g)))
        if g.number_of_nodes():
            out.append(g)
    return out


def _pad(hs: Sequence[np.ndarray]) -> np.ndarray:
    L = max(len(h) for h in hs)
    m = np.zeros((len(hs), L))
    for i, h in enumerate(hs):
        s = float(np.sum(h))
        m[i, : len(h)] = np.asarray(h, dtype=np.float64) / s if s else h
    return m


def mmd_tv(h1: Sequence[np.ndarray], h2: Sequence[np.ndarray], sigma: float) -> float:
    a, b = _pad(h1), _pad(h2)
    L = max(a.shape[1], b.shape[1])
    a = np.pad(a, ((0, 0), (0, L - a.shape[1])))
    b = np.pad(b, ((0, 0), (0, L - b.shape[1])))

    def k(x, y):
        d = 0.5 * np.abs(x[:, None, :] - y[None, :, :]).sum(-1)
        return np.exp(-d * d / (2 * sigma * sigma)).mean()

    return float(k(a, a) + k(b, b) - 2 * k(a, b))


def degree_mmd(g1: List[nx.Graph], g2: List[nx.Graph]) -> float:
    return mmd_tv([np.array(nx.degree_histogram(g)) for g in g1], [np.array(nx.degree_histogram(g)) for g in g2], 1.0)


def clustering_mmd(g1: List[nx.Graph], g2: List[nx.Graph], bins: int = 100) -> float:
    def hist(g):
        return np.histogram(list(nx.clustering(g).values()), bins=bins, range=(0.0, 1.0), density=False)[0]
    return mmd_tv([hist(g) for g in g1], [hist(g) for g in g2], 0.1)


def edge_density(adjs: np.ndarray, flags: np.ndarray) -> float:
    """Edges / possible edges among the active nodes, averaged over the batch."""
    n = flags.sum(1)
    e = np.triu(adjs, 1).sum((1, 2))
    return float(np.mean(e / np.maximum(n * (n - 1) / 2, 1)))
