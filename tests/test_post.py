"""SURVEY 8f rank 2 -- the steps either side of the sampler, against golden vectors recorded from the unmodified
reference (tests/golden/make_golden_post.py): cc_from_incidence (cc_utils.py:156-265) as a batched device pass, and
init_flags (cc_utils.py:883-914) as a cached-table gather.  Bit exact."""
import json
from pathlib import Path

import numpy as np
import pytest
import torch

from ccsd_b200 import post

GOLD = json.loads((Path(__file__).resolve().parent / "golden" / "post_reference.json").read_text())
HAS_CUDA = torch.cuda.is_available()
DEV = "cuda" if HAS_CUDA else "cpu"


def _make_batch(seed, B, N, F, d_min, d_max):
    """Same seeded inputs as tests/golden/make_golden_post.py::make_batch."""
    g = torch.Generator().manual_seed(seed)
    E = N * (N - 1) // 2
    K = len(post.cell_table(N, d_min, d_max))
    n = torch.randint(max(2, N // 2), N + 1, (B,), generator=g)
    flags = (torch.arange(N)[None, :] < n[:, None]).float()
    x = (torch.rand(B, N, F, generator=g) > 0.4).float() * flags[:, :, None]
    a = torch.triu((torch.rand(B, N, N, generator=g) > 0.6).float(), 1)
    adj = (a + a.transpose(1, 2)) * flags[:, :, None] * flags[:, None, :]
    r2 = torch.randn(B, E, K, generator=g)
    r2 = torch.where(torch.rand(B, E, K, generator=g) > 0.93, r2, torch.zeros(()))
    r2[:, :, ::3] = 0
    r2[0, 1, 1] = -2.5
    r2[0, 2, 1] = 2.5
    return x, adj, r2


def _check_cc(case):
    x, adj, r2 = _make_batch(case["seed"], case["B"], case["N"], case["F"], case["d_min"], case["d_max"])
    got = post.ccs_from_incidence(x.to(DEV), adj.to(DEV), r2.to(DEV), case["d_min"], case["d_max"])
    assert len(got) == case["B"]
    for calls, ref in zip(got, case["calls"]):
        assert len(calls) == len(ref)
        for (cell, rank, attr), (rcell, rrank, rattr) in zip(calls, ref):
            assert list(cell) == rcell and rank == rrank
            assert set(attr) == set(rattr)
            for k in attr:
                assert np.float32(attr[k]) == np.float32(rattr[k]), (cell, k)
    # the device reduction itself against torch
    present, row, label = post.cc_cells(r2.to(DEV))
    assert torch.equal(present.cpu().bool(), (r2 != 0).any(1))
    assert torch.equal(row.cpu().long(), r2.abs().argmax(1))
    assert torch.equal(label.cpu(), torch.gather(r2, 1, r2.abs().argmax(1, keepdim=True)).squeeze(1))


def _check_flags():
    shape = GOLD["community_small_adjs_shape"]
    adjs = np.unpackbits(np.array(GOLD["community_small_adjs_packbits"], np.uint8))[: int(np.prod(shape))].reshape(shape)
    fs = post.FlagSampler(torch.from_numpy(adjs.astype(np.float32)), 20, device=DEV)
    for case in GOLD["init_flags"]:
        np.random.seed(case["seed"])
        fl = fs.sample(case["batch_size"] or 16)
        assert fl.dtype == torch.float32 and fl.device.type == DEV
        assert torch.equal(fl.cpu().long(), torch.tensor(case["flags"]))


@pytest.mark.skipif(HAS_CUDA, reason="host-emulation variant")
@pytest.mark.parametrize("i", range(len(GOLD["cc_from_incidence"])))
def test_cc_from_incidence_emulated(i):
    _check_cc(GOLD["cc_from_incidence"][i])


@pytest.mark.skipif(HAS_CUDA, reason="CPU variant")
def test_flag_sampler_cpu():
    _check_flags()


def test_graphs_to_tensor_matches_the_reference_tensor():
    """graphs_to_tensor on networkx graphs rebuilt from the shipped tensor gives the tensor back."""
    import networkx as nx
    shape = GOLD["community_small_adjs_shape"]
    adjs = np.unpackbits(np.array(GOLD["community_small_adjs_packbits"], np.uint8))[: int(np.prod(shape))].reshape(shape)
    graphs = []
    for a in adjs[:5]:
        n = int((a.sum(-1) > 0).sum())
        g = nx.from_numpy_array(a[:n, :n].astype(float))
        nx.set_node_attributes(g, 0, "feature")
        graphs.append(g)
    t = post.graphs_to_tensor(graphs, 20)
    assert torch.equal(t, torch.from_numpy(adjs[:5].astype(np.float32)))


@pytest.mark.gpu
@pytest.mark.parametrize("i", range(len(GOLD["cc_from_incidence"])))
def test_cc_from_incidence_gpu(i):
    _check_cc(GOLD["cc_from_incidence"][i])


@pytest.mark.gpu
def test_flag_sampler_gpu():
    _check_flags()


@pytest.mark.gpu
def test_cc_cells_full_size_properties():
    """community_small_CC batch (1024 x 190 x 1140): presence / first-argmax / label against torch on the device."""
    g = torch.Generator(device="cuda").manual_seed(0)
    r2 = torch.randn(1024, 190, 1140, device="cuda", generator=g)
    r2 = torch.where(torch.rand(r2.shape, device="cuda", generator=g) > 0.99, r2, torch.zeros((), device="cuda"))
    present, row, label = post.cc_cells(r2)
    am = r2.abs().argmax(1)
    assert torch.equal(present.bool(), (r2 != 0).any(1))
    assert torch.equal(label, torch.gather(r2, 1, am[:, None, :]).squeeze(1))
    assert torch.equal(row.long()[present.bool()], am[present.bool()])
