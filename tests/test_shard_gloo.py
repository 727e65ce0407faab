"""Multi-rank path on CPU: world_size-2 gloo processes, each sampling its shard of the batch through the
host-emulation build, one all-gather of the results (ccsd_b200/shard.py).  With no corrector there is no
batch coupling, and Philox is keyed by the GLOBAL sample index, so the gathered result must equal the
single-process run bit for bit; with the Langevin corrector every rank uses its shard-local batch mean
(the reference's divide_batch semantics, ccsd/src/sampler.py:224-232), which is checked against a
single-process run of the same shard."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ccsd_b200.shard import gather_rows, shard_bounds, sharded_sample
from ccsd_b200.solver import get_pc_sampler
from tests.helpers import Config

pytestmark = pytest.mark.skipif(torch.cuda.is_available(), reason="CPU (gloo + host emulation) test")
B, STEPS = 5, 3   # odd batch: ragged shards (3 + 2)


def _flags(cfg):
    g = torch.Generator().manual_seed(3)
    n = torch.randint(max(2, cfg.N // 2), cfg.N + 1, (B,), generator=g)
    return (torch.arange(cfg.N)[None, :] < n[:, None]).to(torch.float32)


def _maker(cfg, corrector):
    sd = cfg.sdes()
    sh = cfg.shipped

    def make(b):
        return get_pc_sampler(sd[0], sd[1], cfg.shapes(b)[0], cfg.shapes(b)[1], predictor="Reverse", corrector=corrector,
                              snr=sh["snr"], scale_eps=sh["scale_eps"], continuous=True, denoise=True, eps=1e-4, device="cpu",
                              is_cc=True, sde_rank2=sd[2], shape_rank2=cfg.shapes(b)[2], d_min=cfg.d_min, d_max=cfg.d_max)
    return make


def _worker_tiny(rank, world, port, out):
    """More ranks than samples: rank 1 gets an empty shard and must still take part in the gather."""
    from ccsd_b200 import _native as nat

    nat.enable_test_emulation(os.environ["CCSD_B200_TEST_EMU"])
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = Config("qm9_cc")
    res = sharded_sample(_maker(cfg, "None"), cfg.holders, _flags(cfg)[:1], max_steps=2, record_traj=False)   # seed broadcast
    torch.save([t.clone() for t in res], out + f".{rank}")
    dist.destroy_process_group()


def _worker(rank, world, port, corrector, out):
    from ccsd_b200 import _native as nat

    nat.enable_test_emulation(os.environ["CCSD_B200_TEST_EMU"])   # spawned process: conftest did not run here
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = Config("qm9_cc")
    res = sharded_sample(_maker(cfg, corrector), cfg.holders, _flags(cfg), seed=17, max_steps=STEPS, record_traj=False)
    if rank == 0:
        torch.save([t.clone() for t in res], out)
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_bounds_and_ragged_gather_single_process():
    assert [shard_bounds(5, 2, r) for r in range(2)] == [(0, 3), (3, 5)]
    assert [shard_bounds(8, 3, r) for r in range(3)] == [(0, 3), (3, 6), (6, 8)]
    t = torch.arange(6.).reshape(3, 2)
    assert torch.equal(gather_rows(t, [3]), t)


@pytest.mark.parametrize("corrector", ["None", "Langevin"])
def test_two_rank_gloo_matches_single_process(tmp_path, corrector):
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(2, _free_port(), corrector, out), nprocs=2, join=True)
    got = torch.load(out)
    cfg = Config("qm9_cc")
    flags = _flags(cfg)
    make = _maker(cfg, corrector)
    if corrector == "None":
        ref = make(B)(*cfg.holders, flags, seed=17, sample_offset=0, max_steps=STEPS, record_traj=False)[:3]
    else:   # shard-local batch means: the reference result is the concatenation of the two shard runs
        parts = []
        for r in range(2):
            lo, hi = shard_bounds(B, 2, r)
            parts.append(make(hi - lo)(*cfg.holders, flags[lo:hi], seed=17, sample_offset=lo, max_steps=STEPS, record_traj=False)[:3])
        ref = [torch.cat([p[k] for p in parts]) for k in range(3)]
    for g, r in zip(got, ref):
        assert g.shape == r.shape and torch.equal(g, r)


def test_more_ranks_than_samples(tmp_path):
    """One sample, two ranks: the rank with the empty shard joins the gather with zero rows (no hang), the seed is
    broadcast from rank 0, and both ranks end up with the same full result."""
    out = str(tmp_path / "tiny.pt")
    mp.spawn(_worker_tiny, args=(2, _free_port(), out), nprocs=2, join=True)
    a, b = torch.load(out + ".0"), torch.load(out + ".1")
    cfg = Config("qm9_cc")
    assert [tuple(t.shape) for t in a] == [tuple(s) for s in cfg.shapes(1)]
    for p, q in zip(a, b):
        assert torch.equal(p, q)
