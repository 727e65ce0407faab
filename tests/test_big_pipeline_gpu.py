"""Large-graph pipeline (big_pipe.cuh, N > 64) on the GPU: tensor-core kernels (tc_agg, tc_afinal on full planes) against its
own fp32 FMA kernels (big_agg_kernel, big_final_kernel) and against the oracle, and the per-pair edge kernel against the
row-tile one."""
import os

import pytest
import torch

from tests.helpers import Config, rel_err
from tests.parity_cases import make_engine

pytestmark = pytest.mark.gpu


def _engine(cfg, B, no_tc):
    old = os.environ.get("CCSD_B200_NO_TC")
    os.environ["CCSD_B200_NO_TC"] = "1" if no_tc else "0"
    try:
        return make_engine(cfg, B, "cuda")
    finally:
        if old is None:
            del os.environ["CCSD_B200_NO_TC"]
        else:
            os.environ["CCSD_B200_NO_TC"] = old


@pytest.mark.parametrize("name,B", [("enzymes", 3), ("grid", 1)])
def test_tensor_core_kernels_match_fp32_kernels(name, B):
    cfg = Config(name)
    x, adj, r2, flags = cfg.random_state(B, seed=7)
    ref = cfg.oracle_models[1](x, adj, flags)
    a = _engine(cfg, B, no_tc=True).score(1, x, adj, r2, flags).cpu()
    b = _engine(cfg, B, no_tc=False).score(1, x, adj, r2, flags).cpu()
    assert rel_err(a, ref) < 2e-5, rel_err(a, ref)      # fp32 FMA pipeline
    assert rel_err(b, ref) < 1e-4, rel_err(b, ref)      # bf16x3 aggregation + final MLP
    assert rel_err(b, a) < 1e-4
    # symmetric, zero diagonal, masked rows: properties of the adjacency score at any size
    assert torch.equal(b, b.transpose(-1, -2))
    assert float(b.diagonal(dim1=-2, dim2=-1).abs().max()) == 0.0
    dead = flags == 0
    if dead.any():
        assert float(b[dead].abs().max()) == 0.0


def test_full_size_grid_batch_properties():
    """BASELINE configs[4] at its bench size (N = 361, B = 64): 3 sampler steps with Philox noise; size-independent properties."""
    cfg = Config("grid")
    B = 64
    _, _, _, flags = cfg.random_state(B, seed=3)
    eng = make_engine(cfg, B, "cuda", "PC", "Reverse", "Langevin")
    eng.init(flags.cuda(), seed=5)
    eng.run(0, 3)
    x, adj = [t.cpu() for t in eng.read(True)]
    assert torch.isfinite(x).all() and torch.isfinite(adj).all()
    assert torch.equal(adj, adj.transpose(-1, -2))
    assert float(adj.diagonal(dim1=-2, dim2=-1).abs().max()) == 0.0
    dead = flags == 0
    assert float(x[dead].abs().max()) == 0.0 and float(adj[dead].abs().max()) == 0.0
    # determinism per seed and shard invariance of the Philox stream (sample 5 alone = sample 5 of the batch)
    eng.init(flags.cuda(), seed=5)
    eng.run(0, 3)
    x2, adj2 = [t.cpu() for t in eng.read(True)]
    assert torch.equal(x, x2) and torch.equal(adj, adj2)


@pytest.mark.parametrize("N", [65, 127, 128, 129, 192, 200, 256, 257])
def test_tile_boundary_sizes(N):
    """The ENZYMES checkpoint at graph sizes around the tile sizes of the large-graph kernels (128-row tiles and 64-node
    contraction chunks of tc_agg, 128-column segments of tc_afinal, 32 x 32 mirror tiles): the score networks have no
    N-dependent parameters."""
    cfg = Config("enzymes")
    cfg.N = N
    B = 3
    x, adj, r2, flags = cfg.random_state(B, seed=N)
    eng = make_engine(cfg, B, "cuda")
    for w in (0, 1):
        ref = cfg.oracle_models[w](x, adj, flags)
        out = eng.score(w, x, adj, r2, flags).cpu()
        assert torch.isfinite(out).all()
        assert rel_err(out, ref) < 1e-4, (N, w, rel_err(out, ref))
