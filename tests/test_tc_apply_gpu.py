"""tcgen05 H.F kernel (MN-major A operand, resident K-major H, double-buffered TMEM) against the fp32
FMA apply kernel: raw ScoreNetworkF output, and sampler steps with injected noise in both modes."""
import os

import pytest
import torch

from tests.helpers import Config, rel_err
from tests.parity_cases import make_engine, sampler_parity

pytestmark = pytest.mark.gpu


def _engine(cfg, B, no_tc, **kw):
    old = os.environ.get("CCSD_B200_NO_TC")
    os.environ["CCSD_B200_NO_TC"] = "1" if no_tc else "0"
    try:
        return make_engine(cfg, B, "cuda", **kw)
    finally:
        if old is None:
            del os.environ["CCSD_B200_NO_TC"]
        else:
            os.environ["CCSD_B200_NO_TC"] = old


@pytest.mark.parametrize("name,B", [("community_small_cc", 3), ("qm9_cc", 160), ("enzymes_small_cc", 9)])
def test_tc_apply_matches_fp32_kernel(name, B):
    cfg = Config(name)
    x, adj, r2, flags = cfg.random_state(B, seed=4, r2_scale=0.5)
    ref = cfg.oracle_models[2](x, adj, r2, flags)
    a = _engine(cfg, B, no_tc=True).score(2, x, adj, r2, flags).cpu()
    b = _engine(cfg, B, no_tc=False).score(2, x, adj, r2, flags).cpu()
    # bf16x3 keeps ~2^-16 relative error per product; with these deliberately large inputs (|H F| ~ 100x
    # the output) the deepest network (ENZYMES) reaches 5e-5 -- still inside the 1e-4 parity bar
    assert rel_err(a, ref) < 1e-5
    assert rel_err(b, ref) < 1e-4, rel_err(b, ref)
    x, adj, r2, flags = cfg.random_state(B, seed=5, r2_scale=0.1)   # realistic magnitudes
    ref = cfg.oracle_models[2](x, adj, r2, flags)
    b = _engine(cfg, B, no_tc=False).score(2, x, adj, r2, flags).cpu()
    assert rel_err(b, ref) < 1e-5, rel_err(b, ref)


@pytest.mark.parametrize("name,sampler,pred,corr", [("qm9_cc", "PC", "Reverse", "Langevin"),
                                                    ("enzymes_small_cc", "S4", "None", "None")])
def test_tc_sampler_steps(name, sampler, pred, corr):
    res = sampler_parity(name, sampler, pred, corr, 8, 3, "cuda")
    for k, (e_ret, e_state, agree) in res.items():
        assert e_ret < 1e-4 and e_state < 1e-4, (name, k, e_ret, e_state)
