"""tcgen05 Gram kernel (bf16x3 split, TMEM accumulation) against the fp32 FMA Gram kernel and an
fp64 reference on the same inputs, for aligned and unaligned K, one and two M tiles."""
import pytest
import torch

from tests.helpers import Config, rel_err
from tests.parity_cases import make_engine

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,B", [("community_small_cc", 5), ("qm9_cc", 37), ("enzymes_small_cc", 9)])
def test_tc_gram_matches_fp32_and_fp64(name, B):
    cfg = Config(name)
    _, _, r2, flags = cfg.random_state(B, seed=3, r2_scale=1.0)
    eng = make_engine(cfg, B, "cuda")
    eng.init(flags, seed=0)
    H0, P0 = eng.debug_gram(r2, use_tc=False)
    H1, P1 = eng.debug_gram(r2, use_tc=True)
    ref = (r2.double() @ r2.double().transpose(-1, -2)) * (1 - torch.eye(cfg.E, dtype=torch.float64))
    assert rel_err(H0, ref) < 1e-5
    assert rel_err(H1, ref) < 2e-5, rel_err(H1, ref)
    assert rel_err(H1, H0) < 2e-5
    assert rel_err(P1, P0) < 2e-5, rel_err(P1, P0)
    assert torch.equal(H1.diagonal(dim1=-2, dim2=-1), torch.zeros_like(H1.diagonal(dim1=-2, dim2=-1)))


def test_large_e_gram_and_apply_match_fp32_and_fp64():
    """grid_small_CC (E = 1176 > 192, K = 18424): the K-chunked tcgen05 Gram product (tc_r2big.cuh, upper tiles + mirror)
    against the fp32 FMA kernel and fp64; H exactly symmetric with a zero diagonal; ScoreNetworkF through the tcgen05
    H . F against the fp32 path and the oracle."""
    import os
    cfg = Config("grid_small_cc")
    B = 2
    x, adj, r2, flags = cfg.random_state(B, seed=3, r2_scale=0.3)
    eng = make_engine(cfg, B, "cuda")
    assert eng.lib.ccsd_plan_info(eng.handle, 18) == 1
    eng.init(flags, seed=0)
    H0, P0 = eng.debug_gram(r2, use_tc=False)
    H1, P1 = eng.debug_gram(r2, use_tc=True)
    ref = (r2.double() @ r2.double().transpose(-1, -2)) * (1 - torch.eye(cfg.E, dtype=torch.float64))
    assert rel_err(H0, ref) < 1e-5
    # bf16x3 keeps 2^-16 per PRODUCT: over an 18424-term contraction of zero-mean rows the entries are ~sqrt(K) smaller than
    # sum |terms|, so the bound is relative to |f_e| |f_c| (Cauchy-Schwarz), not to the entry
    scale = r2.double().pow(2).sum(-1).max().item()
    assert (H1.cpu().double() - ref).abs().max().item() < 3e-5 * scale
    assert rel_err(H1, ref) < 3e-4, rel_err(H1, ref)
    assert rel_err(P1, P0) < 1e-4, rel_err(P1, P0)
    assert torch.equal(H1, H1.transpose(-1, -2))
    assert torch.equal(H1.diagonal(dim1=-2, dim2=-1), torch.zeros_like(H1.diagonal(dim1=-2, dim2=-1)))
    out_tc = eng.score(2, x, adj, r2, flags).cpu()
    os.environ["CCSD_B200_NO_TC_BIG"] = "1"
    try:
        out_fp = make_engine(cfg, B, "cuda").score(2, x, adj, r2, flags).cpu()
    finally:
        del os.environ["CCSD_B200_NO_TC_BIG"]
    ref_s = cfg.oracle_models[2](x, adj, r2, flags)
    assert rel_err(out_fp, ref_s) < 1e-5
    assert rel_err(out_tc, ref_s) < 1e-4, rel_err(out_tc, ref_s)
