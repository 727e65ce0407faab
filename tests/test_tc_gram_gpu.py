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
