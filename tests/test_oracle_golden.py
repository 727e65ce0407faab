"""Pin the oracle: (1) the reference's OWN known-answer tests (tests/models/*.py of the reference,
replayed into tests/golden/kat_reference_tests.npz by make_golden.py), (2) outputs of the unmodified
reference on seeded inputs with the shipped checkpoints (io_<cfg>.npz), (3) the exact tensors the
reference's utility tests expect (tests/utils/test_graph_utils.py, test_cc_utils.py).  CPU only."""
import json

import numpy as np
import pytest
import torch

from oracle import ccsd_oracle as O
from tests.helpers import GOLDEN, Config, check_compressed, rel_err


def _kat():
    z = np.load(GOLDEN / "kat_reference_tests.npz")
    return z, {e["test"]: e for e in json.loads(bytes(z["index"]).decode())}


def _sd(z, t):
    p = f"{t}/sd/"
    return {k[len(p):]: torch.from_numpy(z[k]) for k in z.files if k.startswith(p)}


def _assert_kat(z, t, outs, picks):
    """`picks[i]` extracts from the oracle outputs the slice the reference's i-th assertion looked at."""
    for i, pick in enumerate(picks):
        exp = torch.from_numpy(z[f"{t}/assert{i}/expected"])
        atol = float(z[f"{t}/assert{i}/atol"])
        got = pick(outs)
        assert got.shape == exp.shape
        assert torch.allclose(got, exp, atol=atol), (t, i, (got - exp).abs().max())
    for i, o in enumerate(outs):  # and the full reference output, tightly
        ref = torch.from_numpy(z[f"{t}/out{i}"])
        assert (o - ref).abs().max() <= 2e-6 * (1 + ref.abs().max()), (t, i)


def test_kat_score_network_a_cc():
    """reference tests/models/test_ScoreNetwork_A_CC.py:119-162 (5x5 expected score)."""
    z, idx = _kat()
    t = "test_ScoreNetworkA_CC"
    hp = idx[t]["hp"]
    x, adj, r2 = (torch.from_numpy(z[f"{t}/arg{i}"]) for i in range(3))
    out = O.score_network_a_cc(_sd(z, t), hp, x, adj, r2, None)
    _assert_kat(z, t, [out], [lambda o: o[0]])


def test_kat_score_network_f():
    """reference tests/models/test_ScoreNetwork_F.py:69-106."""
    z, idx = _kat()
    t = "test_ScoreNetworkF"
    hp = idx[t]["hp"]
    r2 = torch.from_numpy(z[f"{t}/arg2"])
    out = O.score_network_f(_sd(z, t), hp, r2, None)
    _assert_kat(z, t, [out], [lambda o: o[0][0, :, 0]])


def test_kat_dense_hcn_conv():
    """reference tests/models/test_hodge_layers.py:143-187."""
    z, _ = _kat()
    t = "test_DenseHCNConv"
    h, r2 = torch.from_numpy(z[f"{t}/arg0"]), torch.from_numpy(z[f"{t}/arg1"])
    out = O.dense_hcn(_sd(z, t), h, r2)
    _assert_kat(z, t, [out], [lambda o: o[0]])


def test_kat_hodge_network_layer():
    """reference tests/models/test_hodge_layers.py:190-241."""
    z, idx = _kat()
    t = "test_HodgeNetworkLayer"
    hp = idx[t]["hp"]
    r2c = torch.from_numpy(z[f"{t}/arg0"])
    out = O.hodge_network_layer(_sd(z, t), r2c, idx[t]["arg1"], hp["d_min"], hp["d_max"], None)
    _assert_kat(z, t, [out], [lambda o: o[0][0, 0, 0]])


def test_kat_hodge_attention():
    """reference tests/models/test_hodge_attention.py:96-140."""
    z, idx = _kat()
    t = "test_HodgeAttention"
    h, r2 = torch.from_numpy(z[f"{t}/arg0"]), torch.from_numpy(z[f"{t}/arg1"])
    v, a = O.hodge_attention(_sd(z, t), h, r2, idx[t]["hp"]["num_heads"])
    _assert_kat(z, t, [v, a], [lambda o: o[0][0, :, 0], lambda o: o[1][0, 0, :]])


def test_kat_hodge_adj_attention_layer():
    """reference tests/models/test_hodge_attention.py:143-209."""
    z, idx = _kat()
    t = "test_HodgeAdjAttentionLayer"
    hp = idx[t]["hp"]
    h, r2 = torch.from_numpy(z[f"{t}/arg0"]), torch.from_numpy(z[f"{t}/arg1"])
    ho, ro = O.hodge_adj_attention_layer(_sd(z, t), h, r2, None, hp["num_heads"], hp["N"], hp["d_min"], hp["d_max"])
    _assert_kat(z, t, [ho, ro], [lambda o: o[0][0, 0, 0], lambda o: o[1][0, 0]])


def test_kat_baseline_block():
    """reference tests/models/test_hodge_layers.py:244-306 (out_rank2[0, 0, :], out_hodge_adj[0, 0, :])."""
    z, _ = _kat()
    t = "test_BaselineBlock"
    h, r2 = torch.from_numpy(z[f"{t}/arg0"]), torch.from_numpy(z[f"{t}/arg1"])
    ro, ho = O.baseline_block(_sd(z, t), h, r2)
    _assert_kat(z, t, [ro, ho], [lambda o: o[0][0, 0, :], lambda o: o[1][0, 0, :]])


def test_kat_hodge_baseline_layer():
    """reference tests/models/test_hodge_layers.py:309-375 (out_hodge_adj[0, 0, 0], out_rank2[0, 0])."""
    z, idx = _kat()
    t = "test_HodgeBaselineLayer"
    hp = idx[t]["hp"]
    h, r2 = torch.from_numpy(z[f"{t}/arg0"]), torch.from_numpy(z[f"{t}/arg1"])
    ho, ro = O.hodge_baseline_layer(_sd(z, t), h, r2, None, hp["N"], hp["d_min"], hp["d_max"])
    _assert_kat(z, t, [ho, ro], [lambda o: o[0][0, 0, 0], lambda o: o[1][0, 0]])


def test_kat_score_network_a_base_cc():
    """reference tests/models/test_ScoreNetwork_A_Base_CC.py:115-158 (5x5 expected score)."""
    z, idx = _kat()
    t = "test_ScoreNetworkA_Base_CC"
    hp = idx[t]["hp"]
    x, adj, r2 = (torch.from_numpy(z[f"{t}/arg{i}"]) for i in range(3))
    out = O.score_network_a_base_cc(_sd(z, t), hp, x, adj, r2, None)
    _assert_kat(z, t, [out], [lambda o: o[0]])


# ---- utility known answers (values from the reference's tests/utils) -------------------------
def test_mask_x_and_adjs():
    """reference tests/utils/test_graph_utils.py:35-59."""
    x = torch.ones(2, 3, 2)
    flags = torch.tensor([[1.0, 1.0, 0.0], [1.0, 0.0, 0.0]])
    out = O.mask_x(x, flags)
    assert out[0, 2].abs().sum() == 0 and out[1, 1:].abs().sum() == 0 and out[0, :2].sum() == 4
    a = torch.ones(2, 3, 3)
    m = O.mask_adjs(a, flags)
    assert m[0].tolist() == [[1, 1, 0], [1, 1, 0], [0, 0, 0]] and m[1].sum() == 1
    m4 = O.mask_adjs(torch.ones(2, 4, 3, 3), flags)
    assert torch.equal(m4[:, 0], m)


def test_pow_tensor_and_quantize():
    """reference tests/utils/test_graph_utils.py:251-324."""
    a = torch.tensor([[[0.0, 1.0], [1.0, 0.0]]])
    p = O.pow_tensor(a, 3)
    assert p.shape == (1, 3, 2, 2)
    assert torch.equal(p[0, 1], torch.eye(2)) and torch.equal(p[0, 2], a[0])
    q = O.quantize(torch.tensor([0.2, 0.5, 0.7]))
    assert q.tolist() == [0.0, 1.0, 1.0]
    qm = O.quantize_mol(torch.tensor([0.4, 0.5, 1.49, 1.5, 2.49, 2.5, 9.0]))
    assert qm.tolist() == [0, 1, 1, 2, 2, 3, 3] and qm.dtype == torch.int64


def test_rank2_dim_and_flags():
    """reference tests/utils/test_cc_utils.py:355, 626-679 (get_rank2_dim, get_rank2_flags, mask_rank2)."""
    assert O.rank2_dim(5, 3, 4) == (10, 15)
    assert O.rank2_dim(9, 3, 9) == (36, 466)
    assert O.rank2_dim(20, 3, 3) == (190, 1140)
    flags = torch.tensor([[1.0, 1.0, 1.0, 1.0, 0.0]])
    fl, fr = O.edge_flags(flags), O.cell_flags(flags, 3, 4)
    # edges containing node 4: (0,4)=3, (1,4)=6, (2,4)=8, (3,4)=9
    assert fl[0].tolist() == [1, 1, 1, 0, 1, 1, 0, 1, 0, 0]
    # cells of size 3 over 5 nodes: those without node 4 are 012, 013, 023, 123 -> indices 0,1,3,6 ; size 4: 0123 -> 10
    assert [i for i, v in enumerate(fr[0].tolist()) if v == 1] == [0, 1, 3, 6, 10]
    r = O.mask_rank2(torch.ones(1, 10, 15), 5, 3, 4, flags)
    assert r.sum() == 6 * 5


def test_hodge_dual_round_trip():
    """reference tests/utils/test_cc_utils.py:1514-1548."""
    a = torch.tensor([[[0.0, 1.0, 2.0], [1.0, 0.0, 3.0], [2.0, 3.0, 0.0]]])
    h = O.adj_to_hodgedual(a)
    assert torch.equal(h[0], torch.diag(torch.tensor([1.0, 2.0, 3.0])))
    assert torch.equal(O.hodgedual_to_adj(h), a)
    # only the diagonal is read back
    h2 = h + (1 - torch.eye(3)) * 7.0
    assert torch.equal(O.hodgedual_to_adj(h2), a)


def test_pow_tensor_cc_hodge_mask():
    """reference tests/utils/test_cc_utils.py:956-1068."""
    f = torch.tensor([[[1.0, 0.0], [1.0, 1.0], [0.0, 2.0]]])
    H = f @ f.transpose(-1, -2)
    out = O.pow_tensor_cc(f, 2, True)
    Hm = H * (1 - torch.eye(3))
    assert torch.equal(out[:, 0], f) and torch.allclose(out[:, 1], Hm @ f)
    out2 = O.pow_tensor_cc(f, 2, False)
    assert torch.allclose(out2[:, 1], H @ f)


# ---- outputs of the unmodified reference on the shipped checkpoints --------------------------
CFGS = ["qm9", "community_small", "ego_small", "qm9_cc", "community_small_cc", "enzymes_small_cc", "ego_small_cc",
        "qm9_base_cc", "community_small_base_cc", "enzymes_small_base_cc", "ego_small_cc_v2", "zinc250k", "enzymes_small",
        "enzymes", "grid", "grid_small_cc"]


@pytest.mark.parametrize("name", CFGS)
def test_networks_match_reference_outputs(name):
    cfg = Config(name)
    io = cfg.io()
    flags, x, adj = (torch.from_numpy(io[k]) for k in ("flags", "x", "adj"))
    B = x.shape[0]
    args = [x, adj]
    if cfg.is_cc:
        g = torch.Generator().manual_seed(int(io["seed_inputs"]))
        # replay make_golden.make_config's generator stream
        n = torch.randint(max(2, cfg.N // 2), cfg.N + 1, (B,), generator=g)
        torch.randn(B, cfg.N, cfg.F, generator=g)
        torch.randn(B, cfg.N, cfg.N, generator=g)
        r2 = O.mask_rank2(torch.randn(B, cfg.E, cfg.K, generator=g) * 0.3, cfg.N, cfg.d_min, cfg.d_max, flags)
        args.append(r2)
    args.append(flags)
    for k, m in zip(cfg.keys, cfg.oracle_models):
        err = check_compressed(io, f"net_{k}", m(*args), 1e-5)
        assert err < 1e-5, (name, k, err)


@pytest.mark.parametrize("name", CFGS)
def test_samplers_match_reference_runs(name):
    """3 steps of the shipped sampler (and the PC/S4 alternative) on the real 1000-step schedule;
    the oracle replays torch's global CPU generator exactly as the reference consumed it."""
    cfg = Config(name)
    io = cfg.io()
    flags = torch.from_numpy(io["flags"])
    B = flags.shape[0]
    sh = cfg.shipped
    runs = sorted({k.split("/")[0] for k in io.files if k.startswith("run_")})
    assert runs
    for tag in runs:
        _, pred, corr = tag.split("_")
        steps = int(io[f"{tag}/steps"])
        torch.manual_seed(int(io[f"{tag}/seed"]))
        kw = dict(snr=sh["snr"], scale_eps=sh["scale_eps"], denoise=True, eps=1e-4, d_min=cfg.d_min, d_max=cfg.d_max,
                  noise=O.NoiseSource(seed=None), max_steps=steps)
        if pred == "S4":
            res, n = O.s4_solver(cfg.oracle_models, cfg.sdes(), cfg.shapes(B), flags, **kw)
            assert n == 0
        else:
            res, n = O.pc_sampler(cfg.oracle_models, cfg.sdes(), cfg.shapes(B), flags, predictor=pred, corrector=corr,
                                  n_steps=1, **kw)
            assert n == 2000
        for k, t in zip(cfg.keys, res):
            err = check_compressed(io, f"{tag}/{k}", t, 1e-5)
            assert err < 2e-5, (name, tag, k, err)
