"""Parity tests proper (B200): the CUDA path through the C ABI versus the oracle on the same seeded
inputs and injected noise, against the committed reference outputs, and -- at full BASELINE sizes
with Philox noise -- through size-independent properties of the domain."""
import numpy as np
import pytest
import torch

from ccsd_b200 import _native as nat
from ccsd_b200.solver import get_pc_sampler, S4_solver, quantize
from oracle import ccsd_oracle as O
from tests.helpers import Config, check_compressed, rel_err
from tests.parity_cases import SCORE_TOL, make_engine, sampler_parity, score_parity

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_product_library_is_loaded():
    assert not nat.is_emulation()
    assert str(nat.lib_path()).endswith("ccsd_b200/_lib/libccsd_b200.so")


@pytest.mark.parametrize("name,B", [("qm9", 64), ("community_small", 16), ("qm9_cc", 16), ("enzymes_small_cc", 8),
                                    ("community_small_cc", 4), ("ego_small", 8), ("ego_small_cc", 2),
                                    ("qm9_base_cc", 16), ("community_small_base_cc", 4), ("enzymes_small_base_cc", 8),
                                    ("ego_small_cc_v2", 2), ("zinc250k", 8), ("enzymes_small", 16),
                                    ("enzymes", 4), ("grid", 2), ("grid_small_cc", 1)])
def test_score_parity(name, B):
    """per-step score outputs within 1e-4 relative of the fp32 reference path (north_star)."""
    errs = score_parity(name, B, DEV)
    for k, e in errs.items():
        assert e < SCORE_TOL, (name, k, e)


@pytest.mark.parametrize("name", ["qm9", "community_small", "ego_small", "qm9_cc", "community_small_cc", "enzymes_small_cc",
                                  "ego_small_cc", "qm9_base_cc", "community_small_base_cc", "enzymes_small_base_cc",
                                  "ego_small_cc_v2", "zinc250k", "enzymes_small", "enzymes", "grid", "grid_small_cc"])
def test_scores_against_committed_reference_outputs(name):
    """The same inputs the unmodified reference was run on (tests/golden/io_<cfg>.npz)."""
    cfg = Config(name)
    io = cfg.io()
    flags, x, adj = (torch.from_numpy(io[k]) for k in ("flags", "x", "adj"))
    B = x.shape[0]
    r2 = None
    if cfg.is_cc:
        g = torch.Generator().manual_seed(int(io["seed_inputs"]))
        torch.randint(max(2, cfg.N // 2), cfg.N + 1, (B,), generator=g)
        torch.randn(B, cfg.N, cfg.F, generator=g)
        torch.randn(B, cfg.N, cfg.N, generator=g)
        r2 = O.mask_rank2(torch.randn(B, cfg.E, cfg.K, generator=g) * 0.3, cfg.N, cfg.d_min, cfg.d_max, flags)
    eng = make_engine(cfg, B, DEV)
    for w, k in enumerate(cfg.keys):
        out = eng.score(w, x, adj, r2, flags).cpu()
        assert check_compressed(io, f"net_{k}", out, SCORE_TOL) < SCORE_TOL, (name, k)


@pytest.mark.parametrize("name,sampler,pred,corr,B,steps", [
    ("qm9", "PC", "Reverse", "Langevin", 32, 4),
    ("qm9", "PC", "Euler", "None", 32, 4),
    ("qm9", "S4", "None", "None", 32, 4),
    ("community_small", "PC", "Euler", "Langevin", 8, 3),
    ("qm9_cc", "PC", "Reverse", "Langevin", 8, 4),
    ("qm9_cc", "PC", "Euler", "Langevin", 8, 3),
    ("qm9_cc", "S4", "None", "None", 8, 4),
    ("enzymes_small_cc", "S4", "None", "None", 4, 3),
    ("enzymes_small_cc", "PC", "Reverse", "Langevin", 4, 2),
    ("community_small_cc", "PC", "Euler", "Langevin", 2, 2),
    ("ego_small", "PC", "Euler", "None", 8, 3),
    ("qm9_base_cc", "PC", "Reverse", "Langevin", 8, 1),
    ("qm9_base_cc", "PC", "Reverse", "Langevin", 8, 3),
    ("community_small_base_cc", "PC", "Euler", "Langevin", 2, 2),
    ("enzymes_small_base_cc", "S4", "None", "None", 4, 2),
    ("ego_small_cc_v2", "PC", "Euler", "None", 2, 2),
    ("zinc250k", "PC", "Reverse", "Langevin", 8, 3),
    ("enzymes_small", "S4", "None", "None", 8, 3),
    ("enzymes", "PC", "Reverse", "Langevin", 4, 3),
    ("enzymes", "S4", "None", "None", 4, 2),
    ("grid", "PC", "Reverse", "Langevin", 2, 2),
    ("grid_small_cc", "PC", "Reverse", "Langevin", 1, 1),
    ("ego_small_cc", "PC", "Euler", "None", 2, 2),
    ("ego_small_cc", "PC", "Reverse", "Langevin", 2, 2),
])
def test_sampler_steps_with_injected_noise(name, sampler, pred, corr, B, steps):
    res = sampler_parity(name, sampler, pred, corr, B, steps, DEV)
    tol = 1e-4   # flat bar (the ill-conditioned non-affine ScoreNetworkF of the Base_CC checkpoints keeps its rank-2 contractions in fp32)
    for k, (e_ret, e_state, agree) in res.items():
        assert e_ret < tol and e_state < tol, (name, k, e_ret, e_state)
        assert agree >= 0.999, (name, k, agree)


@pytest.mark.parametrize("pred,corr,pf,kind", [
    ("Reverse", "Langevin", False, "subVP"),
    ("Euler", "None", False, "subVP"),
    ("Reverse", "Langevin", True, "VP"),
    ("Euler", "Langevin", True, "VE"),
])
def test_sampler_sde_variants(pred, corr, pf, kind):
    """SDE kinds / probability flow that no shipped config selects (subVP: std without the square root and the
    Euler-type discretisation it inherits, sde.py:746, 93-111; probability flow: sde.py:204-235)."""
    res = sampler_parity("qm9_cc", "PC", pred, corr, 4, 3, DEV, probability_flow=pf, sde_kind=kind)
    for k, (e_ret, e_state, agree) in res.items():
        assert e_ret < 1e-4 and e_state < 1e-4, (pred, corr, pf, kind, k, e_ret, e_state)


@pytest.mark.parametrize("name,pred,n_lang,B", [("qm9", "Reverse", 2, 32), ("qm9_cc", "Reverse", 2, 8), ("qm9_cc", "Euler", 3, 8),
                                                ("enzymes_small_cc", "Reverse", 2, 4), ("community_small_cc", "Euler", 2, 2),
                                                ("enzymes", "Reverse", 2, 2)])
def test_langevin_inner_steps(name, pred, n_lang, B):
    """Langevin n_steps > 1 (solver.py:692-701, 760-785): per-object inner loops, the other objects at their
    pre-corrector values, the reference's draw order."""
    res = sampler_parity(name, "PC", pred, "Langevin", B, 2, DEV, n_steps=n_lang)
    for k, (e_ret, e_state, agree) in res.items():
        assert e_ret < 1e-4 and e_state < 1e-4, (name, k, e_ret, e_state)
        assert agree >= 0.999, (name, k, agree)


@pytest.mark.parametrize("name,B", [("synth_gmh_mlpconv", 13), ("synth_gmh_mlpconv2", 9)])
def test_gmh_and_mlp_conv_variants(name, B):
    """ScoreNetworkX_GMH (ScoreNetwork_X.py:156-341) and conv == "MLP" attention (attention.py:170-180): the reference's own
    classes with default init (fixtures from tests/golden/make_golden.py SYNTH): committed reference outputs, oracle parity
    of the scores, sampler steps with injected noise."""
    cfg = Config(name)
    io = cfg.io()
    flags, x, adj = (torch.from_numpy(io[k]) for k in ("flags", "x", "adj"))
    eng = make_engine(cfg, x.shape[0], DEV)
    for w, k in enumerate(cfg.keys):
        assert check_compressed(io, f"net_{k}", eng.score(w, x, adj, None, flags).cpu(), SCORE_TOL) < SCORE_TOL, (name, k)
    for k, e in score_parity(name, B, DEV).items():
        assert e < SCORE_TOL, (name, k, e)
    res = sampler_parity(name, "PC", "Reverse", "Langevin", B, 3, DEV)
    for k, (e_ret, e_state, agree) in res.items():
        assert e_ret < 1e-4 and e_state < 1e-4, (name, k, e_ret, e_state)


def test_longer_horizon_with_injected_noise():
    """60 sampler iterations (360 network evaluations) on the real schedule with the reference's own noise
    stream: the per-step error (~1e-6) must not compound into a different sample -- quantised adjacency and
    incidence agree on >= 99.9 % of the entries (north_star) and the states stay within 1e-3."""
    res = sampler_parity("qm9_cc", "PC", "Reverse", "Langevin", 4, 60, DEV)
    for k, (e_ret, e_state, agree) in res.items():
        assert e_ret < 1e-3 and e_state < 1e-3, (k, e_ret, e_state)
        assert agree >= 0.999, (k, agree)


def test_not_denoised_returns_state():
    res = sampler_parity("qm9_cc", "PC", "Reverse", "Langevin", 4, 2, DEV, denoise=False)
    for k, (e_ret, e_state, _) in res.items():
        assert e_ret < 1e-4 and e_state < 1e-4


def _flags(cfg, B, seed=0):
    g = torch.Generator().manual_seed(seed)
    n = torch.randint(max(2, cfg.N // 2), cfg.N + 1, (B,), generator=g)
    return (torch.arange(cfg.N)[None, :] < n[:, None]).to(torch.float32)


def _check_invariants(cfg, outs, flags):
    """Properties every sampler output has in the reference: finite, adjacency symmetric with zero
    diagonal, everything outside the node mask exactly zero."""
    x, adj = outs[0].cpu(), outs[1].cpu()
    assert torch.isfinite(x).all() and torch.isfinite(adj).all()
    assert torch.equal(adj, adj.transpose(-1, -2))
    assert adj.diagonal(dim1=-2, dim2=-1).abs().max() == 0
    assert torch.equal(x, O.mask_x(x, flags)) and torch.equal(adj, O.mask_adjs(adj, flags))
    if cfg.is_cc:
        r2 = outs[2].cpu()
        assert torch.isfinite(r2).all()
        assert torch.equal(r2, O.mask_rank2(r2, cfg.N, cfg.d_min, cfg.d_max, flags))


def test_full_run_qm9_cc_philox_properties():
    """Whole 1000-step PC run (shipped QM9_CC sampler) with Philox noise: invariants, plausible
    statistics, diff_traj contract, n_evals."""
    cfg = Config("qm9_cc")
    B = 256
    flags = _flags(cfg, B)
    sd = cfg.sdes()
    sh = cfg.shipped
    fn = get_pc_sampler(sd[0], sd[1], cfg.shapes(B)[0], cfg.shapes(B)[1], predictor="Reverse", corrector="Langevin",
                        snr=sh["snr"], scale_eps=sh["scale_eps"], n_steps=1, continuous=True, denoise=True, eps=1e-4,
                        device=DEV, is_cc=True, sde_rank2=sd[2], shape_rank2=cfg.shapes(B)[2], d_min=cfg.d_min,
                        d_max=cfg.d_max)
    x, adj, r2, n, traj = fn(*cfg.holders, flags.to(DEV), seed=11)
    assert n == 2000 and len(traj) == 1000 and len(traj[0]) == 3
    _check_invariants(cfg, (x, adj, r2), flags)
    assert torch.allclose(traj[-1][1], adj[0]) and torch.allclose(traj[-1][0], x[0])
    q = quantize(adj, mol=True).cpu()
    assert q.max() <= 3
    dens = (q > 0).float().mean().item()
    assert 0.01 < dens < 0.6, dens  # molecules: sparse bonded graphs
    # deterministic given the seed, different for another seed
    x2, adj2, r22, _, _ = fn(*cfg.holders, flags.to(DEV), seed=11)
    assert torch.equal(adj2, adj) and torch.equal(r22, r2)
    x3, adj3, _, _, _ = fn(*cfg.holders, flags.to(DEV), seed=12)
    assert not torch.equal(adj3, adj)


def test_full_size_community_small_cc_invariants():
    """BASELINE configs[1] at full size (B=1024, E=190, K=1140), 30 steps on the real schedule."""
    cfg = Config("community_small_cc")
    B = 1024
    flags = _flags(cfg, B)
    sd = cfg.sdes()
    sh = cfg.shipped
    fn = get_pc_sampler(sd[0], sd[1], cfg.shapes(B)[0], cfg.shapes(B)[1], predictor="Euler", corrector="Langevin",
                        snr=sh["snr"], scale_eps=sh["scale_eps"], n_steps=1, continuous=True, denoise=True, eps=1e-4,
                        device=DEV, is_cc=True, sde_rank2=sd[2], shape_rank2=cfg.shapes(B)[2], d_min=cfg.d_min,
                        d_max=cfg.d_max)
    outs = fn(*cfg.holders, flags.to(DEV), seed=3, max_steps=30, record_traj=False)
    _check_invariants(cfg, outs[:3], flags)


def test_shard_invariance_without_batch_coupling():
    """Philox is keyed by the GLOBAL sample index, so with no Langevin batch mean (corrector None)
    a batch run as one piece or as shards with sample_offset gives the same samples.  Shards whose sizes
    are multiples of every kernel's work-group size (tc_apply packs 192 // E = 5 consecutive QM9_CC samples
    per group, tc_attn 128 // N = 14 graphs per MMA tile: lcm 70) reproduce the whole-batch run bit for bit;
    other splits change only the position of a sample inside its group, i.e. the fp32 summation order of
    the block-diagonal products (1e-6 relative per step)."""
    cfg = Config("qm9_cc")
    sd = cfg.sdes()
    unit = 70
    flags = _flags(cfg, 2 * unit)

    def run(fl, off):
        B = fl.shape[0]
        fn = get_pc_sampler(sd[0], sd[1], cfg.shapes(B)[0], cfg.shapes(B)[1], predictor="Reverse", corrector="None",
                            continuous=True, denoise=True, eps=1e-4, device=DEV, is_cc=True, sde_rank2=sd[2],
                            shape_rank2=cfg.shapes(B)[2], d_min=cfg.d_min, d_max=cfg.d_max)
        return fn(*cfg.holders, fl.to(DEV), seed=5, sample_offset=off, max_steps=20, record_traj=False)[:3]

    whole = run(flags, 0)
    a, b = run(flags[:unit], 0), run(flags[unit:], unit)          # group-aligned shards
    for w, p, q in zip(whole, a, b):
        assert torch.equal(w, torch.cat([p, q]))
    a, b = run(flags[:4], 0), run(flags[4:], 4)                    # unaligned shards
    for w, p, q in zip(whole, a, b):
        assert rel_err(torch.cat([p, q]), w) < 1e-4


def test_philox_noise_is_standard_normal_and_masked():
    """Prior sampling through the kernels: moments of the unmasked entries, symmetry, masks."""
    cfg = Config("qm9_cc")
    B = 512
    flags = _flags(cfg, B)
    eng = make_engine(cfg, B, DEV)
    eng.init(flags, seed=123)
    x, adj, r2 = [t.cpu() for t in eng.read(False)]
    _check_invariants(cfg, (x, adj, r2), flags)
    for t, m in ((x, O.mask_x(torch.ones_like(x), flags)), (r2, O.mask_rank2(torch.ones_like(r2), cfg.N, cfg.d_min, cfg.d_max, flags))):
        v = t[m > 0]
        assert abs(v.mean().item()) < 0.02 and abs(v.std().item() - 1.0) < 0.02
        assert abs((v ** 4).mean().item() - 3.0) < 0.3  # kurtosis of a Gaussian
    iu = torch.triu_indices(cfg.N, cfg.N, 1)
    v = adj[:, iu[0], iu[1]][O.mask_adjs(torch.ones_like(adj), flags)[:, iu[0], iu[1]] > 0]
    assert abs(v.mean().item()) < 0.05 and abs(v.std().item() - 1.0) < 0.05


def test_s4_factory_and_quantize():
    cfg = Config("enzymes_small_cc")
    B = 16
    flags = _flags(cfg, B)
    sd = cfg.sdes()
    fn = S4_solver(sd[0], sd[1], cfg.shapes(B)[0], cfg.shapes(B)[1], snr=0.15, scale_eps=0.7, continuous=True,
                   denoise=True, eps=1e-4, device=DEV, is_cc=True, sde_rank2=sd[2], shape_rank2=cfg.shapes(B)[2],
                   d_min=cfg.d_min, d_max=cfg.d_max)
    x, adj, r2, n, traj = fn(*cfg.holders, flags.to(DEV), seed=1, max_steps=10)
    assert n == 0 and len(traj) == 10 and len(traj[0]) == 3   # one [x[0], adj[0], rank2[0]] entry per executed step
    _check_invariants(cfg, (x, adj, r2), flags)
    q = quantize(adj).cpu()
    assert torch.equal(q.float(), O.quantize(adj.cpu()))
    qm = quantize(adj * 3, mol=True).cpu()
    assert torch.equal(qm.long(), O.quantize_mol(adj.cpu() * 3))


def test_mol_onehot_on_device():
    """Molecule post-processing (sampler.py:814-825) on the device: bit exact against the reference's tensor code, at the
    QM9 bench batch and at ZINC250k's shape."""
    from ccsd_b200.solver import mol_onehot
    for B, N, F in ((10000, 9, 4), (2048, 38, 9)):
        g = torch.Generator().manual_seed(B)
        x = torch.rand(B, N, F, generator=g) * 1.2
        adj = torch.rand(B, N, N, generator=g) * 3.4 - 0.2
        xo, ao = mol_onehot(x.to(DEV), adj.to(DEV))
        xr, ar = O.mol_onehot(x, adj)
        assert torch.equal(xo.cpu(), xr) and torch.equal(ao.cpu(), ar)


@pytest.mark.parametrize("name,sampler,pred,B", [("community_small", "PC", "Euler", 37), ("qm9", "PC", "Reverse", 64),
                                                 ("enzymes_small", "S4", "Euler", 16), ("enzymes", "PC", "Reverse", 2)])
def test_graph_replay_equals_eager(name, sampler, pred, B, monkeypatch):
    """Graph-only plans replay ONE captured step with the step index / diff_traj slots read from device memory
    (ccsd_plan_run, StepDev): the same kernels with the same Philox counters, so state, means and the recorded
    trajectory must equal the eager launch sequence bit for bit."""
    cfg = Config(name)
    g = torch.Generator().manual_seed(3)
    n = torch.randint(max(2, cfg.N // 2), cfg.N + 1, (B,), generator=g)
    flags = (torch.arange(cfg.N)[None, :] < n[:, None]).float().to(DEV)
    outs = []
    for no_graph in (False, True):
        if no_graph:
            monkeypatch.setenv("CCSD_B200_NO_GRAPH", "1")
        else:
            monkeypatch.delenv("CCSD_B200_NO_GRAPH", raising=False)
        eng = make_engine(cfg, B, DEV, sampler=sampler, predictor=pred)
        assert int(eng.lib.ccsd_plan_info(eng.handle, 19)) == (0 if no_graph else 1)
        eng.enable_traj()
        eng.init(flags, seed=11)
        eng.run(0, 36)
        torch.cuda.synchronize()
        outs.append([t.clone() for t in eng.read(False)] + [t.clone() for t in eng.read(True)] + [t[:36].clone() for t in eng.traj]
                    + [eng.launches])
    assert outs[0][-1] == outs[1][-1] + 34, (outs[0][-1], outs[1][-1])   # one step_advance per replayed step
    for a, b in zip(outs[0][:-1], outs[1][:-1]):
        assert torch.equal(a, b)
