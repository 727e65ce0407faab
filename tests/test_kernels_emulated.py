"""CPU-side check of the kernels' index arithmetic, weight packing and masks: the SAME kernel
sources compiled for the host (one thread per block, tests/_emu) against the oracle.  This is not
a product path -- it only exists so that a box without a GPU can catch layout bugs; the parity
tests proper are tests/test_parity_gpu.py."""
import pytest
import torch

from ccsd_b200 import _native as nat
from tests.parity_cases import SCORE_TOL, sampler_parity, score_parity

pytestmark = pytest.mark.skipif(torch.cuda.is_available(), reason="host emulation is only used where there is no GPU")


def test_emulation_build_is_flagged():
    assert nat.is_emulation()
    assert b"EMULATION" in nat.load().ccsd_version()


@pytest.mark.parametrize("name,B", [("qm9", 3), ("qm9_cc", 2), ("community_small", 2), ("enzymes_small_cc", 1), ("ego_small", 2),
                                    ("qm9_base_cc", 2), ("community_small_base_cc", 1),
                                    ("enzymes_small_base_cc", 1), ("zinc250k", 1), ("enzymes_small", 2),
                                    ("enzymes", 1), ("grid", 1)])   # N = 125, 361: the large-graph pipeline (big_pipe.cuh)
def test_scores(name, B):
    for k, e in score_parity(name, B, "cpu").items():
        assert e < SCORE_TOL, (name, k, e)


def test_scores_community_small_cc():
    for k, e in score_parity("community_small_cc", 1, "cpu").items():
        assert e < SCORE_TOL, (k, e)


@pytest.mark.parametrize("name,sampler,pred,corr", [
    ("qm9", "PC", "Reverse", "Langevin"),
    ("qm9", "PC", "Euler", "None"),
    ("qm9", "S4", "None", "None"),
    ("qm9_cc", "PC", "Reverse", "Langevin"),
    ("qm9_cc", "PC", "Euler", "Langevin"),
    ("qm9_cc", "S4", "None", "None"),
    ("enzymes_small_cc", "S4", "None", "None"),
    ("qm9_base_cc", "PC", "Reverse", "Langevin"),
    ("enzymes_small", "S4", "None", "None"),
    ("enzymes", "PC", "Reverse", "Langevin"),
])
def test_sampler_steps(name, sampler, pred, corr):
    res = sampler_parity(name, sampler, pred, corr, B=2, steps=2, device="cpu")
    for k, (e_ret, e_state, agree) in res.items():
        assert e_ret < 1e-4 and e_state < 1e-4, (name, k, e_ret, e_state)
        assert agree >= 0.999


@pytest.mark.parametrize("name,pred,n_lang", [("qm9", "Reverse", 2), ("qm9_cc", "Reverse", 2), ("qm9_cc", "Euler", 3),
                                              ("enzymes_small_cc", "Reverse", 2)])
def test_langevin_inner_steps(name, pred, n_lang):
    """Langevin n_steps > 1: every object's corrector loops on its own with the other objects at their pre-corrector
    values (solver.py:692-701, 760-785, 1123-1140); noise draws in the reference's order."""
    res = sampler_parity(name, "PC", pred, "Langevin", B=2, steps=2, device="cpu", n_steps=n_lang)
    for k, (e_ret, e_state, agree) in res.items():
        assert e_ret < 1e-4 and e_state < 1e-4, (name, k, e_ret, e_state)
        assert agree >= 0.999


def test_sampler_not_denoised():
    res = sampler_parity("qm9_cc", "PC", "Reverse", "Langevin", B=2, steps=2, device="cpu", denoise=False)
    for k, (e_ret, e_state, _) in res.items():
        assert e_ret < 1e-4 and e_state < 1e-4


@pytest.mark.parametrize("pred,corr,pf,kind", [
    ("Reverse", "Langevin", False, "subVP"),   # subVP: std without the square root, Euler-type discretize (sde.py:746, 93-111)
    ("Euler", "Langevin", False, "subVP"),
    ("Reverse", "None", True, "VP"),           # probability flow: half score term, no diffusion (sde.py:204-235)
    ("Euler", "Langevin", True, "VE"),
])
def test_sampler_sde_variants(pred, corr, pf, kind):
    res = sampler_parity("qm9_cc", "PC", pred, corr, B=2, steps=2, device="cpu", probability_flow=pf, sde_kind=kind)
    for k, (e_ret, e_state, agree) in res.items():
        assert e_ret < 1e-4 and e_state < 1e-4, (pred, corr, pf, kind, k, e_ret, e_state)


def test_large_graph_pipeline_on_small_graphs():
    """CCSD_B200_FORCE_BIG routes graph-only plans with N <= 64 through the large-graph kernels (big_pipe.cuh): same
    checkpoints, same oracle, same bar -- and the two pipelines must agree with each other."""
    import os
    import subprocess
    import sys
    code = (
        "import sys; sys.path.insert(0, '.')\n"
        "import tests.conftest\n"
        "from tests.parity_cases import score_parity, sampler_parity\n"
        "for n, B in (('qm9', 3), ('community_small', 2), ('zinc250k', 1)):\n"
        "    for k, e in score_parity(n, B, 'cpu').items():\n"
        "        assert e < 1e-4, (n, k, e)\n"
        "for args in (('qm9', 'PC', 'Reverse', 'Langevin', 3, 2), ('community_small', 'PC', 'Euler', 'Langevin', 2, 2),\n"
        "             ('enzymes_small', 'S4', 'None', 'None', 2, 2)):\n"
        "    for k, v in sampler_parity(*args, 'cpu').items():\n"
        "        assert v[0] < 1e-4 and v[1] < 1e-4 and v[2] >= 0.999, (args, k, v)\n"
        "from ccsd_b200 import _native as nat\n"
        "print('ok')\n")
    env = dict(os.environ, CCSD_B200_FORCE_BIG="1")
    r = subprocess.run([sys.executable, "-c", code], cwd=str(__import__("pathlib").Path(__file__).resolve().parents[1]), env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr


def test_mol_onehot_matches_reference_postprocessing():
    """ccsd_mol_onehot vs the reference's own tensor code (sampler.py:814-825, restated in the oracle): bit exact."""
    import torch
    from ccsd_b200.solver import mol_onehot
    from oracle import ccsd_oracle as O
    g = torch.Generator().manual_seed(3)
    x = torch.rand(5, 9, 4, generator=g) * 1.2
    adj = torch.rand(5, 9, 9, generator=g) * 3.4 - 0.2
    adj[0, 0, :4] = torch.tensor([0.5, 1.5, 2.5, 0.4999])   # the thresholds themselves
    xo, ao = mol_onehot(x, adj)
    xr, ar = O.mol_onehot(x, adj)
    assert xo.dtype == torch.int64 and ao.dtype == torch.int64 and ao.shape == (5, 4, 9, 9) and xo.shape == (5, 9, 5)
    assert torch.equal(xo, xr) and torch.equal(ao, ar)


@pytest.mark.parametrize("N", [65, 96, 129])
def test_large_graph_pipeline_at_tile_boundaries(N):
    """The score networks have no N-dependent parameters, so the ENZYMES checkpoint (N = 125) is evaluated at other graph
    sizes around the large-graph pipeline's tile sizes (8-row groups, 32-row chunks, 64 / 128-column segments)."""
    import torch
    from tests.helpers import Config, rel_err
    from tests.parity_cases import make_engine
    cfg = Config("enzymes")
    cfg.N = N
    x, adj, r2, flags = cfg.random_state(2, seed=N)
    eng = make_engine(cfg, 2, "cpu")
    for w in (0, 1):
        ref = cfg.oracle_models[w](x, adj, flags)
        out = eng.score(w, x, adj, r2, flags).cpu()
        assert torch.isfinite(out).all()
        assert rel_err(out, ref) < 1e-4, (N, w, rel_err(out, ref))


@pytest.mark.parametrize("name,B", [("synth_gmh_mlpconv", 3), ("synth_gmh_mlpconv2", 2)])
def test_gmh_and_mlp_conv_variants(name, B):
    """ScoreNetworkX_GMH (ScoreNetwork_X.py:156-341) and the conv == "MLP" attention variant (attention.py:170-180): the
    reference's own classes with default init (tests/golden/make_golden.py SYNTH), scores and sampler steps."""
    for k, e in score_parity(name, B, "cpu").items():
        assert e < SCORE_TOL, (name, k, e)
    res = sampler_parity(name, "PC", "Reverse", "Langevin", B=B, steps=2, device="cpu")
    for k, (e_ret, e_state, agree) in res.items():
        assert e_ret < 1e-4 and e_state < 1e-4, (name, k, e_ret, e_state)


def test_gmh_against_committed_reference_outputs():
    import torch
    from tests.helpers import Config, check_compressed
    from tests.parity_cases import make_engine
    for name in ("synth_gmh_mlpconv", "synth_gmh_mlpconv2"):
        cfg = Config(name)
        io = cfg.io()
        flags, x, adj = (torch.from_numpy(io[k]) for k in ("flags", "x", "adj"))
        eng = make_engine(cfg, x.shape[0], "cpu")
        for w, k in enumerate(cfg.keys):
            out = eng.score(w, x, adj, None, flags).cpu()
            assert check_compressed(io, f"net_{k}", out, SCORE_TOL) < SCORE_TOL, (name, k)
