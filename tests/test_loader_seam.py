"""The factory seam of the reference (ccsd/src/utils/loader.py: load_sde, load_sampling_fn, load_model,
load_model_from_ckpt) and the module mirrors of ccsd_b200/models.py, driven the way Sampler_*.sample() drives
them (ccsd/src/sampler.py:138, 415, 727, 1104) with reference-shaped configs.  Runs on the B200 under ``-m gpu``
(product library) and, where there is no GPU, through the host-emulation build."""
import copy

import pytest
import torch

from ccsd_b200 import loader
from ccsd_b200.solver import InjectedNoise
from oracle import ccsd_oracle as O
from tests.helpers import Config, rel_err

HAS_CUDA = torch.cuda.is_available()
DEV = "cuda" if HAS_CUDA else "cpu"


def _configs(cfg: Config, B: int, predictor: str, corrector: str, n_steps: int = 1):
    """config_train / config_module / config_sample with the field names of config/sample_*.yaml + the checkpoint config."""
    s = cfg.meta["sde"]
    train = {"sde": {k: dict(s[k]) for k in cfg.keys},
             "data": {"data": "QM9" if cfg.name.startswith("qm9") else cfg.name, "batch_size": B, "max_node_num": cfg.N,
                      "max_feat_num": cfg.F}}
    module = {"predictor": predictor, "corrector": corrector, "snr": cfg.shipped["snr"], "scale_eps": cfg.shipped["scale_eps"],
              "n_steps": n_steps}
    sample = {"probability_flow": False, "noise_removal": True, "eps": 1e-4, "n_samples": B}
    return train, module, sample


def _models(cfg: Config, prefix: bool = False):
    out = []
    for k, h in zip(cfg.keys, cfg.holders):
        sd = {("module." + n if prefix else n): t.clone() for n, t in h.state_dict().items()}
        out.append(loader.load_model_from_ckpt(dict(cfg.meta["params"][k]), sd, DEV))
    return out


def _seam_forward(name: str, B: int):
    cfg = Config(name)
    x, adj, r2, flags = cfg.random_state(B, seed=11)
    models = _models(cfg, prefix=True)   # DataParallel-style keys are stripped (loader.py:635-637)
    args = [t.to(DEV) for t in ((x, adj, r2, flags) if cfg.is_cc else (x, adj, flags))]
    ref_args = (x, adj, r2, flags) if cfg.is_cc else (x, adj, flags)
    for m, om, k in zip(models, cfg.oracle_models, cfg.keys):
        out = m(*args).cpu()
        assert rel_err(out, om(*ref_args)) < 1e-4, (name, k)
    return cfg, models, args, ref_args


def _seam_sampler(name: str, B: int, predictor: str, corrector: str, steps: int = 2):
    cfg = Config(name)
    train, module, sample = _configs(cfg, B, predictor, corrector)
    fn = loader.load_sampling_fn(train, module, sample, [0] if HAS_CUDA else "cpu", is_cc=cfg.is_cc, d_min=cfg.d_min, d_max=cfg.d_max)
    models = _models(cfg)
    _, _, _, flags = cfg.random_state(B, seed=4)
    src = O.NoiseSource(seed=4)
    kw = dict(snr=cfg.shipped["snr"], scale_eps=cfg.shipped["scale_eps"], denoise=True, eps=1e-4, d_min=cfg.d_min, d_max=cfg.d_max,
              noise=src, max_steps=steps)
    if predictor == "S4":
        ref, _ = O.s4_solver(cfg.oracle_models, cfg.sdes(), cfg.shapes(B), flags, **kw)
        n_draws = 3
    else:
        ref, _ = O.pc_sampler(cfg.oracle_models, cfg.sdes(), cfg.shapes(B), flags, predictor=predictor, corrector=corrector, n_steps=1, **kw)
        n_draws = 2 if corrector == "Langevin" else 1
    inj = InjectedNoise.from_flat_log(src.log, len(cfg.keys), n_draws, steps)
    out = fn(*models, flags.to(DEV), noise=inj, max_steps=steps)
    n_obj = len(cfg.keys)
    for k in range(n_obj):
        assert rel_err(out[k].cpu(), ref[k]) < 1e-4, (name, cfg.keys[k])
    n_total = cfg.sdes()[1].N
    assert out[n_obj] == (0 if predictor == "S4" else n_total * 2)          # solver.py:1001, 1172, 1369, 1559
    assert len(out[n_obj + 1]) == steps and len(out[n_obj + 1][0]) == n_obj  # diff_traj: [x[0], adj[0](, rank2[0])] per step
    return fn, models, flags, out


def _weights_are_live(name: str, B: int):
    """An in-place weight change between two calls (EMA copy_to writes through .data, ccsd/src/utils/ema.py) is seen."""
    cfg, models, args, ref_args = _seam_forward(name, B)
    m = models[1]
    before = m(*args).clone()
    with torch.no_grad():
        for p in m.final.parameters():
            p.data.mul_(1.5)          # .data: the version counter does not move
    after = m(*args)
    assert rel_err(after, before) > 1e-3
    hp = cfg.meta["params"]["adj"]
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    ref = O.Model(hp["model_type"], hp, sd, is_cc=cfg.is_cc)(*ref_args)
    assert rel_err(after.cpu(), ref) < 1e-4


def _traj_is_a_copy():
    fn, models, flags, out1 = _seam_sampler("qm9", 6, "Reverse", "Langevin")
    traj1 = [[t.clone() for t in step] for step in out1[3]]
    out2 = fn(*models, flags.to(DEV), seed=99, max_steps=2)          # same factory call, same cached plan
    for a, b in zip(traj1, out1[3]):
        for u, v in zip(a, b):
            assert torch.equal(u, v)                                  # the first call's diff_traj did not change
    assert not torch.equal(out2[3][0][1], out1[3][0][1])


def test_load_sde_kinds():
    assert type(loader.load_sde({"type": "VP", "beta_min": 0.1, "beta_max": 1.0, "num_scales": 10})).__name__ == "VPSDE"
    ve = loader.load_sde({"type": "VE", "beta_min": 0.2, "beta_max": 1.0, "num_scales": 10})
    assert (ve.sigma_min, ve.sigma_max) == (0.2, 1.0)                  # loader.py:261-262
    assert type(loader.load_sde({"type": "subVP", "beta_min": 0.1, "beta_max": 1.0, "num_scales": 10})).__name__ == "subVPSDE"
    with pytest.raises(NotImplementedError):
        loader.load_sde({"type": "XX", "beta_min": 0.1, "beta_max": 1.0, "num_scales": 10})


def test_load_model_rejects_unknown_types():
    with pytest.raises(ValueError):
        loader.load_model({"model_type": "Nope"})


CASES_FWD = [("qm9", 5), ("qm9_cc", 3), ("qm9_base_cc", 2), ("synth_gmh_mlpconv", 4)]   # last: load_model("ScoreNetworkX_GMH"), conv="MLP"
CASES_SMP = [("qm9", 6, "Reverse", "Langevin"), ("qm9_cc", 3, "Reverse", "Langevin"), ("qm9_cc", 3, "S4", "None")]


@pytest.mark.skipif(HAS_CUDA, reason="host-emulation variant")
@pytest.mark.parametrize("name,B", CASES_FWD)
def test_model_forward_seam_emulated(name, B):
    _seam_forward(name, B)


@pytest.mark.skipif(HAS_CUDA, reason="host-emulation variant")
@pytest.mark.parametrize("name,B,pred,corr", CASES_SMP)
def test_load_sampling_fn_seam_emulated(name, B, pred, corr):
    _seam_sampler(name, B, pred, corr)


@pytest.mark.skipif(HAS_CUDA, reason="host-emulation variant")
def test_weights_are_live_and_traj_is_a_copy_emulated():
    _weights_are_live("qm9", 4)
    _traj_is_a_copy()


@pytest.mark.gpu
@pytest.mark.parametrize("name,B", CASES_FWD + [("community_small_cc", 7), ("enzymes_small_cc", 11), ("grid", 2)])
def test_model_forward_seam_gpu(name, B):
    _seam_forward(name, B)


@pytest.mark.gpu
@pytest.mark.parametrize("name,B,pred,corr", CASES_SMP + [("community_small_cc", 7, "Euler", "Langevin"), ("enzymes_small_cc", 11, "S4", "None")])
def test_load_sampling_fn_seam_gpu(name, B, pred, corr):
    _seam_sampler(name, B, pred, corr)


@pytest.mark.gpu
def test_weights_are_live_and_traj_is_a_copy_gpu():
    _weights_are_live("qm9_cc", 8)
    _traj_is_a_copy()


@pytest.mark.gpu
def test_second_device_if_present():
    """A plan on cuda:1 launches on cuda:1 (no ABI call selected the device)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("single-GPU box")
    cfg = Config("qm9")
    x, adj, _, flags = cfg.random_state(4, seed=2)
    m = loader.load_model_from_ckpt(dict(cfg.meta["params"]["adj"]), cfg.holders[1].state_dict(), [1])
    out = m(x.to("cuda:1"), adj.to("cuda:1"), flags.to("cuda:1"))
    assert out.device.index == 1 and rel_err(out.cpu(), cfg.oracle_models[1](x, adj, flags)) < 1e-4
