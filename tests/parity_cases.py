"""Parity cases shared by the CPU (host-emulation) and GPU test files: the kernels behind the C ABI
versus the oracle on the same seeded inputs and the same injected noise stream."""
from __future__ import annotations

import torch

from ccsd_b200.solver import Engine, InjectedNoise
from oracle import ccsd_oracle as O
from tests.helpers import Config, rel_err

SCORE_TOL = 1e-4  # BASELINE.json north_star: per-step score outputs within 1e-4 relative


def make_engine(cfg: Config, B: int, device: str, sampler="PC", predictor="Euler", corrector="Langevin", snr=None,
                scale_eps=None, denoise=True, probability_flow=False, sdes=None, n_steps=1):
    sh = cfg.shipped
    return Engine(cfg.holders, sdes or cfg.sdes(), cfg.shapes(B), sampler=sampler, predictor=predictor, corrector=corrector,
                  snr=sh["snr"] if snr is None else snr, scale_eps=sh["scale_eps"] if scale_eps is None else scale_eps,
                  n_steps=n_steps, denoise=denoise, eps=1e-4, device=device, d_min=cfg.d_min, d_max=cfg.d_max,
                  probability_flow=probability_flow)


def sdes_of_kind(cfg: Config, kind: str):
    """The config's SDEs with their type replaced (subVP has no shipped checkpoint; the networks do not care)."""
    s = cfg.meta["sde"]
    return [O.make_sde(kind, s[k]["beta_min"], s[k]["beta_max"], s[k]["num_scales"]) for k in cfg.keys]


def score_parity(name: str, B: int, device: str, seed: int = 1):
    """model(x, adj[, rank2], flags) of every network vs the oracle.  Returns {net: rel err}."""
    cfg = Config(name)
    x, adj, r2, flags = cfg.random_state(B, seed)
    eng = make_engine(cfg, B, device)
    args = (x, adj, r2, flags) if cfg.is_cc else (x, adj, flags)
    errs = {}
    for w, (k, m) in enumerate(zip(cfg.keys, cfg.oracle_models)):
        ref = m(*args)
        out = eng.score(w, x, adj, r2, flags).cpu()
        assert torch.isfinite(out).all()
        errs[k] = rel_err(out, ref)
    return errs


def sampler_parity(name: str, sampler: str, predictor: str, corrector: str, B: int, steps: int, device: str,
                   seed: int = 5, denoise: bool = True, probability_flow: bool = False, sde_kind: str = None,
                   n_steps: int = 1):
    """`steps` sampler iterations on the real schedule with an injected noise stream.  Returns per
    object (rel err of the returned tensor, rel err of the raw state, quantised agreement)."""
    cfg = Config(name)
    _, _, _, flags = cfg.random_state(B, seed)
    sh = cfg.shipped
    src = O.NoiseSource(seed=seed)
    rec = []
    kw = dict(snr=sh["snr"], scale_eps=sh["scale_eps"], denoise=denoise, eps=1e-4, d_min=cfg.d_min, d_max=cfg.d_max,
              noise=src, max_steps=steps, record=rec)
    sdes = sdes_of_kind(cfg, sde_kind) if sde_kind else cfg.sdes()
    if sampler == "S4":
        res, _ = O.s4_solver(cfg.oracle_models, sdes, cfg.shapes(B), flags, **kw)
    else:
        res, _ = O.pc_sampler(cfg.oracle_models, sdes, cfg.shapes(B), flags, predictor=predictor,
                              corrector=corrector, n_steps=n_steps, probability_flow=probability_flow, **kw)
    eng = make_engine(cfg, B, device, sampler, predictor, corrector, denoise=denoise, probability_flow=probability_flow,
                      sdes=sdes, n_steps=n_steps)
    inj = InjectedNoise.from_flat_log(src.log, len(cfg.keys), eng.n_draws, steps,
                                      n_lang=n_steps if (sampler == "PC" and corrector == "Langevin") else None)
    eng.init(flags, prior=inj.prior)
    for i in range(steps):
        eng.step(i, inj.steps[i])
    ret = [t.cpu() for t in eng.read(denoise)]
    state = [t.cpu() for t in eng.read(False)]
    out = {}
    for k, key in enumerate(cfg.keys):
        agree = 1.0
        if key != "x":
            agree = (O.quantize(ret[k]) == O.quantize(res[k])).float().mean().item()
        out[key] = (rel_err(ret[k], res[k]), rel_err(state[k], rec[-1][0][k]), agree)
    return out
