"""Validate oracle/ccsd_oracle.py against the UNMODIFIED reference imported from
/root/reference (build container only; the GPU box has no /root/reference).

Run:  python -m oracle.validate_against_reference            (prints max |diff| per check)

Checks, for every shipped checkpoint the hot path covers:
  * each score network on random inputs + random prefix flags,
  * 3 sampler steps (PC: every predictor x corrector; S4) on the real 1000-step schedule,
    the reference run with torch.manual_seed(s) and the oracle replaying the same global
    generator stream (randn / randn_like consume it identically).
Test infrastructure only.
"""
from __future__ import annotations

import sys

import torch

from oracle import refstubs

refstubs.install()

from ccsd.src import solver as rsolver  # noqa: E402
from ccsd.src import sde as rsde  # noqa: E402
from ccsd.src.utils import loader as rloader  # noqa: E402

from oracle import ccsd_oracle as O  # noqa: E402

CKPTS = {
    "community_small": "community_small/gdss_community_small",
    "qm9": "QM9/gdss_qm9_retrained",
    "qm9_cc": "QM9/ccsd_qm9_CC",
    "community_small_cc": "community_small_CC/ccsd_community_small_CC",
    "enzymes_small_cc": "ENZYMES_small_CC/ccsd_enzymes_small_CC",
    "ego_small_cc": "ego_small_CC/ccsd_ego_small_CC",
}


def load(name):
    ck = torch.load(f"/root/reference/checkpoints/{CKPTS[name]}.pth", map_location="cpu", weights_only=False)
    is_cc = "params_rank2" in ck
    keys = ["x", "adj"] + (["rank2"] if is_cc else [])
    ref_models, ora_models = [], []
    for k in keys:
        p = dict(ck[f"params_{k}"])
        m = rloader.load_model(dict(p))
        sd = {kk[7:] if kk.startswith("module.") else kk: v for kk, v in ck[f"{k}_state_dict"].items()}
        m.load_state_dict(sd)
        m.eval()
        ref_models.append(m)
        ora_models.append(O.Model(p["model_type"], p, sd, is_cc=is_cc))
    return ck, is_cc, ref_models, ora_models


def rand_flags(B, N, gen):
    n = torch.randint(max(2, N // 2), N + 1, (B,), generator=gen)
    return (torch.arange(N)[None, :] < n[:, None]).to(torch.float32)


def make_sdes(ck, is_cc):
    cfg = ck["model_config"].sde
    keys = ["x", "adj"] + (["rank2"] if is_cc else [])
    ref = [rloader.load_sde(cfg[k]) for k in keys]
    ora = [O.make_sde(cfg[k].type, cfg[k].beta_min, cfg[k].beta_max, cfg[k].num_scales) for k in keys]
    return ref, ora


def check(name, B=3):
    ck, is_cc, rm, om = load(name)
    d = ck["model_config"].data
    N, Fd = d.max_node_num, d.max_feat_num
    d_min, d_max = (d.d_min, d.d_max) if is_cc else (None, None)
    g = torch.Generator().manual_seed(1)
    flags = rand_flags(B, N, g)
    x = O.mask_x(torch.randn(B, N, Fd, generator=g), flags)
    adj = O.mask_adjs(O.symmetrize_noise(torch.randn(B, N, N, generator=g)), flags)
    worst = 0.0
    with torch.no_grad():
        if is_cc:
            E, K = O.rank2_dim(N, d_min, d_max)
            r2 = O.mask_rank2(torch.randn(B, E, K, generator=g) * 0.3, N, d_min, d_max, flags)
            args = (x, adj, r2, flags)
        else:
            args = (x, adj, flags)
        for nm, a, b in zip(["x", "adj", "rank2"], rm, om):
            ra, rb = a(*args), b(*args)
            diff = (ra - rb).abs().max().item()
            rel = diff / (ra.abs().max().item() + 1e-30)
            print(f"  [{name}] net {nm}: max|d|={diff:.3e} rel={rel:.3e}")
            worst = max(worst, rel)
    # samplers
    rs, os_ = make_sdes(ck, is_cc)
    shapes = [(B, N, Fd), (B, N, N)] + ([(B, E, K)] if is_cc else [])
    combos = [("Euler", "Langevin", 1), ("Reverse", "Langevin", 1), ("Reverse", "None", 1), ("S4", "None", 1),
              ("Reverse", "Langevin", 2)]   # last: Langevin n_steps = 2 (per-object inner loops, solver.py:692, 760)
    for pred, corr, n_lang in combos:
        if pred == "S4" and any(isinstance(s, rsde.subVPSDE) for s in rs):
            continue
        kw = dict(
            predictor=pred, corrector=corr, snr=0.15, scale_eps=0.7, n_steps=n_lang, probability_flow=False,
            continuous=True, denoise=True, eps=1e-4, device="cpu",
        )
        if is_cc:
            kw.update(is_cc=True, sde_rank2=rs[2], shape_rank2=shapes[2], d_min=d_min, d_max=d_max)
        fac = rsolver.S4_solver if pred == "S4" else rsolver.get_pc_sampler
        steps = 3
        # run the reference for `steps` iterations only: shrink diff_steps by patching trange
        import ccsd.src.solver as S

        orig = S.trange
        S.trange = lambda a, b, **k: range(a, min(b, steps))
        try:
            fn = fac(rs[0], rs[1], shapes[0], shapes[1], **kw)
            torch.manual_seed(7)
            out = fn(*rm, flags)
        finally:
            S.trange = orig
        torch.manual_seed(7)
        src = O.NoiseSource(seed=None)
        okw = dict(snr=0.15, scale_eps=0.7, denoise=True, eps=1e-4, d_min=d_min, d_max=d_max, noise=src, max_steps=steps)
        if pred == "S4":
            res, _ = O.s4_solver(om, os_, shapes, flags, **okw)
        else:
            res, _ = O.pc_sampler(om, os_, shapes, flags, predictor=pred, corrector=corr, n_steps=n_lang, **okw)
        for nm, a, b in zip(["x", "adj", "rank2"], out[: len(res)], res):
            diff = (a - b).abs().max().item()
            rel = diff / (a.abs().max().item() + 1e-30)
            print(f"  [{name}] {pred}+{corr} n_steps={n_lang} {nm}: max|d|={diff:.3e} rel={rel:.3e}")
            worst = max(worst, rel)
    return worst


if __name__ == "__main__":
    names = sys.argv[1:] or list(CKPTS)
    w = 0.0
    for n in names:
        w = max(w, check(n))
    print("WORST relative difference:", w)
    sys.exit(0 if w < 1e-4 else 1)
