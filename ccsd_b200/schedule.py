"""Host-side schedule: every per-step scalar the kernels need, one row per (step, object).

All samples of a batch share ``t`` (``vec_t = ones(B) * timesteps[i]``, ccsd/src/solver.py:976-977),
so everything the reference computes from ``t`` is a scalar per step and object.  They are computed
here with the reference's own torch fp32 expressions, in the reference's operation order (e.g.
``timestep = (t * (N - 1) / T).long()`` on the fp32 ``linspace``), so the scalars agree with the
reference's -- they are never re-derived inside a kernel.

Columns follow ``ccsd_objcoef_t`` (include/ccsd_b200.h).
"""
from __future__ import annotations

from typing import Sequence

import numpy as np
import torch

from .sde import sde_kind

COLS = ("score_scale", "lg_alpha", "pa", "pb", "pc", "s4_alpha", "s4_m1", "s4_s1", "s4_sd", "s4_m2", "s4_s2", "pad_")


def _tables(sde, kind):
    N = sde.N
    if kind in ("VP", "subVP"):
        betas = torch.linspace(sde.beta_0 / N, sde.beta_1 / N, N)  # sde.py:364, 689
        return betas, 1.0 - betas
    sig = torch.exp(torch.linspace(np.log(sde.sigma_min), np.log(sde.sigma_max), N))  # sde.py:523-525
    return sig, None


def _marginal_std(sde, kind, t):
    if kind == "VE":
        return sde.sigma_min * (sde.sigma_max / sde.sigma_min) ** t  # sde.py:579
    lmc = -0.25 * t ** 2 * (sde.beta_1 - sde.beta_0) - 0.5 * t * sde.beta_0  # sde.py:419-421, 742-744
    if kind == "VP":
        return torch.sqrt(1.0 - torch.exp(2.0 * lmc))
    return 1 - torch.exp(2.0 * lmc)  # subVP: no sqrt (sde.py:746)


def _sde_coeffs(sde, kind, t):
    """(c_d, g): drift = c_d * x, diffusion g.  sde.py:387-404, 545-565, 709-728."""
    if kind == "VE":
        sigma = sde.sigma_min * (sde.sigma_max / sde.sigma_min) ** t
        g = sigma * torch.sqrt(torch.tensor(2 * (np.log(sde.sigma_max) - np.log(sde.sigma_min))))
        return torch.zeros_like(t), g
    beta_t = sde.beta_0 + t * (sde.beta_1 - sde.beta_0)
    if kind == "VP":
        return -0.5 * beta_t, torch.sqrt(beta_t)
    discount = 1.0 - torch.exp(-2 * sde.beta_0 * t - (sde.beta_1 - sde.beta_0) * t ** 2)
    return -0.5 * beta_t, torch.sqrt(beta_t * discount)


def _transition(sde, kind, t, dt):
    """(m, std): mean = m * x.  sde.py:485-503, 650-669."""
    if kind == "VP":
        lmc = 0.25 * dt * (2 * sde.beta_0 + (2 * t + dt) * (sde.beta_1 - sde.beta_0))
        return torch.exp(-lmc), torch.sqrt(1.0 - torch.exp(2.0 * lmc))
    if kind == "VE":
        std = torch.square(sde.sigma_min * (sde.sigma_max / sde.sigma_min) ** t) - torch.square(
            sde.sigma_min * (sde.sigma_max / sde.sigma_min) ** (t + dt)
        )
        return torch.ones_like(t), torch.sqrt(std)
    raise NotImplementedError("subVPSDE has no transition(); the S4 solver does not support it (sde.py:672-786)")


def build_schedule(
    sdes: Sequence, *, sampler: str, predictor: str, probability_flow: bool, eps: float
) -> np.ndarray:
    """Return a float32 array [n_diff_steps, 3, 12] (object order x, adj, rank2; rank2 row is a copy
    of adj's when the plan is graph-only).

    sampler: "PC" or "S4"; predictor: "Reverse" | "Euler" (PC only).
    """
    sde_adj = sdes[1]
    n_steps = int(sde_adj.N)
    timesteps = torch.linspace(sde_adj.T, eps, n_steps)  # solver.py:969-970 / 1119-1120
    out = np.zeros((n_steps, 3, len(COLS)), np.float32)
    pf = 0.5 if probability_flow else 1.0
    dt_s4 = -1.0 / n_steps  # solver.py:1277, 1421
    for o, sde in enumerate(sdes):
        kind = sde_kind(sde)
        N, T = int(sde.N), sde.T
        t = timesteps.clone()
        tab, alphas = _tables(sde, kind)
        col = {}
        std = _marginal_std(sde, kind, t)
        col["score_scale"] = torch.ones_like(t) if kind == "VE" else -1.0 / std  # losses.py:66-70, 95-99
        idx = (t * (N - 1) / T).long()  # sde.py:477, 639; solver.py:685, 753
        col["lg_alpha"] = alphas[idx] if kind in ("VP", "subVP") else torch.ones_like(t)
        c_d, g = _sde_coeffs(sde, kind, t)
        if sampler == "PC":
            if predictor == "Reverse":
                if kind == "VP":  # sde.py:465-483
                    beta, alpha = tab[idx], alphas[idx]
                    G = torch.sqrt(beta)
                    a_coef = 1.0 - (torch.sqrt(alpha) - 1.0)  # x - (sqrt(alpha) x - x)
                elif kind == "VE":  # sde.py:625-648
                    sigma = tab[idx]
                    adjacent = torch.where(idx == 0, torch.zeros_like(t), tab[idx - 1])
                    G = torch.sqrt(sigma ** 2 - adjacent ** 2)
                    a_coef = torch.ones_like(t)
                else:  # subVP inherits SDE.discretize (Euler), sde.py:93-111
                    dt = 1 / N
                    G = g * torch.sqrt(torch.tensor(dt))
                    a_coef = 1.0 - c_d * dt
                col["pa"] = a_coef
                col["pb"] = G ** 2 * pf  # mean = x - (f - G^2 score pf)   (sde.py:229-235, solver.py:388)
                col["pc"] = torch.zeros_like(G) if probability_flow else G
            elif predictor == "Euler":  # solver.py:227-244; sde.py:200-207
                dt = -1.0 / N
                col["pa"] = 1.0 + c_d * dt
                col["pb"] = -(g ** 2) * pf * dt
                col["pc"] = torch.zeros_like(g) if probability_flow else g * np.sqrt(-dt)
            else:
                raise NotImplementedError(f"Predictor {predictor} not yet supported. Select from [Reverse, Euler].")
        else:  # S4, solver.py:1287-1353 / 1431-1534
            idx_x = (t * (int(sdes[0].N) - 1) / sdes[0].T).long()  # solver.py:1297, 1446 (sde_x for all)
            col["s4_alpha"] = alphas[idx_x] if kind == "VP" else torch.ones_like(t)  # VPSDE only (:1306, 1455)
            vec_dt = torch.ones_like(t) * (dt_s4 / 2)
            m1, s1 = _transition(sde, kind, t, vec_dt)
            m2, s2 = _transition(sde, kind, t + vec_dt, vec_dt)
            col["s4_m1"], col["s4_s1"], col["s4_m2"], col["s4_s2"] = m1, s1, m2, s2
            col["s4_sd"] = -(g ** 2) * dt_s4
        for name, v in col.items():
            out[:, o, COLS.index(name)] = v.to(torch.float32).numpy()
    if len(sdes) == 2:
        out[:, 2, :] = out[:, 1, :]
    return out
