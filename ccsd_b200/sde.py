"""SDE objects with the reference's attribute surface (ccsd/src/sde.py:345-786).

The sampler factories accept the reference's own ``VPSDE`` / ``VESDE`` / ``subVPSDE`` instances
(duck-typed on class name and the attributes below); these mirrors exist so that
``ccsd_b200.loader.load_sde`` works where the reference package is not importable.  They carry the
coefficient tables and scalar maths only -- the per-step scalars the kernels consume are derived in
``ccsd_b200.schedule`` with the same torch fp32 expressions the reference evaluates.
"""
from __future__ import annotations

import numpy as np
import torch


class SDE:
    def __init__(self, N: int) -> None:
        self.N = N

    @property
    def T(self) -> int:
        return 1


class VPSDE(SDE):
    """ccsd/src/sde.py:345-503."""

    def __init__(self, beta_min: float = 0.1, beta_max: float = 20.0, N: int = 1000) -> None:
        super().__init__(N)
        self.beta_0, self.beta_1 = beta_min, beta_max
        self.discrete_betas = torch.linspace(beta_min / N, beta_max / N, N)
        self.alphas = 1.0 - self.discrete_betas

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}(N={self.N}, beta_min={self.beta_0}, beta_max={self.beta_1}, T={self.T})"


class VESDE(SDE):
    """ccsd/src/sde.py:506-669."""

    def __init__(self, sigma_min: float = 0.01, sigma_max: float = 50.0, N: int = 1000) -> None:
        super().__init__(N)
        self.sigma_min, self.sigma_max = sigma_min, sigma_max
        self.discrete_sigmas = torch.exp(torch.linspace(np.log(sigma_min), np.log(sigma_max), N))

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}(N={self.N}, sigma_min={self.sigma_min}, sigma_max={self.sigma_max}, T={self.T})"


class subVPSDE(SDE):
    """ccsd/src/sde.py:672-786."""

    def __init__(self, beta_min: float = 0.1, beta_max: float = 20.0, N: int = 1000) -> None:
        super().__init__(N)
        self.beta_0, self.beta_1 = beta_min, beta_max
        self.discrete_betas = torch.linspace(beta_min / N, beta_max / N, N)
        self.alphas = 1.0 - self.discrete_betas

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}(N={self.N}, beta_min={self.beta_0}, beta_max={self.beta_1}, T={self.T})"


def sde_kind(sde) -> str:
    """'VP' | 'VE' | 'subVP' from a reference or mirror SDE object; NotImplementedError otherwise
    (matches ccsd/src/losses.py:101-102, 195-196)."""
    name = type(sde).__name__
    for cls in type(sde).__mro__:
        if cls.__name__ in ("VPSDE", "VESDE", "subVPSDE"):
            name = cls.__name__
            break
    if name == "VPSDE":
        return "VP"
    if name == "VESDE":
        return "VE"
    if name == "subVPSDE":
        return "subVP"
    raise NotImplementedError(f"SDE class {type(sde).__name__} not supported.")
