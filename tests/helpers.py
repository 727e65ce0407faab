"""Shared test helpers: golden fixtures -> oracle models + weight holders for the packer."""
from __future__ import annotations

import json
from pathlib import Path
from typing import Dict, List, Tuple

import numpy as np
import torch

from oracle import ccsd_oracle as O

GOLDEN = Path(__file__).resolve().parent / "golden"
STRIDE = 97


class Holder:
    """Duck-typed stand-in for a reference Module: hyper-parameter attributes + state_dict()."""

    def __init__(self, kind: str, hp: dict, sd: Dict[str, torch.Tensor]):
        self.__dict__.update(hp)
        self.model_type = kind
        self._sd = sd

    def state_dict(self):
        return self._sd

    def eval(self):
        return self


class Config:
    def __init__(self, name: str):
        z = np.load(GOLDEN / f"weights_{name}.npz")
        self.name = name
        self.meta = json.loads(bytes(z["meta"]).decode())
        self.is_cc = self.meta["is_cc"]
        self.keys = ["x", "adj"] + (["rank2"] if self.is_cc else [])
        d = self.meta["data"]
        self.N, self.F = d["max_node_num"], d["max_feat_num"]
        self.d_min, self.d_max = (d["d_min"], d["d_max"]) if self.is_cc else (None, None)
        self.E, self.K = O.rank2_dim(self.N, self.d_min, self.d_max) if self.is_cc else (0, 0)
        self.oracle_models, self.holders = [], []
        for k in self.keys:
            hp = self.meta["params"][k]
            sd = {kk[len(k) + 1:]: torch.from_numpy(z[kk]) for kk in z.files if kk.startswith(k + "/")}
            self.oracle_models.append(O.Model(hp["model_type"], hp, sd, is_cc=self.is_cc))
            self.holders.append(Holder(hp["model_type"], hp, sd))
        self.shipped = self.meta["shipped_sampler"]

    def sdes(self):
        s = self.meta["sde"]
        return [O.make_sde(s[k]["type"], s[k]["beta_min"], s[k]["beta_max"], s[k]["num_scales"]) for k in self.keys]

    def shapes(self, B: int):
        sh = [(B, self.N, self.F), (B, self.N, self.N)]
        if self.is_cc:
            sh.append((B, self.E, self.K))
        return sh

    def random_state(self, B: int, seed: int = 1, r2_scale: float = 0.3):
        g = torch.Generator().manual_seed(seed)
        n = torch.randint(max(2, self.N // 2), self.N + 1, (B,), generator=g)
        n[0] = self.N
        flags = (torch.arange(self.N)[None, :] < n[:, None]).to(torch.float32)
        x = O.mask_x(torch.randn(B, self.N, self.F, generator=g), flags)
        adj = O.mask_adjs(O.symmetrize_noise(torch.randn(B, self.N, self.N, generator=g)), flags)
        r2 = None
        if self.is_cc:
            r2 = O.mask_rank2(torch.randn(B, self.E, self.K, generator=g) * r2_scale, self.N, self.d_min, self.d_max, flags)
        return x, adj, r2, flags

    def io(self):
        return np.load(GOLDEN / f"io_{self.name}.npz")


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / max |b|  -- the relative error used for the 1e-4 score-parity bar."""
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


def check_compressed(store, key: str, t: torch.Tensor, tol: float) -> float:
    """Compare a tensor with a golden entry written by make_golden.compress()."""
    t = t.detach().cpu().to(torch.float32)
    if f"{key}/full" in store.files:
        ref = torch.from_numpy(store[f"{key}/full"])
        return rel_err(t, ref)
    ref = torch.from_numpy(store[f"{key}/sample"])
    got = t.reshape(-1)[::STRIDE]
    e1 = rel_err(got, ref)
    s_ref, a_ref = float(store[f"{key}/sum"]), float(store[f"{key}/abssum"])
    e2 = abs(t.double().sum().item() - s_ref) / (a_ref + 1e-30)
    return max(e1, e2)
