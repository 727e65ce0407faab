"""The factory seam of the reference (ccsd/src/utils/loader.py): ``load_sde`` (:242-275),
``load_sampling_fn`` (:337-458), ``load_model`` (:70-100), ``load_model_from_ckpt`` (:619-653) and
``load_seed`` (:35-55), with the same signatures and config attribute names -- so that replacing

    from ccsd.src.utils.loader import load_sampling_fn

by the import from this module routes ``Sampler_*.sample()`` (ccsd/src/sampler.py:138,415,727,1104)
onto the CUDA path.  Configs may be EasyDicts or any object/dict with the same fields.
"""
from __future__ import annotations

import random
from typing import Any, Callable, List, Optional, Union

import numpy as np
import torch

from .models import load_model  # noqa: F401  (re-export, loader.py:70)
from .packer import rank2_dim
from .sde import VESDE, VPSDE, subVPSDE
from .solver import S4_solver, get_pc_sampler


def _get(cfg: Any, name: str):
    return cfg[name] if isinstance(cfg, dict) else getattr(cfg, name)


def load_seed(seed: int) -> int:
    """loader.py:35-55."""
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    return seed


def load_sde(config_sde: Any):
    """loader.py:242-275 (VE is built with sigma_min=beta_min, sigma_max=beta_max)."""
    sde_type = _get(config_sde, "type")
    beta_min, beta_max, num_scales = _get(config_sde, "beta_min"), _get(config_sde, "beta_max"), _get(config_sde, "num_scales")
    if sde_type == "VP":
        return VPSDE(beta_min=beta_min, beta_max=beta_max, N=num_scales)
    if sde_type == "VE":
        return VESDE(sigma_min=beta_min, sigma_max=beta_max, N=num_scales)
    if sde_type == "subVP":
        return subVPSDE(beta_min=beta_min, beta_max=beta_max, N=num_scales)
    raise NotImplementedError(f"SDE class {sde_type} not (yet) supported.")


def load_sampling_fn(
    config_train: Any, config_module: Any, config_sample: Any, device: Union[str, List[str], List[int]],
    is_cc: bool = False, d_min: Optional[int] = None, d_max: Optional[int] = None, divide_batch: Optional[int] = None,
) -> Callable:
    """loader.py:337-458: same shape rules (n_samples for QM9/ZINC250k, data.batch_size otherwise,
    // divide_batch), same keyword hand-off to ``S4_solver`` / ``get_pc_sampler``."""
    sde_cfg, data = _get(config_train, "sde"), _get(config_train, "data")
    sde_x, sde_adj = load_sde(_get(sde_cfg, "x")), load_sde(_get(sde_cfg, "adj"))
    sde_rank2 = load_sde(_get(sde_cfg, "rank2")) if is_cc else None
    max_node_num = _get(data, "max_node_num")
    device_id = f"cuda:{device[0]}" if isinstance(device, list) else device
    get_sampler = S4_solver if _get(config_module, "predictor") == "S4" else get_pc_sampler
    if _get(data, "data") in ["QM9", "ZINC250k"]:
        total = _get(config_sample, "n_samples")
    else:
        total = _get(data, "batch_size")
    batch_size = total if divide_batch is None else total // divide_batch
    shape_x = (batch_size, max_node_num, _get(data, "max_feat_num"))
    shape_adj = (batch_size, max_node_num, max_node_num)
    kw = dict(
        sde_x=sde_x, sde_adj=sde_adj, shape_x=shape_x, shape_adj=shape_adj,
        predictor=_get(config_module, "predictor"), corrector=_get(config_module, "corrector"),
        snr=_get(config_module, "snr"), scale_eps=_get(config_module, "scale_eps"), n_steps=_get(config_module, "n_steps"),
        probability_flow=_get(config_sample, "probability_flow"), continuous=True,
        denoise=_get(config_sample, "noise_removal"), eps=_get(config_sample, "eps"), device=device_id,
    )
    if is_cc:
        E, K = rank2_dim(max_node_num, d_min, d_max)
        kw.update(is_cc=True, sde_rank2=sde_rank2, shape_rank2=(batch_size, E, K), d_min=d_min, d_max=d_max)
    return get_sampler(**kw)


def load_model_from_ckpt(params: dict, state_dict: dict, device: Union[str, List[str], List[int]]) -> torch.nn.Module:
    """loader.py:619-653: build, strip DataParallel's ``module.`` prefix, load.  No DataParallel wrap:
    multi-GPU sampling shards the batch across processes instead (ccsd_b200/shard.py)."""
    model = load_model(params)
    if "module." in list(state_dict.keys())[0]:
        state_dict = {k[7:]: v for k, v in state_dict.items()}
    model.load_state_dict(state_dict)
    dev = f"cuda:{device[0]}" if isinstance(device, list) else device
    return model.to(dev)
