"""ccsd_b200 -- B200-native reverse-SDE sampler for CCSD (drop-in for ccsd/src/solver.py).

Public surface mirrors the reference: ``get_pc_sampler`` / ``S4_solver`` (ccsd/src/solver.py:856,
1179), SDE classes (ccsd/src/sde.py), ``load_sde`` / ``load_sampling_fn`` (ccsd/src/utils/loader.py:
242, 337).  The compute path is the CUDA extension in ``ccsd_b200/_lib``; there is no CPU fallback.
"""
from .sde import VPSDE, VESDE, subVPSDE  # noqa: F401
from .solver import Engine, InjectedNoise, S4_solver, get_pc_sampler, mol_onehot, quantize  # noqa: F401

__version__ = "0.1.0"
