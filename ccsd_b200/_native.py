"""ctypes binding of the C ABI declared in include/ccsd_b200.h.

The product library is ``ccsd_b200/_lib/libccsd_b200.so`` (built in-tree by ``ccsd_b200.build``
for sm_100a).  There is no CPU implementation: if the library is missing, or torch sees no CUDA
device, every entry point of the package raises.  ``CCSD_B200_LIB`` may point at another CUDA build of
the same ABI (A/B experiments).  The host-emulation build of the kernels (``tests/_emu``, test
infrastructure, see ccsd_b200/csrc/common.cuh) is refused unless the test-suite itself asked for it
with ``enable_test_emulation()`` -- no environment variable turns the package into a CPU path.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

MAX_LAYERS, MAX_CH, MAX_MLP, MAX_HODGE, MAX_F_LAYERS = 8, 8, 4, 2, 4
OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA, ERR_STATE = 0, -1, -2, -3, -4
SAMPLER_PC, SAMPLER_S4 = 0, 1
NET_X, NET_ADJ, NET_RANK2 = 0, 1, 2

i32, f32 = C.c_int32, C.c_float


class Mlp(C.Structure):
    _fields_ = [("nl", i32), ("din", i32), ("dhid", i32), ("dout", i32), ("w", i32 * MAX_MLP), ("b", i32 * MAX_MLP)]


class Gcn(C.Structure):
    _fields_ = [("din", i32), ("dout", i32), ("w", i32), ("b", i32)]


class AttnLayer(C.Structure):
    _fields_ = [
        ("c_in", i32), ("c_out", i32), ("conv_in", i32), ("attn_dim", i32), ("conv_out", i32),
        ("q", Gcn * MAX_CH), ("k", Gcn * MAX_CH), ("v", Gcn * MAX_CH), ("vw", Gcn * MAX_CH), ("conv_mlp", i32), ("qm", Mlp * MAX_CH), ("km", Mlp * MAX_CH),
        ("mlp", Mlp), ("multi_channel", Mlp),
    ]


class NetX(C.Structure):
    _fields_ = [("nfeat", i32), ("depth", i32), ("nhid", i32), ("fdim", i32), ("gcn", Gcn * MAX_LAYERS), ("fin", Mlp),
                ("gmh", i32), ("gmh_c_init", i32), ("gmh_heads", i32), ("glayer", AttnLayer * MAX_LAYERS)]


class HodgeLayer(C.Structure):
    _fields_ = [
        ("c_in", i32), ("c_out", i32), ("attn_dim", i32), ("proj_row", i32),
        ("bq", i32 * MAX_CH), ("bk", i32 * MAX_CH), ("mlp_attention", Mlp), ("mlp_value", Mlp),
    ]


class HBaseLayer(C.Structure):
    _fields_ = [
        ("c_in", i32), ("c_out", i32), ("hid", i32),
        ("w1", i32 * MAX_CH), ("b1", i32 * MAX_CH), ("w2", i32 * MAX_CH), ("b2", i32 * MAX_CH), ("mlp_hodge", Mlp),
    ]


class NetA(C.Structure):
    _fields_ = [
        ("is_cc", i32), ("num_layers", i32), ("c_init", i32), ("num_heads", i32), ("fdim", i32),
        ("layer", AttnLayer * MAX_LAYERS), ("num_layers_h", i32), ("num_heads_h", i32),
        ("n_proj_rows", i32 * MAX_HODGE), ("proj_w", i32), ("hodge", HodgeLayer * MAX_HODGE),
        ("base_cc", i32), ("hbase", HBaseLayer * MAX_HODGE), ("fin", Mlp),
    ]


class NetF(C.Structure):
    _fields_ = [
        ("num_layers", i32), ("cnum", i32), ("fdim", i32), ("use_hodge_mask", i32),
        ("layer", Mlp * MAX_F_LAYERS), ("fin", Mlp), ("affine", i32), ("aff", f32 * 3),
    ]


class ObjCoef(C.Structure):
    _fields_ = [(n, f32) for n in (
        "score_scale", "lg_alpha", "pa", "pb", "pc", "s4_alpha", "s4_m1", "s4_s1", "s4_sd", "s4_m2", "s4_s2", "pad_")]


class PlanDesc(C.Structure):
    _fields_ = [
        ("B", i32), ("N", i32), ("F", i32), ("is_cc", i32), ("E", i32), ("K", i32), ("d_min", i32), ("d_max", i32),
        ("sampler", i32), ("use_corrector", i32), ("n_lang_steps", i32), ("denoise", i32), ("n_diff_steps", i32), ("nets", i32),
        ("snr", f32), ("scale_eps", f32), ("netx", NetX), ("neta", NetA), ("netf", NetF),
    ]


_LIB = None
_TEST_EMULATION = None   # path of the host-emulation build, set only by enable_test_emulation()


def enable_test_emulation(path) -> None:
    """TEST-SUITE ONLY (tests/conftest.py, tests/test_shard_gloo.py workers): drive the host-emulation build
    of the kernels so that `pytest -m "not gpu"` can check index arithmetic and packing without a GPU."""
    global _TEST_EMULATION, _LIB
    _TEST_EMULATION = Path(path)
    _LIB = None


def lib_path() -> Path:
    if _TEST_EMULATION is not None:
        return _TEST_EMULATION
    env = os.environ.get("CCSD_B200_LIB")
    if env:
        return Path(env)
    return Path(__file__).resolve().parent / "_lib" / "libccsd_b200.so"


def load():
    """Load the shared library (once).  Raises RuntimeError when it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    p = lib_path()
    if not p.exists():
        raise RuntimeError(
            f"ccsd_b200: native library {p} not found. Build it with `python -m ccsd_b200.build` "
            "(nvcc, sm_100a). There is no CPU fallback."
        )
    lib = C.CDLL(str(p))
    vp, sz = C.c_void_p, C.c_size_t
    lib.ccsd_plan_desc_size.restype = C.c_int
    lib.ccsd_objcoef_size.restype = C.c_int
    lib.ccsd_plan_create.restype = C.c_int
    lib.ccsd_plan_create.argtypes = [C.POINTER(PlanDesc), vp, vp, sz, C.POINTER(vp)]
    lib.ccsd_plan_destroy.restype = None
    lib.ccsd_plan_destroy.argtypes = [vp]
    lib.ccsd_plan_workspace_bytes.restype = sz
    lib.ccsd_plan_workspace_bytes.argtypes = [vp]
    lib.ccsd_plan_bind.restype = C.c_int
    lib.ccsd_plan_bind.argtypes = [vp, vp, sz, vp]
    lib.ccsd_plan_set_traj.restype = C.c_int
    lib.ccsd_plan_set_traj.argtypes = [vp, vp, vp, vp]
    lib.ccsd_plan_init.restype = C.c_int
    lib.ccsd_plan_init.argtypes = [vp, vp, vp, vp, vp, C.c_uint64, C.c_int64, vp]
    lib.ccsd_plan_step.restype = C.c_int
    lib.ccsd_plan_step.argtypes = [vp, C.c_int, vp, vp, vp, vp]
    lib.ccsd_plan_run.restype = C.c_int
    lib.ccsd_plan_run.argtypes = [vp, C.c_int, C.c_int, vp]
    lib.ccsd_plan_read.restype = C.c_int
    lib.ccsd_plan_read.argtypes = [vp, C.c_int, vp, vp, vp, vp]
    lib.ccsd_score_eval.restype = C.c_int
    lib.ccsd_score_eval.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp, vp]
    lib.ccsd_mol_onehot.restype = C.c_int
    lib.ccsd_mol_onehot.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, vp]
    lib.ccsd_quantize.restype = C.c_int
    lib.ccsd_quantize.argtypes = [vp, vp, sz, C.c_float, C.c_int, vp]
    lib.ccsd_cc_cells.restype = C.c_int
    lib.ccsd_cc_cells.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, vp]
    lib.ccsd_plan_launch_count.restype = C.c_int64
    lib.ccsd_plan_launch_count.argtypes = [vp]
    lib.ccsd_plan_info.restype = C.c_int
    lib.ccsd_plan_info.argtypes = [vp, C.c_int]
    lib.ccsd_plan_set_profiling.restype = C.c_int
    lib.ccsd_plan_set_profiling.argtypes = [vp, C.c_int]
    lib.ccsd_plan_get_profile.restype = C.c_int
    lib.ccsd_plan_get_profile.argtypes = [vp, C.c_int, vp, C.c_int, vp]
    lib.ccsd_debug_gram.restype = C.c_int
    lib.ccsd_debug_gram.argtypes = [vp, vp, vp, vp, C.c_int, vp]
    lib.ccsd_debug_apply_trace.restype = C.c_int
    lib.ccsd_debug_apply_trace.argtypes = [vp, vp]
    lib.ccsd_last_error.restype = C.c_char_p
    lib.ccsd_version.restype = C.c_char_p
    if b"EMULATION" in lib.ccsd_version() and _TEST_EMULATION is None:
        raise RuntimeError(
            f"ccsd_b200: {p} is the host-emulation build of the kernels (test infrastructure); the package has no "
            "CPU path. Build the CUDA library with `python -m ccsd_b200.build`."
        )
    if lib.ccsd_plan_desc_size() != C.sizeof(PlanDesc) or lib.ccsd_objcoef_size() != C.sizeof(ObjCoef):
        raise RuntimeError("ccsd_b200: ABI mismatch between ccsd_b200/_native.py and the shared library")
    _LIB = lib
    return lib


def is_emulation() -> bool:
    return b"EMULATION" in load().ccsd_version()


def check(code: int) -> None:
    """Map C error codes to the reference's exception types."""
    if code == OK:
        return
    msg = load().ccsd_last_error().decode()
    if code == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    if code == ERR_INVALID:
        raise ValueError(msg)
    raise RuntimeError(msg)
