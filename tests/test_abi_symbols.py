"""The C-ABI library loads without a GPU and exports every function include/ccsd_b200.h declares."""
import ctypes
import re
from pathlib import Path

from ccsd_b200 import build

ROOT = Path(__file__).resolve().parents[1]


def _declared():
    text = (ROOT / "include" / "ccsd_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ccsd_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    names = _declared()
    for must in ("ccsd_plan_create", "ccsd_plan_bind", "ccsd_plan_init", "ccsd_plan_step", "ccsd_plan_run", "ccsd_plan_read",
                 "ccsd_score_eval", "ccsd_quantize", "ccsd_mol_onehot", "ccsd_last_error"):
        assert must in names


def test_product_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(str(build.build_cuda()))   # nvcc cross-compiles here; loading needs no GPU
    for name in _declared():
        assert hasattr(lib, name), name
    lib.ccsd_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.ccsd_version()


def test_emulation_library_exports_the_same_surface():
    lib = ctypes.CDLL(str(build.build_emu()))
    for name in _declared():
        assert hasattr(lib, name), name


def test_emulation_build_is_refused_outside_the_test_hook(monkeypatch):
    """No environment variable turns the package into a CPU path: pointing CCSD_B200_LIB at the emulation
    build raises unless the test-suite's own hook enabled it."""
    import pytest
    from ccsd_b200 import _native as nat

    saved = (nat._TEST_EMULATION, nat._LIB)
    try:
        nat._TEST_EMULATION, nat._LIB = None, None
        monkeypatch.setenv("CCSD_B200_LIB", str(build.build_emu()))
        with pytest.raises(RuntimeError, match="host-emulation"):
            nat.load()
    finally:
        nat._TEST_EMULATION, nat._LIB = saved
