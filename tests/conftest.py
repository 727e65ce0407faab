"""pytest configuration.

* ``-m "not gpu"``: runs here without a GPU.  Kernel code is exercised through the host-emulation
  build (tests/_emu/libccsd_b200_emu.so, one thread per block -- see ccsd_b200/csrc/common.cuh);
  it is built on demand and selected through ccsd_b200._native.enable_test_emulation() *only* when torch
  sees no CUDA device (no environment variable makes the product package load it).
* ``-m gpu``: parity tests proper, on a B200, through the product library and the C ABI.
"""
import os
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

HAS_CUDA = torch.cuda.is_available()
if not HAS_CUDA:
    from ccsd_b200 import _native as _nat
    from ccsd_b200 import build as _build

    _EMU = str(_build.build_emu())
    os.environ["CCSD_B200_TEST_EMU"] = _EMU     # for spawned worker processes of the test-suite (they call the hook too)
    _nat.enable_test_emulation(_EMU)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def device():
    return "cuda" if HAS_CUDA else "cpu"
