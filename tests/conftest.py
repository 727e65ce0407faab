"""pytest configuration.

* ``-m "not gpu"``: runs here without a GPU.  Kernel code is exercised through the host-emulation
  build (tests/_emu/libccsd_b200_emu.so, one thread per block -- see ccsd_b200/csrc/common.cuh);
  it is built on demand and selected via CCSD_B200_LIB *only* when torch sees no CUDA device.
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
if not HAS_CUDA and "CCSD_B200_LIB" not in os.environ:
    from ccsd_b200 import build as _build

    os.environ["CCSD_B200_LIB"] = str(_build.build_emu())


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def device():
    return "cuda" if HAS_CUDA else "cpu"
