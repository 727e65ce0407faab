"""Build recipes (in-tree, explicit nvcc): ``python -m ccsd_b200.build [--emu] [--force]``.

  product:   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo  ->  ccsd_b200/_lib/libccsd_b200.so
  --emu:     g++ -x c++ -DCCSD_EMU (host emulation of the kernels, CPU tests only)
             ->  tests/_emu/libccsd_b200_emu.so
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "ccsd_b200" / "csrc"
LIB = ROOT / "ccsd_b200" / "_lib" / "libccsd_b200.so"
EMU = ROOT / "tests" / "_emu" / "libccsd_b200_emu.so"


def _sources():
    return sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) +
                  [ROOT / "include" / "ccsd_b200.h"])


def _stale(target: Path) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(s.stat().st_mtime > t for s in _sources())


def build_cuda(force: bool = False, verbose: bool = False) -> Path:
    if not force and not _stale(LIB):
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    LIB.parent.mkdir(parents=True, exist_ok=True)
    cmd = [
        nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
        "-Xcompiler", "-fPIC", "-shared", "-o", str(LIB), str(CSRC / "ccsd_b200.cu"),
    ]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return LIB


def build_emu(force: bool = False) -> Path:
    if not force and not _stale(EMU):
        return EMU
    EMU.parent.mkdir(parents=True, exist_ok=True)
    cmd = [
        "g++", "-x", "c++", "-DCCSD_EMU", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", str(EMU),
        str(CSRC / "ccsd_b200.cu"),
    ]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ (emulation build) failed:\n" + r.stdout + r.stderr)
    return EMU


if __name__ == "__main__":
    force = "--force" in sys.argv
    if "--emu" in sys.argv:
        print(build_emu(force))
    else:
        print(build_cuda(force, verbose="-v" in sys.argv))
