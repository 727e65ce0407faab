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


def _nvcc_base():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    return [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]


# translation units: (object name, source, extra flags, headers it depends on besides its own source)
_APPLY_DEPS = ["tc_apply.cuh", "r2_kernels.cuh", "prims.cuh", "plan_dev.h", "common.cuh", "tc_common.cuh"]


def _units():
    units = [("ccsd_b200", "ccsd_b200.cu", [], None)]   # None: depends on every header
    for k in range(5):
        units.append((f"tc_apply_f{k}", "tc_apply_tu.cu", [f"-DTA_FMODE={k}"], _APPLY_DEPS))
    units.append(("tc_r2big", "tc_r2big_tu.cu", [], ["tc_r2big.cuh", "r2_kernels.cuh", "prims.cuh", "plan_dev.h", "common.cuh", "tc_common.cuh"]))
    units.append(("tc_hnorm", "tc_hnorm_tu.cu", [], ["tc_hnorm.cuh", "r2_kernels.cuh", "prims.cuh", "plan_dev.h", "common.cuh", "tc_common.cuh"]))
    units.append(("tc_edge", "tc_edge_tu.cu", [], ["tc_edge.cuh", "xa_pipe.cuh", "r2_kernels.cuh", "prims.cuh", "plan_dev.h", "common.cuh", "tc_common.cuh"]))
    units.append(("tc_xfin", "tc_xfin_tu.cu", [], ["tc_xfin.cuh", "xa_pipe.cuh", "r2_kernels.cuh", "prims.cuh", "plan_dev.h", "common.cuh", "tc_common.cuh"]))
    units.append(("tc_attn", "tc_attn_tu.cu", [], ["tc_attn.cuh", "xa_pipe.cuh", "r2_kernels.cuh", "prims.cuh", "plan_dev.h", "common.cuh", "tc_common.cuh"]))
    return units


def build_cuda(force: bool = False, verbose: bool = False) -> Path:
    """Compile the translation units in parallel (objects cached under build/obj by source mtime) and link."""
    if not force and not _stale(LIB):
        return LIB
    from concurrent.futures import ThreadPoolExecutor

    LIB.parent.mkdir(parents=True, exist_ok=True)
    objdir = ROOT / "build" / "obj"
    objdir.mkdir(parents=True, exist_ok=True)
    hdr_all = [s for s in _sources() if s.suffix != ".cu"]
    jobs = []
    for name, src, flags, deps in _units():
        obj = objdir / f"{name}.o"
        dep_files = [CSRC / src] + (hdr_all if deps is None else [CSRC / d for d in deps] + [ROOT / "include" / "ccsd_b200.h"])
        fresh = obj.exists() and all(f.stat().st_mtime <= obj.stat().st_mtime for f in dep_files)
        if force or not fresh:
            cmd = _nvcc_base() + flags + (["-Xptxas=-v"] if verbose else []) + ["-c", str(CSRC / src), "-o", str(obj)]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        return r.stderr

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        for log in ex.map(run, jobs):
            if verbose:
                print(log)
    objs = [str(objdir / f"{name}.o") for name, _, _, _ in _units()]
    r = subprocess.run(_nvcc_base() + ["-shared", "-o", str(LIB)] + objs, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
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
