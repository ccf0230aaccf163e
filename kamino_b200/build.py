"""Build libkamino_b200.so (hand-written sm_100a kernels + C ABI) with nvcc, in-tree.

    python -m kamino_b200.build            # build if stale
    python -m kamino_b200.build --force

The library has no CPU fallback; it is cross-compiled here (no GPU needed) and travels to the GPU
box with the repository snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
INCLUDE = PKG.parent / "include"
LIB_DIR = PKG / "_lib"
LIB_PATH = LIB_DIR / "libkamino_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "--extended-lambda", "-Xcompiler", "-fPIC", "-shared",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found; libkamino_b200.so cannot be built (there is no CPU fallback)")


def sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _stale() -> bool:
    if not LIB_PATH.exists():
        return True
    t = LIB_PATH.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(INCLUDE.glob("*.h"))
    return any(p.stat().st_mtime > t for p in deps)


def build_variant(name: str, defines: list[str]) -> Path:
    """Tuning builds (tools/ only): the same sources with -D overrides, as _lib/libkamino_b200_<name>.so
    (loaded through the KAMINO_B200_LIB environment variable)."""
    LIB_DIR.mkdir(exist_ok=True)
    out = LIB_DIR / f"libkamino_b200_{name}.so"
    cmd = [_nvcc(), *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-I", str(INCLUDE), "-o", str(out), *map(str, sources())]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError(f"nvcc failed building {out.name}")
    return out


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not _stale():
        return LIB_PATH
    LIB_DIR.mkdir(exist_ok=True)
    cmd = [_nvcc(), *NVCC_FLAGS, "-I", str(INCLUDE), "-o", str(LIB_PATH), *map(str, sources())]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    (LIB_DIR / "build.log").write_text(" ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError("nvcc failed building libkamino_b200.so")
    if verbose:
        sys.stderr.write(proc.stderr)
    return LIB_PATH


HOST_BIN = LIB_DIR / "kamino-b200"
HOST_SRC = CSRC / "host"


def build_host(force: bool = False) -> Path:
    """`kamino-b200`: the C++ host pipeline (kamino's CLI) linked against libkamino_b200.so."""
    lib = build(force=False)
    srcs = sorted(HOST_SRC.glob("*.cpp"))
    deps = srcs + list(HOST_SRC.glob("*.hpp")) + [lib]
    if not force and HOST_BIN.exists() and all(p.stat().st_mtime <= HOST_BIN.stat().st_mtime for p in deps):
        return HOST_BIN
    cxx = os.environ.get("CXX") or shutil.which("g++") or "g++"
    cmd = [cxx, "-O2", "-std=c++17", "-ffp-contract=off", "-Wall", "-Wextra", "-pthread", "-I", str(INCLUDE),
           "-o", str(HOST_BIN), *map(str, srcs), "-L", str(LIB_DIR), "-lkamino_b200", "-lz",
           "-Wl,-rpath,$ORIGIN"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError("g++ failed building kamino-b200")
    if proc.stderr.strip():
        sys.stderr.write(proc.stderr)
    return HOST_BIN


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
    print(build_host(force="--force" in sys.argv))
