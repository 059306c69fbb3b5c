"""Build libpadel_b200.so (hand-written sm_100a CUDA + C ABI) in-tree with nvcc.

The library is torch-free: plain `nvcc -gencode arch=compute_100a,code=sm_100a` on csrc/*.cu, a few seconds per TU.
The built .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB = HERE / "libpadel_b200.so"
OBJ = HERE / "build"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
]
# Bring-up build (`python -m padel_analytics_b200.build --debug`): the same sources + csrc/debug/*.cu with
# -DPB_DEBUG_BUILD -> libpadel_b200_debug.so, which adds the pb_debug_* hooks the scripts/exp_*.py experiments use.
# It never overwrites the product library (whose exported ABI is exactly include/padel_b200.h); load it with
# PADEL_B200_LIB=padel_analytics_b200/libpadel_b200_debug.so.
DEBUG_BUILD = "--debug" in sys.argv or os.environ.get("PADEL_B200_DEBUG_BUILD") == "1"
if DEBUG_BUILD:
    LIB = HERE / "libpadel_b200_debug.so"
    OBJ = HERE / "build_debug"
    NVCC_FLAGS.append("-DPB_DEBUG_BUILD")


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _digest(paths) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def sources():
    return sorted(CSRC.glob("*.cu")) + (sorted((CSRC / "debug").glob("*.cu")) if DEBUG_BUILD else [])


def build_lib(force: bool = False, verbose: bool = False) -> Path:
    srcs = sources()
    deps = srcs + sorted(CSRC.glob("*.h")) + sorted(CSRC.glob("*.cuh")) + [HERE.parent / "include" / "padel_b200.h",
                                                                          CSRC / "exports.map"]
    stamp = OBJ / "stamp.txt"
    dig = _digest(deps)
    if not force and LIB.exists() and stamp.exists() and stamp.read_text() == dig:
        return LIB
    OBJ.mkdir(exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src: Path) -> Path:
        obj = OBJ / (src.stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr, file=sys.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    cmd = [nvcc, "-shared", "-o", str(LIB), *map(str, objs), "-cudart", "static",
           "-Xlinker", f"--version-script={CSRC / 'exports.map'}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    stamp.write_text(dig)
    return LIB


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv))
