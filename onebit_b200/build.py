"""Builds libonebit_b200.so in-tree with nvcc for sm_100a (no torch headers involved).

    python -m onebit_b200.build [--force] [--verbose]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot; nothing is JIT-compiled
at import time on the box unless the library is missing or older than its sources.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
SUFFIX = os.environ.get("ONEBIT_LIB_SUFFIX", "")          # e.g. "_trace" for an instrumented side build
EXTRA = os.environ.get("ONEBIT_NVCC_EXTRA", "").split()   # e.g. -DONEBIT_TRACE
LIB = PKG / f"libonebit_b200{SUFFIX}.so"
OBJ = PKG / f"build{SUFFIX}"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "--use_fast_math", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
         "-ccbin", "/usr/bin/g++"]
# --use_fast_math only affects intrinsics we do not use on the numerics path (no div/exp there);
# it is dropped for files listed here to keep IEEE division / sqrt in the normalisation kernels.
PRECISE = {"layernorm.cu", "decoder.cu", "persist_step.cu"}


def sources():
    return sorted(CSRC.glob("*.cu"))


def _stale() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "onebit_b200.h"]
    return any(d.stat().st_mtime > t for d in deps)


def _compile(src: Path, verbose: bool) -> Path:
    obj = OBJ / (src.stem + ".o")
    flags = [f for f in FLAGS if not (src.name in PRECISE and f == "--use_fast_math")]
    cmd = [NVCC, *ARCH, *flags, *EXTRA, "-c", str(src), "-o", str(obj)]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
    if verbose and r.stderr:
        print(r.stderr)
    return obj


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not _stale():
        return LIB
    OBJ.mkdir(exist_ok=True)
    deps_t = max(p.stat().st_mtime for p in list(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "onebit_b200.h"])
    todo = []
    for s in sources():
        o = OBJ / (s.stem + ".o")
        if force or not o.exists() or o.stat().st_mtime < max(s.stat().st_mtime, deps_t):
            todo.append(s)
    with ThreadPoolExecutor(max_workers=min(8, max(1, len(todo)))) as ex:
        list(ex.map(lambda s: _compile(s, verbose), todo))
    objs = [str(OBJ / (s.stem + ".o")) for s in sources()]
    cmd = [NVCC, *ARCH, "-shared", "-ccbin", "/usr/bin/g++", "-o", str(LIB), *objs, "-lcudart_static", "-lcuda"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


TORCH_LIB = PKG / "libonebit_b200_torch.so"
TORCH_SRC = PKG / "csrc_torch" / "torch_ops.cpp"


def build_torch_ops(force: bool = False) -> Path:
    """The torch dispatcher shim (TORCH_LIBRARY onebit_b200: bitlinear, bitlinear_nolayernorm, pack_signs,
    unpack_signs) over the C ABI. Host C++ only (g++ against the installed torch headers); links libonebit_b200.so."""
    build(force=False)
    deps = [TORCH_SRC, PKG.parent / "include" / "onebit_b200.h"]
    if not force and TORCH_LIB.exists() and all(d.stat().st_mtime <= TORCH_LIB.stat().st_mtime for d in deps):
        return TORCH_LIB
    import torch
    tdir = Path(torch.__file__).resolve().parent
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    cmd = ["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", f"-D_GLIBCXX_USE_CXX11_ABI={abi}",
           f"-I{tdir / 'include'}", f"-I{tdir / 'include' / 'torch' / 'csrc' / 'api' / 'include'}",
           "-I/usr/local/cuda/include", f"-I{PKG.parent / 'include'}", str(TORCH_SRC), "-o", str(TORCH_LIB),
           f"-L{tdir / 'lib'}", "-ltorch", "-ltorch_cpu", "-lc10", "-lc10_cuda", "-ltorch_cuda",
           f"-L{PKG}", f"-l:{LIB.name}", "-Wl,-rpath,$ORIGIN", f"-Wl,-rpath,{tdir / 'lib'}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"g++ failed for {TORCH_SRC.name}:\n{r.stdout}\n{r.stderr}")
    return TORCH_LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(p)
    print(build_torch_ops(force="--force" in sys.argv))
