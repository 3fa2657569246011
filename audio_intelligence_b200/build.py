"""Build liba2sb_b200.so (hand-written sm_100a CUDA kernels + C-ABI) in-tree with nvcc.

The heavy kernel families (one per n_fft and direction) are separate translation units so they
compile in parallel; nvcc cross-compiles without a GPU.  The result is placed next to this file
(git-ignored, but it travels to the GPU box with the repo snapshot).
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "liba2sb_b200.so")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v",
]
N_INST = 8


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the B200 path cannot be built (there is no CPU fallback)")
    return exe


def _sources_mtime() -> float:
    m = 0.0
    for d in (CSRC, os.path.join(HERE, "..", "include")):
        for f in os.listdir(d):
            m = max(m, os.path.getmtime(os.path.join(d, f)))
    return max(m, os.path.getmtime(__file__))


def _compile(args):
    src, obj, extra = args
    cmd = [_nvcc(), *NVCC_FLAGS, *extra, "-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    with open(obj + ".log", "w") as fh:
        fh.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj


def build(force: bool = False, verbose: bool = False, defines: tuple = (), suffix: str = "") -> str:
    """Compile (if stale) and return the path of liba2sb_b200.so.  `defines`/`suffix` build an
    experiment variant (e.g. -DA2SB_PLAIN_STORES -> liba2sb_b200_plain.so) next to the product library."""
    lib = LIB.replace(".so", f"{suffix}.so")
    obj = OBJ + suffix
    if not force and os.path.exists(lib) and os.path.getmtime(lib) >= _sources_mtime():
        return lib
    os.makedirs(obj, exist_ok=True)
    extra = [f"-D{d}" for d in defines]
    jobs = [(os.path.join(CSRC, "a2sb_api.cu"), os.path.join(obj, "api.o"), extra)]
    for k in range(1, N_INST + 1):
        jobs.append((os.path.join(CSRC, "inst.cu"), os.path.join(obj, f"inst{k}.o"), [f"-DA2SB_INST={k}", *extra]))
    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 4)) as ex:
        objs = list(ex.map(_compile, jobs))
    cmd = [_nvcc(), "-shared", "-o", lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if verbose:
        for o in objs:
            sys.stdout.write(open(o + ".log").read())
    return lib


if __name__ == "__main__":
    _defs = tuple(a[2:] for a in sys.argv if a.startswith("-D"))
    _suf = next((a.split("=", 1)[1] for a in sys.argv if a.startswith("--suffix=")), "")
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, defines=_defs, suffix=_suf))
