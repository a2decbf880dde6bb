"""Build libcolpali_b200.so (sm_100a only) in-tree with nvcc.  `python -m colpali_b200.build`."""

from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libcolpali_b200.so")
SOURCES = ["cabi.cu", "maxsim_sm100.cu", "maxsim_kpipe_sm100.cu", "maxsim_pair_sm100.cu", "loss_sm100.cu", "smooth_bwd_sm100.cu", "exchange_sm100.cu", "head_sm100.cu",
           "head_wide_sm100.cu", "dense_sm100.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",  # explicit -gencode: the -arch shorthand emits compute_100 PTX
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]
LINK_FLAGS = ["-shared", "-cudart", "static"]
OBJ_DIR = os.path.join(HERE, "build")  # per-source objects (git-ignored): only what changed is recompiled


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build the sm_100a extension")


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "colpali_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _stale(obj: str, src: str, headers: list) -> bool:
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    return any(os.path.getmtime(d) > t for d in [src, *headers])


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every source to build/<name>.o (in parallel, only the stale ones) and link the shared library."""
    if not force and not needs_build():
        return OUT
    from concurrent.futures import ThreadPoolExecutor

    os.makedirs(OBJ_DIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(HERE, "..", "include", "colpali_b200.h"))
    nvcc = _nvcc()

    def compile_one(name: str):
        src, obj = os.path.join(CSRC, name), os.path.join(OBJ_DIR, name[:-3] + ".o")
        if force or _stale(obj, src, headers):
            cmd = [nvcc, *NVCC_FLAGS, *(["-Xptxas", "-v"] if verbose else []), "-c", "-o", obj, src]
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                raise RuntimeError(f"nvcc failed on {name} ({res.returncode}):\n{res.stdout}\n{res.stderr}")
            if verbose:
                print(" ".join(cmd), res.stderr, sep="\n", flush=True)
        return obj

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as pool:
        objs = list(pool.map(compile_one, SOURCES))
    res = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", *LINK_FLAGS, "-o", OUT, *objs],
                         capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed ({res.returncode}):\n{res.stdout}\n{res.stderr}")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
