"""
In-tree build of libmfk_b200.so (the C-ABI CUDA core) with nvcc for sm_100a.

    python -m matrix_factorization_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU.  The .so is written next to the sources
(matrix_factorization_b200/csrc/libmfk_b200.so) so it travels with the tree; it is
git-ignored.  Nothing here falls back to another backend: if nvcc is missing the build fails.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
LIB = os.path.join(CSRC, "libmfk_b200.so")
SOURCES = ["mfk_plan.cu", "mfk_sgd.cu", "mfk_eval.cu", "mfk_als.cu", "mfk_score.cu", "mfk_score_tc.cu", "mfk_host.cu", "mfk_prep.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ARCH + ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
                     "--expt-relaxed-constexpr", "-I", INCLUDE]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libmfk_b200.so cannot be built (no fallback backend exists)")


def _deps():
    out = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h", ".inc"))]
    out.append(os.path.join(INCLUDE, "mfk.h"))
    out.append(os.path.abspath(__file__))
    return out


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in _deps())


def build(force: bool = False, verbose: bool = False, ptxas_info: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    nvcc = _nvcc()
    objdir = os.path.join(CSRC, "build")
    os.makedirs(objdir, exist_ok=True)
    headers_t = max(os.path.getmtime(d) for d in _deps() if not d.endswith(".cu"))

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        sp = os.path.join(CSRC, src)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(sp), headers_t):
            return obj, ""
        extra = ["-DMFK_RING_PROFILE=1", "-DMFK_BATCH_PROFILE=1"] if os.environ.get("MFK_RING_PROFILE") == "1" else []
        extra += ["-D" + d for d in os.environ.get("MFK_NVCC_DEFS", "").split(",") if d]  # experiments
        cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if ptxas_info else []) + ["-c", sp, "-o", obj]
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 2)) as ex:
        results = list(ex.map(compile_one, SOURCES))
    if ptxas_info:
        for _, log in results:
            print(log)
    cmd = [nvcc] + ARCH + ["-shared", "-o", LIB] + [o for o, _ in results] + ["-cudart", "static"]
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, ptxas_info="--ptxas" in sys.argv))
