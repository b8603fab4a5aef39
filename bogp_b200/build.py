"""Builds libbogp_sm100.so (hand-written sm_100a CUDA + the C ABI of include/bogp.h) in-tree with nvcc.

No torch.utils.cpp_extension: the library has a plain C ABI (no torch types), so it is a straight
``nvcc -shared``.  The .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libbogp_sm100.so"
STAMP = PKG / ".libbogp_sm100.stamp"
SOURCES = ["hostmath.cu", "api_mfp.cu", "kernels_simt.cu", "kernels_umma.cu", "kernels_umma16.cu", "kernels_env.cu", "kernels_gp.cu", "kernels_gpfit.cu", "kernels_gpfit_coop.cu", "kernels_ts.cu", "kernels_peer.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--use_fast_math=false"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _digest() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h"))
                    + [PKG.parent / "include" / "bogp.h", Path(__file__)]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every CUDA source for sm_100a and link the shared library.  Returns its path."""
    digest = _digest() + ("+dbg" if os.environ.get("BOGP_UMMA_DBG_BUILD") == "1" else "") + ("+coopdbg" if os.environ.get("BOGP_COOP_DBG_BUILD") == "1" else "")
    if not force and LIB.exists() and STAMP.exists() and STAMP.read_text().strip() == digest:
        return LIB
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")]
    if os.environ.get("BOGP_UMMA_DBG_BUILD") == "1":      # per-role wait counters in the tcgen05 kernel (tools/umma_dbg_counters.py)
        flags.append("-DBOGP_UMMA_DBG=1")
    if os.environ.get("BOGP_COOP_DBG_BUILD") == "1":      # per-phase cycle counts of the cooperative GP fit (one step, CTA 0)
        flags.append("-DBOGP_COOP_DBG=1")
    objdir = PKG / "build"
    objdir.mkdir(exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = objdir / (src + ".o")
        cmd = [_nvcc(), *flags, "-c", str(CSRC / src), "-o", str(obj)]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = []
    for src, obj, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        if verbose and out.strip():
            print(out, file=sys.stderr)
        objs.append(str(obj))
    cmd = [_nvcc(), "-shared", "-o", str(LIB), *objs, "-lcudart"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    STAMP.write_text(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
