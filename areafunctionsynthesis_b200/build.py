"""Build libafs_b200.so (CUDA kernels + C ABI) in-tree for sm_100a.

    python -m areafunctionsynthesis_b200.build [--force]

nvcc cross-compiles without a GPU.  The TL translation unit is built with
-fmad=false: its results are bit-identical to the reference's g++ build
(no FMA contraction); see csrc/tl_kernels.cu.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# (profiling builds: AFS_TDS_TIMING=1 AFS_LIB_TAG=_timing python -m areafunctionsynthesis_b200.build --force writes
#  libafs_b200_timing.so beside the product library; tools/tds_timing.py loads it through AFS_LIB)
TAG = os.environ.get("AFS_LIB_TAG", "")
LIB = os.path.join(HERE, f"libafs_b200{TAG}.so")
OBJ = os.path.join(HERE, "build" + TAG)
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
EXTRA_DEFS = (["-DAFS_TDS_TIMING"] if os.environ.get("AFS_TDS_TIMING") else []) + os.environ.get("AFS_EXTRA_DEFS", "").split()
COMMON = [*EXTRA_DEFS, "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off"]
# per-file device flags
UNITS = {
    "tl_kernels.cu": ["-fmad=false"],
    "tl_fast_kernels.cu": [],
    "tds_kernels.cu": [],
    "probe_kernels.cu": [],
    "raster_kernels.cu": ["-fmad=false"],
    "afs_api.cu": [],
    "afs_host.cpp": [],
    "afs_multi.cpp": [],
}


def _sources_mtime():
    m = os.path.getmtime(os.path.abspath(__file__))
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in os.listdir(root):
            m = max(m, os.path.getmtime(os.path.join(root, f)))
    return m


def build_library(force=False, verbose=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _sources_mtime():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    objs = []
    procs = []
    for src, extra in UNITS.items():
        obj = os.path.join(OBJ, src.rsplit(".", 1)[0] + ".o")
        cmd = [NVCC, *ARCH, *COMMON, *extra, "-Xptxas", "-v", "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    log = []
    for src, cmd, p in procs:
        out, _ = p.communicate()
        log.append(f"$ {' '.join(cmd)}\n{out}")
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
    cmd = [NVCC, *ARCH, "-shared", "-o", LIB, *objs]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log.append(f"$ {' '.join(cmd)}\n{r.stdout}")
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    with open(os.path.join(OBJ, "build.log"), "w") as fh:
        fh.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
