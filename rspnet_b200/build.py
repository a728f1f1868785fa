"""In-tree build of the sm_100a shared library (C ABI declared in include/rspnet_b200.h).

nvcc cross-compiles without a GPU; the resulting ``rspnet_b200/lib/librspnet_b200.so`` is git-ignored but
travels with the working tree to the GPU box.
"""
import hashlib
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent
REPO = ROOT.parent
CSRC = ROOT / "csrc"
LIB_DIR = ROOT / "lib"
LIB_PATH = LIB_DIR / "librspnet_b200.so"
SOURCES = ["common.cu", "conv_igemm.cu", "conv_stem.cu", "conv_stem3.cu", "conv_direct.cu", "conv_wgrad_direct.cu", "elementwise.cu", "bn_pool.cu", "bn_pool_bwd.cu", "moco.cu", "peer.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--use_fast_math=false",
]


def _nvcc() -> str:
    cand = os.environ.get("NVCC") or "/usr/local/cuda/bin/nvcc"
    return cand if Path(cand).exists() else "nvcc"


def _digest() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [REPO / "include" / "rspnet_b200.h"]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every .cu for sm_100a and link the shared library. Returns its path."""
    LIB_DIR.mkdir(exist_ok=True)
    stamp = LIB_DIR / "build.stamp"
    digest = _digest()
    if not force and LIB_PATH.exists() and stamp.exists() and stamp.read_text() == digest:
        return LIB_PATH
    objs = []
    procs = []
    for src in SOURCES:
        obj = LIB_DIR / (src + ".o")
        cmd = [_nvcc(), *[f for f in NVCC_FLAGS if f != "--use_fast_math=false"], "-I", str(REPO / "include"),
               "-I", str(CSRC), "-c", str(CSRC / src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(str(obj))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        if verbose and out:
            print(out)
    link = [_nvcc(), "-shared", "-o", str(LIB_PATH), *objs, "-lcudart"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    stamp.write_text(digest)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
