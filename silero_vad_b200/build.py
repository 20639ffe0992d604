"""Builds the CUDA engine in-tree: silero_vad_b200/lib/libsilero_vad_b200.so (sm_100a only).

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box with the tree.
"""
import os
import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
LIB = PKG / "lib" / "libsilero_vad_b200.so"
SOURCES = [PKG / "csrc" / "svad_api.cu", PKG / "csrc" / "svad_segments.cpp"]
HEADERS = sorted((PKG / "csrc").glob("*.h")) + [PKG.parent / "include" / "silero_vad_b200.h"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found")


def stale():
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    return any(p.stat().st_mtime > t for p in SOURCES + HEADERS)


def build(force=False, verbose=False):
    """Compile if missing or older than its sources. Returns the library path."""
    if force or stale():
        LIB.parent.mkdir(exist_ok=True)
        cmd = [nvcc_path(), *NVCC_FLAGS, "-o", str(LIB), *map(str, SOURCES)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode:
            print(r.stdout, r.stderr)
        if r.returncode:
            raise RuntimeError("nvcc failed building %s" % LIB)
        (PKG / "lib" / "ptxas.log").write_text(r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
