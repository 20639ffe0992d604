"""Builds the CUDA engine in-tree: silero_vad_b200/lib/libsilero_vad_b200.so (sm_100a only).

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box with the tree.
"""
import hashlib
import os
import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
LIB = PKG / "lib" / "libsilero_vad_b200.so"
SOURCES = [PKG / "csrc" / "svad_api.cu", PKG / "csrc" / "svad_segments.cpp"]
HEADERS = sorted((PKG / "csrc").glob("*.h")) + sorted((PKG / "csrc").glob("*.cuh")) + [PKG.parent / "include" / "silero_vad_b200.h"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-pthread", "-Xptxas", "-v"] + os.environ.get("SVAD_EXTRA_NVCC", "").split()
HASH = PKG / "lib" / "source_hash.txt"


def nvcc_path():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found")


def source_hash():
    """sha256 over the sources, headers and flags the library is built from (content, not mtimes: the tree is copied
    to the GPU box, where file times say nothing)."""
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for p in sorted(SOURCES + HEADERS):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    return h.hexdigest()


def stale():
    """True when the library is missing or was built from other sources than the ones in the tree."""
    return not (LIB.exists() and HASH.exists() and HASH.read_text().strip() == source_hash())


def build_debug():
    """Second library with -DSVAD_H16_DEBUG (the fp16 kernel dumps every activation of CTA 0's first two steps): only the
    bring-up tools load it (SVAD_DEBUG_LIB=1)."""
    dbg = LIB.with_name("libsilero_vad_b200_dbg.so")
    r = subprocess.run([nvcc_path(), *NVCC_FLAGS, "-DSVAD_H16_DEBUG", "-o", str(dbg), *map(str, SOURCES)], capture_output=True, text=True)
    if r.returncode:
        print(r.stdout, r.stderr)
        raise RuntimeError("nvcc failed building %s" % dbg)
    return dbg


def build_variant(tag, defines):
    """Experiment library lib/libsilero_vad_b200_<tag>.so with extra -D flags (A/B runs of kernel variants on one GPU box);
    SVAD_DEBUG_LIB=<tag> makes the bring-up tools load it."""
    out = LIB.with_name("libsilero_vad_b200_%s.so" % tag)
    r = subprocess.run([nvcc_path(), *NVCC_FLAGS, *defines, "-o", str(out), *map(str, SOURCES)], capture_output=True, text=True)
    if r.returncode:
        print(r.stdout, r.stderr)
        raise RuntimeError("nvcc failed building %s" % out)
    return out


def build(force=False, verbose=False):
    """Compile if missing or built from other sources. Returns the library path."""
    tag = os.environ.get("SVAD_DEBUG_LIB")
    if tag:
        dbg = LIB.with_name("libsilero_vad_b200_%s.so" % ("dbg" if tag == "1" else tag))
        if not dbg.exists():
            raise RuntimeError("debug library missing: run silero_vad_b200.build.build_debug() / build_variant() first")
        return dbg
    if force or stale():
        LIB.parent.mkdir(exist_ok=True)
        cmd = [nvcc_path(), *NVCC_FLAGS, "-o", str(LIB), *map(str, SOURCES)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode:
            print(r.stdout, r.stderr)
        if r.returncode:
            raise RuntimeError("nvcc failed building %s" % LIB)
        (PKG / "lib" / "ptxas.log").write_text(r.stderr)
        HASH.write_text(source_hash() + "\n")
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
