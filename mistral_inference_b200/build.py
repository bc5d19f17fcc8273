"""Builds libmb200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

In-tree so the built .so travels to the GPU box with the gpurun snapshot (a JIT cache would not).
`python -m mistral_inference_b200.build [--force] [--verbose]`
"""
import hashlib
import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libmb200.so"
STAMP = PKG / "csrc" / ".build_stamp"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared",
    "--expt-relaxed-constexpr",
]


def _extra_defines():
    return [f"-D{d}" for d in os.environ.get("MB200_DEFINES", "").split() if d]


def _sources_digest() -> str:
    h = hashlib.sha256()
    for f in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "mistral_b200.h"]):
        h.update(f.name.encode())
        h.update(f.read_bytes())
    h.update(" ".join(NVCC_FLAGS + _extra_defines()).encode())
    return h.hexdigest()


def build_variant(name: str, defines) -> Path:
    """An extra in-tree library with additional -D flags (A/B experiments: MB200_LIB_PATH selects it at run time)."""
    out = PKG / f"libmb200_{name}.so"
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-o", str(out)] + [str(f) for f in sorted(CSRC.glob("*.cu"))]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError(f"nvcc failed ({proc.returncode})")
    return out


def build_library(force: bool = False, verbose: bool = False) -> Path:
    digest = _sources_digest()
    if not force and LIB.exists() and STAMP.exists() and STAMP.read_text().strip() == digest:
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, *NVCC_FLAGS, *_extra_defines(), "-Xptxas", "-v", "-o", str(LIB)] + [str(f) for f in sorted(CSRC.glob("*.cu"))]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    (CSRC / "ptxas.log").write_text(proc.stderr)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError(f"nvcc failed ({proc.returncode}): {' '.join(cmd)}")
    if verbose:
        print(proc.stderr)
    STAMP.write_text(digest)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
