"""Builds libhelios_b200.so (hand-written CUDA for sm_100a + the C ABI) in-tree with nvcc.

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.  --fmad=false keeps every fp32
a*b+c as two IEEE roundings so that hit parameters match the CPU oracle bit for bit; the kernels call fmaf()
explicitly where a fused multiply-add is wanted (node slab tests).
"""
from __future__ import annotations

import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent
CSRC = ROOT / "csrc"
LIB = ROOT / "libhelios_b200.so"
SOURCES = ["hl_builder.cu", "hl_wavefront.cu", "hl_api.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "--fmad=false", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "-shared",
]


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.h")) + [ROOT.parent / "include" / "helios_b200.h"]
    return any(d.stat().st_mtime > t for d in deps)


def build_library(force: bool = False, verbose: bool = False, defines: list[str] | None = None, out: Path | None = None) -> Path:
    """`defines` / `out` build a tuning variant (e.g. ["-DHL_REFILL_MIN=4"]) next to the product library."""
    lib = Path(out) if out else LIB
    if not defines and not force and not needs_build():
        return lib
    lib.parent.mkdir(parents=True, exist_ok=True)
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc, *NVCC_FLAGS, *(defines or []), *( ["-Xptxas", "-v"] if verbose else []), "-o", str(lib), *[str(CSRC / s) for s in SOURCES]]
    print("[helios_b200] " + " ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    return lib


if __name__ == "__main__":
    build_library(force="--force" in sys.argv, verbose="-v" in sys.argv)
