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
SOURCES = ["hl_builder.cu", "hl_wavefront.cu", "hl_api.cu", "hl_comm.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "--fmad=false", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "-shared",
]
LINK_FLAGS = ["-ldl"]  # hl_comm.cu opens libnccl.so.2 at run time


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.h")) + [ROOT.parent / "include" / "helios_b200.h"]
    return any(d.stat().st_mtime > t for d in deps)


def build_library(force: bool = False, verbose: bool = False, defines: list[str] | None = None, out: Path | None = None) -> Path:
    """`defines` / `out` build a tuning variant (e.g. ["-DHL_REFILL_MIN=4"]) next to the product library.
    The translation units are compiled side by side (one nvcc process each), then linked."""
    lib = Path(out) if out else LIB
    if not defines and not force and not needs_build():
        return lib
    lib.parent.mkdir(parents=True, exist_ok=True)
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    objdir = ROOT.parent / "build" / ("obj_" + lib.stem)
    objdir.mkdir(parents=True, exist_ok=True)
    cflags = [f for f in NVCC_FLAGS if f != "-shared"]
    procs = []
    for src in SOURCES:
        cmd = [nvcc, *cflags, *(defines or []), *(["-Xptxas", "-v"] if verbose else []), "-c", "-o", str(objdir / (src + ".o")), str(CSRC / src)]
        print("[helios_b200] " + " ".join(cmd), file=sys.stderr)
        procs.append((cmd, subprocess.Popen(cmd)))
    for cmd, p in procs:
        if p.wait() != 0:
            raise subprocess.CalledProcessError(p.returncode, cmd)
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", str(lib), *[str(objdir / (src + ".o")) for src in SOURCES], *LINK_FLAGS]
    print("[helios_b200] " + " ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    return lib


TEST_VARIANTS = {
    # a 3-entry traversal stack: tests/test_gpu_edges.py checks that overflows are counted and hl_get_counters fails loudly
    "stack3": ["-DHL_STACK_FAST=2", "-DHL_STACK_SPILL=1"],
    # four shared-memory cost rows per treelet: nearly every treelet of the two-level BVH re-split is fitted through global memory
    # (the path a treelet with a long chain of lopsided splits takes): tests/test_gpu_edges.py compares the trees
    "rows4": ["-DHL_TREELET_ROWS=4"],
}


def build_test_variants(force: bool = False) -> dict:
    """test-only builds of the library with other compile-time constants, in-tree (tests/_variants/, git-ignored) so that
    they travel to the GPU box with the snapshot"""
    out = {}
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.h")) + [ROOT.parent / "include" / "helios_b200.h"]
    for name, defines in TEST_VARIANTS.items():
        lib = ROOT.parent / "tests" / "_variants" / f"libhelios_b200_{name}.so"
        if force or not lib.exists() or any(d.stat().st_mtime > lib.stat().st_mtime for d in deps):
            build_library(force=True, defines=defines, out=lib)
        out[name] = lib
    return out


SHIM = ROOT / "shim"
ENGINE_LIB = ROOT / "libhelios_engine.so"
HEADLESS = ROOT / "helios_headless"


def build_shim(force: bool = False) -> Path:
    """C++ host layer with the reference's class surface (helios_b200/shim) + the headless driver.  Plain g++:
    it only talks to the C ABI, so it links against libhelios_b200.so and needs no CUDA headers."""
    srcs = sorted((SHIM / "src").glob("*.cpp"))
    hdrs = list((SHIM / "include").rglob("*.h")) + list((SHIM / "include").rglob("*.hpp")) + [ROOT.parent / "include" / "helios_b200.h"]
    tool = SHIM / "tools" / "helios_headless.cpp"
    newest = max(f.stat().st_mtime for f in [*srcs, *hdrs, tool])
    if not force and ENGINE_LIB.exists() and HEADLESS.exists() and min(ENGINE_LIB.stat().st_mtime, HEADLESS.stat().st_mtime) > newest and ENGINE_LIB.stat().st_mtime > LIB.stat().st_mtime:
        return HEADLESS
    cxx = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++"
    common = [cxx, "-std=c++17", "-O2", "-fPIC", "-Wall", "-Wno-unused-function", f"-I{SHIM / 'include'}", f"-I{ROOT.parent / 'include'}"]
    link = [f"-L{ROOT}", "-lhelios_b200", "-ldl", "-Wl,-rpath,$ORIGIN"]
    cmd = [*common, "-shared", "-o", str(ENGINE_LIB), *map(str, srcs), *link]
    print("[helios_b200] " + " ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    cmd = [*common, "-o", str(HEADLESS), str(tool), "-lhelios_engine", *link]
    print("[helios_b200] " + " ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    return HEADLESS


if __name__ == "__main__":
    build_library(force="--force" in sys.argv, verbose="-v" in sys.argv)
    build_shim(force="--force" in sys.argv)
