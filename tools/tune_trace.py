"""Compile-time tuning of the persistent trace loop.  `build` (run where nvcc is) compiles one library per
variant under build/variants/; `run` (on the GPU box) times configs[1] with each and prints one JSON line per
variant.  HL_TUNE_SET=occ|sched python tools/tune_trace.py build|run [scene]"""
import os, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
VARIANT_SETS = {
    "sched": {
        "tmin1": ["-DHL_TRI_MIN_LANES=1"], "tmin4": ["-DHL_TRI_MIN_LANES=4"], "tmin8": ["-DHL_TRI_MIN_LANES=8"], "tmin12": ["-DHL_TRI_MIN_LANES=12"],
        "tmin16": ["-DHL_TRI_MIN_LANES=16"], "tmin8_tri24": ["-DHL_TRI_MIN_LANES=8", "-DHL_TRI_PER_STEP=24"], "tmin12_tri24": ["-DHL_TRI_MIN_LANES=12", "-DHL_TRI_PER_STEP=24"],
        "tmin8_tri3": ["-DHL_TRI_MIN_LANES=8", "-DHL_TRI_PER_STEP=3"], "tmin8_refill4": ["-DHL_TRI_MIN_LANES=8", "-DHL_REFILL_MIN=4"],
        "tmin8_refill12": ["-DHL_TRI_MIN_LANES=8", "-DHL_REFILL_MIN=12"],
    },
    # occupancy of the persistent trace kernels: register cap (launch bounds), CTA size, fast-stack depth, CTAs per SM
    "occ": {
        "base": [],
        "mb8": ["-DHL_TRACE_MIN_BLOCKS=8"],
        "mb8_s8": ["-DHL_TRACE_MIN_BLOCKS=8", "-DHL_STACK_FAST=8"],
        "mb10_s8": ["-DHL_TRACE_MIN_BLOCKS=10", "-DHL_STACK_FAST=8", "-DHL_TRACE_GRID_MULT=10"],
        "b64_mb16": ["-DHL_TRACE_BLOCK=64", "-DHL_TRACE_MIN_BLOCKS=16", "-DHL_TRACE_GRID_MULT=16"],
        "b256_mb4": ["-DHL_TRACE_BLOCK=256", "-DHL_TRACE_MIN_BLOCKS=4", "-DHL_TRACE_GRID_MULT=4"],
        "g7": ["-DHL_TRACE_GRID_MULT=7"],
        "g6": ["-DHL_TRACE_GRID_MULT=6"],
        "s16": ["-DHL_STACK_FAST=16"],
        "refill4": ["-DHL_REFILL_MIN=4"], "refill16": ["-DHL_REFILL_MIN=16"],
    },
}
VARIANT_SETS["sched2"] = {
    "base": [], "tri1": ["-DHL_TRI_PER_STEP=1"], "tri3": ["-DHL_TRI_PER_STEP=3"], "tri4": ["-DHL_TRI_PER_STEP=4"], "tri24": ["-DHL_TRI_PER_STEP=24"],
    "tmin4": ["-DHL_TRI_MIN_LANES=4"], "tmin8": ["-DHL_TRI_MIN_LANES=8"], "refill4": ["-DHL_REFILL_MIN=4"], "refill12": ["-DHL_REFILL_MIN=12"], "refill16": ["-DHL_REFILL_MIN=16"],
    "refill1": ["-DHL_REFILL_MIN=1"], "s8": ["-DHL_STACK_FAST=8"],
}
VARIANT_SETS["shade"] = {
    "base": [], "mb5": ["-DHL_SHADE_MIN_BLOCKS=5"], "mb6": ["-DHL_SHADE_MIN_BLOCKS=6"], "mb8": ["-DHL_SHADE_MIN_BLOCKS=8"],
    "b64_mb12": ["-DHL_SHADE_BLOCK=64", "-DHL_SHADE_MIN_BLOCKS=12", "-DHL_SHADE_GRID_MULT=16"], "b256_mb3": ["-DHL_SHADE_BLOCK=256", "-DHL_SHADE_MIN_BLOCKS=3", "-DHL_SHADE_GRID_MULT=4"],
    "g16": ["-DHL_SHADE_GRID_MULT=16"], "g4": ["-DHL_SHADE_GRID_MULT=4"],
}
VARIANT_SETS["sah"] = {"c020": ["-DHL_SAH_C_PRIM_TRIANGLE=0.2f"], "c035": [], "c060": ["-DHL_SAH_C_PRIM_TRIANGLE=0.6f"], "c100": ["-DHL_SAH_C_PRIM_TRIANGLE=1.0f"], "c150": ["-DHL_SAH_C_PRIM_TRIANGLE=1.5f"], "c250": ["-DHL_SAH_C_PRIM_TRIANGLE=2.5f"]}
VARIANT_SETS["defslots"] = {"d4": [], "d6": ["-DHL_DEFAULT_WAVE_SLOTS=6"], "d8": ["-DHL_DEFAULT_WAVE_SLOTS=8"]}
VARIANT_SETS["slots"] = {"base": [], "slots3": ["-DHL_WAVE_SLOTS=3"], "slots2": ["-DHL_WAVE_SLOTS=2"], "slots4": ["-DHL_WAVE_SLOTS=4"], "slots6": ["-DHL_WAVE_SLOTS=6"], "mb7": ["-DHL_TRACE_MIN_BLOCKS=7"], "mb6": ["-DHL_TRACE_MIN_BLOCKS=6", "-DHL_TRACE_GRID_MULT=6"]}
VARIANT_SETS["coop"] = {
    "base0": ["-DHL_COOP_LEAVES=0"], "coop4": [], "coop4_mb6": ["-DHL_TRACE_MIN_BLOCKS=6"], "coop4_mb7": ["-DHL_TRACE_MIN_BLOCKS=7"],
    "coop2": ["-DHL_COOP_K=2"], "coop2_mb7": ["-DHL_COOP_K=2", "-DHL_TRACE_MIN_BLOCKS=7"], "coop3_mb6": ["-DHL_COOP_K=3", "-DHL_TRACE_MIN_BLOCKS=6"],
}
VARIANT_SETS["fol"] = {"base": [], "tri3": ["-DHL_TRI_PER_STEP=3"], "tri4": ["-DHL_TRI_PER_STEP=4"], "tri1": ["-DHL_TRI_PER_STEP=1"], "refill4": ["-DHL_REFILL_MIN=4"], "refill16": ["-DHL_REFILL_MIN=16"]}
VARIANT_SETS["treelet"] = {"base": [], "tiny1": ["-DHL_TREELET_TINY=1"], "exact3": ["-DHL_TREELET_EXACT=3"], "prims256": ["-DHL_TREELET_PRIMS=256"], "prims1024": ["-DHL_TREELET_PRIMS=1024"]}
VARIANT_SETS["grid"] = {"g8": [], "g4": ["-DHL_TRACE_GRID_MULT=4"], "g5": ["-DHL_TRACE_GRID_MULT=5"], "g6": ["-DHL_TRACE_GRID_MULT=6"], "g7": ["-DHL_TRACE_GRID_MULT=7"], "g12": ["-DHL_TRACE_GRID_MULT=12"]}
VARIANT_SETS["prefetch"] = {"base": [], "pf": ["-DHL_PREFETCH_NEXT_NODE=1"]}
VARIANTS = VARIANT_SETS[os.environ.get("HL_TUNE_SET", "occ")]
OUT = ROOT / "build" / "variants"
if sys.argv[1] == "build":
    from concurrent.futures import ThreadPoolExecutor
    from helios_b200.build import build_library
    with ThreadPoolExecutor(int(os.environ.get("HL_TUNE_JOBS", "6"))) as ex:  # nvcc runs as subprocesses
        list(ex.map(lambda nd: build_library(force=True, defines=nd[1], out=OUT / f"lib_{nd[0]}.so"), VARIANTS.items()))
else:
    scene = sys.argv[2:] or ["terrain"]
    for n in VARIANTS:
        env = dict(os.environ, HELIOS_B200_LIB=str(OUT / f"lib_{n}.so"))
        r = subprocess.run([sys.executable, str(ROOT / "tools" / "frame_time.py"), *scene], env=env, capture_output=True, text=True)
        print(n, r.stdout.strip() or r.stderr.strip()[-400:], flush=True)
