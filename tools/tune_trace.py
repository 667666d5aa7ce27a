"""Compile-time tuning of the persistent trace loop.  `build` (run where nvcc is) compiles one library per
variant under build/variants/; `run` (on the GPU box) times configs[1] with each and prints one JSON line per
variant.  python tools/tune_trace.py build|run [scene]"""
import os, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
VARIANTS = {
    "refill1": ["-DHL_REFILL_MIN=1"], "refill4": ["-DHL_REFILL_MIN=4"], "refill8": ["-DHL_REFILL_MIN=8"], "refill16": ["-DHL_REFILL_MIN=16"],
    "refill32": ["-DHL_REFILL_MIN=32"],
    "refill8_tri1": ["-DHL_REFILL_MIN=8", "-DHL_TRI_PER_STEP=1"], "refill8_tri2": ["-DHL_REFILL_MIN=8", "-DHL_TRI_PER_STEP=2"],
    "refill4_tri2": ["-DHL_REFILL_MIN=4", "-DHL_TRI_PER_STEP=2"],
}
OUT = ROOT / "build" / "variants"
if sys.argv[1] == "build":
    from helios_b200.build import build_library
    for n, d in VARIANTS.items():
        build_library(defines=d, out=OUT / f"lib_{n}.so")
else:
    scene = sys.argv[2:] or ["terrain"]
    for n in VARIANTS:
        env = dict(os.environ, HELIOS_B200_LIB=str(OUT / f"lib_{n}.so"))
        r = subprocess.run([sys.executable, str(ROOT / "tools" / "frame_time.py"), *scene], env=env, capture_output=True, text=True)
        print(n, r.stdout.strip() or r.stderr.strip()[-400:], flush=True)
