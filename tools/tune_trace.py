"""Compile-time tuning of the persistent trace loop.  `build` (run where nvcc is) compiles one library per
variant under build/variants/; `run` (on the GPU box) times configs[1] with each and prints one JSON line per
variant.  python tools/tune_trace.py build|run [scene]"""
import os, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
VARIANTS = {
    "tmin1": ["-DHL_TRI_MIN_LANES=1"], "tmin4": ["-DHL_TRI_MIN_LANES=4"], "tmin8": ["-DHL_TRI_MIN_LANES=8"], "tmin12": ["-DHL_TRI_MIN_LANES=12"],
    "tmin16": ["-DHL_TRI_MIN_LANES=16"], "tmin8_tri24": ["-DHL_TRI_MIN_LANES=8", "-DHL_TRI_PER_STEP=24"], "tmin12_tri24": ["-DHL_TRI_MIN_LANES=12", "-DHL_TRI_PER_STEP=24"],
    "tmin8_tri3": ["-DHL_TRI_MIN_LANES=8", "-DHL_TRI_PER_STEP=3"], "tmin8_refill4": ["-DHL_TRI_MIN_LANES=8", "-DHL_REFILL_MIN=4"],
    "tmin8_refill12": ["-DHL_TRI_MIN_LANES=8", "-DHL_REFILL_MIN=12"],
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
