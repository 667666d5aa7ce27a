#!/usr/bin/env python3
"""Writes tests/golden/ref_scene/*.tab and tests/golden/glm_pin.bin with the REFERENCE'S OWN CODE (oracle/_ref/ref_scene_tool =
src/engine/resource/scene.cpp + material.cpp compiled where they lie; oracle/_ref/glm_pin_ref = oracle/ref_scene/glm_pin.cpp
against the GLM the reference vendors).  Runs only where /root/reference is mounted."""
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from helios_b200 import scene_io  # noqa: E402
from oracle import oracle  # noqa: E402
from tests.test_ref_scene import REF_SCENES  # noqa: E402

tool, glm_pin = oracle.build_ref_scene()
assert tool is not None and glm_pin is not None, "needs /root/reference"
out = ROOT / "tests" / "golden" / "ref_scene"
out.mkdir(parents=True, exist_ok=True)
with tempfile.TemporaryDirectory() as d:
    for name in sorted(REF_SCENES):
        f = Path(d) / "s.hlsc"
        scene_io.export_scene(REF_SCENES[name](), f)
        r = subprocess.run([str(tool), str(f), str(out / f"{name}.tab")], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        print(name, r.stdout.strip(), (out / f"{name}.tab").stat().st_size, "bytes")
(ROOT / "tests" / "golden" / "glm_pin.bin").write_bytes(subprocess.run([str(glm_pin), "120"], capture_output=True).stdout)
print("glm_pin.bin", (ROOT / "tests" / "golden" / "glm_pin.bin").stat().st_size, "bytes")
