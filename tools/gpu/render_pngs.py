"""renders of the BASELINE scenes through the C++ host (helios_headless) as PNG, for eyeballing: python tools/gpu/render_pngs.py outdir [spp]"""
import subprocess, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
from helios_b200 import scene_io, scenes
from helios_b200.build import build_shim

out = Path(sys.argv[1]); out.mkdir(parents=True, exist_ok=True)
spp = sys.argv[2] if len(sys.argv) > 2 else "64"
exe = str(build_shim())
for name, s in (("cornell", scenes.cornell_box(512, 512)), ("terrain", scenes.terrain_scene(width=960, height=540)), ("foliage", scenes.foliage_scene(n_clusters=20000, width=960, height=540)),
                ("city", scenes.city_scene(width=960, height=540))):
    f = out / f"{name}.hlsc"
    scene_io.export_scene(s, f)
    r = subprocess.run([exe, "--scene", str(f), "--spp", spp, "--out", str(out / f"{name}.png")], capture_output=True, text=True)
    print(name, r.returncode, r.stdout.strip().splitlines()[-1][:200] if r.stdout.strip() else r.stderr[-300:])
    f.unlink()
