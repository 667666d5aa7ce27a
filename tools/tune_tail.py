"""GPU tuning sweep for the wavefront scheduler knobs (hl_set_option): ms/frame on the configs[1] scene."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from helios_b200 import scenes, api
from helios_b200.sky import sky_coefficients

s = scenes.terrain_scene()
ctx = api.Context(s.width, s.height)
ctx.load_scene(s, sky_coeffs=sky_coefficients(s.sun_direction))
pcs = [s.push_constants(f) for f in range(1, 41)]
ref = None
for start, thr in [(2, 0), (1, 32768), (1, 131072), (1, 524288), (2, 32768), (2, 98304), (2, 262144), (2, 1048576), (3, 98304), (1, 4000000)]:
    ctx.set_option(2, start); ctx.set_option(1, thr)
    ctx.accum_clear()
    for pc in pcs[:4]: ctx.render_frame(pc)
    ctx.reset_counters(); ctx.synchronize()
    ctx.event_record(0)
    for pc in pcs[4:36]: ctx.render_frame(pc)
    ctx.event_record(1)
    ms = ctx.event_elapsed_ms(0, 1) / 32
    c = ctx.counters()
    rays = float(c["extension_rays"] + c["shadow_rays"]) / 32
    acc = ctx.read_accum()
    if ref is None: ref = acc
    print(f"tail_start {start} threshold {thr:8d}: {ms:.3f} ms/frame  {rays/ms/1e3:.1f} Mrays/s  rays/frame {rays:.0f}  max|diff vs wavefront| {np.abs(acc-ref).max():.3g}")
ctx.close()
