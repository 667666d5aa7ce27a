#!/usr/bin/env python3
"""Extract the Hosek-Wilkie RGB coefficient tables into a compact binary file.

The sky model needs the published Hosek-Wilkie RGB dataset v1.4a
("An Analytic Model for Full Spectral Sky-Dome Radiance", Hosek & Wilkie, SIGGRAPH 2012), which the
reference vendors as include/gfx/hosek_data_rgb.inl.  This script parses the numeric tables out of that
file (data, not code) and writes them as raw little-endian float64:

    helios_b200/data/hosek_rgb_v1_4a.f64 =  3 x 1080 (datasetRGB1..3)  then  3 x 120 (datasetRGBRad1..3)

Run in the build container (where /root/reference exists); the output is committed so that the GPU box,
which has no reference checkout, can evaluate the sky.
"""
import re
import struct
import sys
from pathlib import Path

src = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/include/gfx/hosek_data_rgb.inl")
dst = Path(__file__).resolve().parent.parent / "helios_b200" / "data" / "hosek_rgb_v1_4a.f64"

text = src.read_text()
arrays = {}
for m in re.finditer(r"double\s+(\w+)\[\]\s*=\s*\{(.*?)\};", text, re.S):
    body = re.sub(r"//[^\n]*", "", m.group(2))
    vals = [float(x) for x in re.findall(r"[-+]?\d+\.?\d*(?:[eE][-+]?\d+)?", body)]
    arrays[m.group(1)] = vals
out = []
for i in (1, 2, 3):
    a = arrays[f"datasetRGB{i}"]
    assert len(a) == 1080, len(a)
    out += a
for i in (1, 2, 3):
    a = arrays[f"datasetRGBRad{i}"]
    assert len(a) == 120, len(a)
    out += a
dst.parent.mkdir(parents=True, exist_ok=True)
dst.write_bytes(struct.pack(f"<{len(out)}d", *out))
print(f"wrote {dst} ({len(out)} doubles)")
