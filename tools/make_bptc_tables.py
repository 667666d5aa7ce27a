#!/usr/bin/env python3
"""Regenerates the BPTC (BC7 / BC6H) partition and anchor tables of helios_b200/shim/src/bc_decode.cpp by DECODING PROBE
BLOCKS with oracle/_ref/ref_bc_tool (nvidia-texture-tools' decoder, the library the reference's asset pipeline uses): the
tables are constants of the format (Khronos Data Format Specification 1.3, BPTC chapter) — this script documents where the
numbers in the source come from and re-derives them.  Runs only where /root/reference is mounted.

  two subsets   BC7 mode 1 blocks, subset 0 endpoints black / subset 1 white, all indices 0 -> the subset of every pixel;
                endpoints (black, white) in both subsets, all index bits 1 -> the two pixels that decode darker are the anchors
  three subsets BC7 mode 2 blocks, the same with three grey levels"""
import json
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import oracle  # noqa: E402

TOOL = oracle.build_ref_bc()
assert TOOL is not None, "needs /root/reference"


def pack(fields):
    v = pos = 0
    for val, n in fields:
        v |= (val & ((1 << n) - 1)) << pos
        pos += n
    assert pos == 128
    return v.to_bytes(16, "little")


def decode(blocks):
    with tempfile.TemporaryDirectory() as d:
        (Path(d) / "i").write_bytes(b"".join(blocks))
        subprocess.check_call([str(TOOL), "7", str(Path(d) / "i"), str(Path(d) / "o")])
        return np.frombuffer((Path(d) / "o").read_bytes(), np.uint8).reshape(len(blocks), 16, 4)


def mode1(p, e, idx, pbits):
    return pack([(0b10, 2), (p, 6)] + [(x, 6) for _ in range(3) for x in e] + [(b, 1) for b in pbits] + [(idx, 46)])


def mode2(p, e, idx):
    return pack([(0b100, 3), (p, 6)] + [(x, 5) for _ in range(3) for x in e] + [(idx, 29)])


P2 = (decode([mode1(p, (0, 0, 63, 63), 0, (0, 1)) for p in range(64)])[:, :, 0] > 128).astype(int)
d = decode([mode1(p, (0, 63, 0, 63), (1 << 46) - 1, (0, 0)) for p in range(64)])
A2 = [[i for i in range(16) if d[p, i, 0] != d[p, :, 0].max()] for p in range(64)]
assert all(a[0] == 0 and len(a) == 2 for a in A2)
g = decode([mode2(p, (0, 0, 15, 15, 31, 31), 0) for p in range(64)])[:, :, 0]
P3 = np.where(g < 60, 0, np.where(g < 200, 1, 2))
d = decode([mode2(p, (0, 31) * 3, (1 << 29) - 1) for p in range(64)])
A3 = [[i for i in range(16) if d[p, i, 0] != d[p, :, 0].max()] for p in range(64)]
assert all(a[0] == 0 and len(a) == 3 for a in A3)
out = {
    "partition2": [hex(sum(int(P2[p, i]) << i for i in range(16))) for p in range(64)],
    "partition3": [hex(sum(int(P3[p, i]) << (2 * i) for i in range(16))) for p in range(64)],
    "anchor2": [a[1] for a in A2],
    "anchor3a": [[i for i in a[1:] if P3[p, i] == 1][0] for p, a in enumerate(A3)],
    "anchor3b": [[i for i in a[1:] if P3[p, i] == 2][0] for p, a in enumerate(A3)],
}
print(json.dumps(out, indent=1))
src = (ROOT / "helios_b200" / "shim" / "src" / "bc_decode.cpp").read_text()
missing = [v for v in out["partition2"] + out["partition3"] if ("0x%04x" % int(v, 16)) not in src and ("0x%08x" % int(v, 16)) not in src]
print("tables in bc_decode.cpp agree" if not missing else f"NOT in bc_decode.cpp: {missing}")
