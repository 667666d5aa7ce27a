"""BC7 and BC6H texture payloads (the reference's format table maps them to VK_FORMAT_BC7_* / BC6H_UFLOAT,
src/engine/core/resource_manager.cpp:14-28, and lets the Vulkan sampler decode them): the host-side decoders of the C++
host layer (helios_b200/shim/src/bc_decode.cpp) against the decoder of the library the reference's asset pipeline itself
uses — nvidia-texture-tools, vendored by AssetCore, compiled where it lies into oracle/_ref/ref_bc_tool (oracle/ref_bc/).

Bar: bit-exact on random blocks of every mode (BC7: 8 modes + the reserved pattern; BC6H: 14 modes + 4 reserved patterns,
unsigned and signed).  Random 128-bit blocks exercise every partition, rotation, index-selection and delta combination.
The oracle's outputs are pinned as SHA-256 digests (tests/golden/bc67_golden.json, written by this file's __main__ where
/root/reference is mounted) so the decoders are held to the reference library's results on machines without it too.
End to end: an AssetCore image file carrying such blocks, loaded through ast::load_image + the ResourceManager's level-0
conversion (tests/ast/my_ast_dump texels), gives the same texels."""
import hashlib
import json
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = ROOT / "tests" / "golden" / "bc67_golden.json"
BC7_MODES = 9  # 8 modes + reserved (all mode bits zero)
BC6_MODES = [(2, 0), (2, 1), (5, 2), (5, 6), (5, 10), (5, 14), (5, 18), (5, 22), (5, 26), (5, 30), (5, 3), (5, 7), (5, 11), (5, 15), (5, 19), (5, 23), (5, 27), (5, 31)]

HARNESS = r"""
#include <utility/bc_decode.h>
#include <cstdio>
#include <cstring>
#include <vector>
int main(int argc, char** argv)
{
    FILE* f = fopen(argv[2], "rb");
    std::vector<unsigned char> in;
    unsigned char buf[4096];
    size_t n;
    while ((n = fread(buf, 1, 4096, f)) > 0) in.insert(in.end(), buf, buf + n);
    fclose(f);
    FILE* o = fopen(argv[3], "wb");
    for (size_t b = 0; b < in.size() / 16; b++)
    {
        if (argv[1][0] == '7')
        {
            std::vector<uint8_t> out;
            if (!helios::decode_bc(9, in.data() + 16 * b, 16, 4, 4, out)) return 1;
            fwrite(out.data(), 1, 64, o);
        }
        else
        {
            std::vector<uint16_t> out;
            if (!helios::decode_bc6h(in.data() + 16 * b, 16, 4, 4, argv[1][1] == 's', out)) return 1;
            fwrite(out.data(), 2, 48, o);
        }
    }
    fclose(o);
    return 0;
}
"""


def bc7_blocks(n, seed):
    rng = np.random.default_rng(seed)
    blocks = rng.integers(0, 256, (n, 16), dtype=np.uint8)
    for i in range(n):  # an even spread over the modes: mode m = m zero bits, then a one
        m = i % BC7_MODES
        v = int.from_bytes(blocks[i].tobytes(), "little") & ~((1 << (m + 1)) - 1)
        if m < 8:
            v |= 1 << m
        blocks[i] = np.frombuffer(v.to_bytes(16, "little"), np.uint8)
    return blocks


def bc6_blocks(n, seed):
    rng = np.random.default_rng(seed)
    blocks = rng.integers(0, 256, (n, 16), dtype=np.uint8)
    for i in range(n):
        nb, mv = BC6_MODES[i % len(BC6_MODES)]
        blocks[i, 0] = (int(blocks[i, 0]) & ~((1 << nb) - 1) & 255) | mv
    return blocks


@pytest.fixture(scope="module")
def mine(tmp_path_factory):
    d = tmp_path_factory.mktemp("bc67")
    (d / "h.cpp").write_text(HARNESS)
    exe = d / "bc_mine"
    subprocess.check_call(["g++", "-O1", "-std=c++17", f"-I{ROOT / 'helios_b200' / 'shim' / 'include'}", "-o", str(exe), str(d / "h.cpp"), str(ROOT / "helios_b200" / "shim" / "src" / "bc_decode.cpp")])

    def run(kind, blocks):
        (d / "in.bin").write_bytes(blocks.tobytes())
        subprocess.check_call([str(exe), kind, str(d / "in.bin"), str(d / "out.bin")])
        return (d / "out.bin").read_bytes()

    return run


def oracle_run(tool, kind, blocks, tmp):
    (tmp / "in.bin").write_bytes(blocks.tobytes())
    subprocess.check_call([str(tool), kind, str(tmp / "in.bin"), str(tmp / "ref.bin")])
    return (tmp / "ref.bin").read_bytes()


CASES = [("7", bc7_blocks, 9 * 400, 11), ("6u", bc6_blocks, 18 * 200, 12), ("6s", bc6_blocks, 18 * 200, 13)]


@pytest.mark.parametrize("kind,gen,n,seed", CASES)
def test_decoders_equal_the_reference_library(kind, gen, n, seed, mine, tmp_path):
    from oracle import oracle

    blocks = gen(n, seed)
    got = mine(kind, blocks)
    gold = json.loads(GOLDEN.read_text())
    assert hashlib.sha256(blocks.tobytes()).hexdigest() == gold[kind]["blocks_sha256"], "block generator changed"
    assert hashlib.sha256(got).hexdigest() == gold[kind]["decoded_sha256"], f"{kind}: decoder output differs from nvidia-texture-tools' (committed digest)"
    tool = oracle.build_ref_bc()
    if tool is not None:  # the library itself, here and now, on ten times as many blocks
        assert oracle_run(tool, kind, blocks, tmp_path) == got
        more = gen(10 * n, seed + 100)
        a = np.frombuffer(oracle_run(tool, kind, more, tmp_path), np.uint8).reshape(10 * n, -1)
        b = np.frombuffer(mine(kind, more), np.uint8).reshape(10 * n, -1)
        bad = (a != b).any(1)
        per_mode = [int(bad[i :: (BC7_MODES if kind == "7" else len(BC6_MODES))].sum()) for i in range(BC7_MODES if kind == "7" else len(BC6_MODES))]
        assert not bad.any(), f"{kind}: mismatching blocks per mode {per_mode}"


def test_bc7_known_blocks(mine):
    """hand-made blocks whose decode follows from the format alone: a mode-6 block (one subset, 7-bit RGBA + P bit, 4-bit
    indices) with endpoints 0 and 127|P=1 -> 0 and 255; index i on pixel i -> the 4-bit weight table"""
    v = 1 << 6  # mode 6
    pos = 7
    for ch in range(4):  # r0 r1 g0 g1 b0 b1 a0 a1, 7 bits each
        v |= 0 << pos
        pos += 7
        v |= 127 << pos
        pos += 7
    v |= 0 << pos  # P0
    pos += 1
    v |= 1 << pos  # P1
    pos += 1
    for i in range(16):  # pixel 0 (anchor) has 3 index bits
        nb = 3 if i == 0 else 4
        v |= (i & ((1 << nb) - 1)) << pos
        pos += nb
    assert pos == 128
    out = np.frombuffer(mine("7", np.frombuffer(v.to_bytes(16, "little"), np.uint8).reshape(1, 16)), np.uint8).reshape(16, 4)
    w = [0, 4, 9, 13, 17, 21, 26, 30, 34, 38, 43, 47, 51, 55, 60, 64]
    want = [((64 - w[i]) * 0 + w[i] * 255 + 32) >> 6 for i in range(16)]
    assert [int(x) for x in out[:, 0]] == want and np.all(out == out[:, :1])


from tests.test_ast_loader import my_dump, run  # noqa: E402,F401  (the loader harness fixture: tests/ast/my_ast_dump)


@pytest.mark.parametrize("kind,comp", [("7", 9), ("6u", 8)])
def test_bc67_images_through_the_asset_loader(kind, comp, mine, my_dump, tmp_path):
    """an AssetCore image file (helios_b200/ast_io.py writes the container the reference's exporter writes) with BC7 / BC6H
    blocks, 10 x 6 texels (ragged: 3 x 2 blocks), loaded by ast::load_image and converted as ResourceManager::load_texture_2d
    does — the texels are the block decoder's, block by block"""
    from helios_b200 import abi, ast_io

    w, h = 10, 6
    blocks = (bc7_blocks if kind == "7" else bc6_blocks)(6, 5)
    ast_io.write_image(tmp_path / "img.ast", "img", [[(w, h, blocks.tobytes())]], 4 if kind == "7" else 3, ast_io.PIXEL_UNORM8 if kind == "7" else ast_io.PIXEL_FLOAT16, compression=comp)
    head = run(my_dump, "texels", tmp_path / "img.ast", 0, 0, tmp_path / "t.bin").split()
    per_block = np.frombuffer(mine(kind, blocks), np.uint8 if kind == "7" else np.uint16).reshape(6, 16, 4 if kind == "7" else 3)
    if kind == "7":
        assert head == [str(abi.TEX_RGBA8_UNORM), str(w), str(h)]
        got = np.fromfile(tmp_path / "t.bin", np.uint8).reshape(h, w, 4)
    else:
        assert head == [str(abi.TEX_RGBA32F), str(w), str(h)]
        got = np.fromfile(tmp_path / "t.bin", np.float32).reshape(h, w, 4)
        assert np.all(got[..., 3] == 1.0)
        got = got[..., :3]
        per_block = per_block.view(np.float16).astype(np.float32)
    for by in range(2):
        for bx in range(3):
            for i in range(16):
                x, y = bx * 4 + (i & 3), by * 4 + (i >> 2)
                if x < w and y < h:
                    assert np.array_equal(got[y, x], per_block[by * 3 + bx, i], equal_nan=True), (bx, by, i)
    if kind == "7":  # BC7 has an sRGB VkFormat, BC6H has none (kCompressedFormats)
        assert run(my_dump, "texels", tmp_path / "img.ast", 1, 0, tmp_path / "t.bin").split() == [str(abi.TEX_RGBA8_SRGB), str(w), str(h)]
    else:
        assert run(my_dump, "texels", tmp_path / "img.ast", 1, 0, tmp_path / "t.bin").split() == ["no", "format"]


if __name__ == "__main__":  # regenerate the golden digests with the reference's library
    sys.path.insert(0, str(ROOT))
    import tempfile

    from oracle import oracle

    tool = oracle.build_ref_bc()
    assert tool is not None, "needs /root/reference"
    gold = {"source": "oracle/_ref/ref_bc_tool = nvidia-texture-tools bc7/ + bc6h/ decoders (AssetCore's vendored copy), SHA-256 of the decoded bytes"}
    with tempfile.TemporaryDirectory() as d:
        for kind, gen, n, seed in CASES:
            blocks = gen(n, seed)
            gold[kind] = {"blocks": n, "seed": seed, "blocks_sha256": hashlib.sha256(blocks.tobytes()).hexdigest(), "decoded_sha256": hashlib.sha256(oracle_run(tool, kind, blocks, Path(d))).hexdigest()}
    GOLDEN.write_text(json.dumps(gold, indent=1) + "\n")
    print(json.dumps(gold, indent=1))
