"""Renderer::save_image_to_disk's encoder (reference: src/engine/gfx/renderer.cpp:651, stbi_write_png of the 8-bit
RGBA tone-mapped image).  The file written by the C++ host layer is decoded here with an independent decoder
(zlib inflate + PNG unfilter + chunk CRC checks) and must return the input pixels exactly."""
import ctypes
import struct
import zlib

import numpy as np
import pytest

from helios_b200.build import ENGINE_LIB, build_library, build_shim


@pytest.fixture(scope="module")
def engine():
    build_library()
    build_shim()
    lib = ctypes.CDLL(str(ENGINE_LIB))
    lib.helios_write_png_rgba8.argtypes = [ctypes.c_char_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_uint32]
    lib.helios_write_png_rgba8.restype = ctypes.c_int
    return lib


def decode_png(data: bytes) -> np.ndarray:
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, hdr, seen_end = 8, b"", None, False
    while pos < len(data):
        n, typ = struct.unpack(">I4s", data[pos : pos + 8])
        body = data[pos + 8 : pos + 8 + n]
        (crc,) = struct.unpack(">I", data[pos + 8 + n : pos + 12 + n])
        assert zlib.crc32(typ + body) == crc, typ
        pos += 12 + n
        if typ == b"IHDR":
            hdr = struct.unpack(">IIBBBBB", body)
        elif typ == b"IDAT":
            idat += body
        elif typ == b"IEND":
            seen_end = True
    assert seen_end and hdr is not None
    w, h, depth, ctype, comp, flt, interlace = hdr
    assert (depth, ctype, comp, flt, interlace) == (8, 6, 0, 0, 0)
    raw = zlib.decompress(idat)  # verifies the Adler-32 trailer too
    assert len(raw) == h * (w * 4 + 1)
    rows = np.frombuffer(raw, np.uint8).reshape(h, w * 4 + 1)
    out = np.zeros((h, w * 4), np.uint8)
    for y in range(h):
        f, line = int(rows[y, 0]), rows[y, 1:].astype(np.int64)
        up = out[y - 1].astype(np.int64) if y else np.zeros(w * 4, np.int64)
        if f == 0:
            cur = line
        elif f == 1:  # Sub: prefix sum per channel mod 256
            cur = (np.cumsum(line.reshape(w, 4), axis=0) % 256).reshape(-1)
        elif f == 2:
            cur = (line + up) % 256
        else:
            raise AssertionError(f"filter {f} not expected from this encoder")
        out[y] = cur.astype(np.uint8)
    return out.reshape(h, w, 4)


def images():
    rng = np.random.default_rng(7)
    yield "noise", rng.integers(0, 256, (37, 53, 4), dtype=np.uint8)
    g = np.zeros((64, 96, 4), np.uint8)
    g[..., 0] = np.arange(96)[None, :] * 2
    g[..., 1] = np.arange(64)[:, None] * 3
    g[..., 2] = 40
    g[..., 3] = 255
    yield "gradient", g
    yield "flat", np.full((120, 300, 4), 200, np.uint8)  # long runs: matches of the maximum length 258
    yield "one_pixel", np.array([[[1, 2, 3, 4]]], np.uint8)
    yield "one_column", rng.integers(0, 256, (70, 1, 4), dtype=np.uint8)
    big = np.zeros((270, 480, 4), np.uint8)  # > 32 KiB window, mixed content
    big[..., :3] = (rng.random((270, 480, 3)) ** 3 * 255).astype(np.uint8)
    big[100:200, 50:400] = (10, 20, 30, 255)
    yield "mixed", big


@pytest.mark.parametrize("name,img", list(images()), ids=[n for n, _ in images()])
def test_png_round_trip(engine, tmp_path, name, img):
    p = tmp_path / f"{name}.png"
    img = np.ascontiguousarray(img)
    h, w, _ = img.shape
    assert engine.helios_write_png_rgba8(str(p).encode(), w, h, img.ctypes.data, w * 4) == 0
    got = decode_png(p.read_bytes())
    assert np.array_equal(got, img)
    if name == "flat":
        assert p.stat().st_size < img.nbytes // 50  # the run-length matches are really used


def test_png_row_stride_and_errors(engine, tmp_path):
    rng = np.random.default_rng(3)
    padded = rng.integers(0, 256, (20, 40, 4), dtype=np.uint8)
    p = tmp_path / "stride.png"
    assert engine.helios_write_png_rgba8(str(p).encode(), 25, 20, padded.ctypes.data, 160) == 0
    assert np.array_equal(decode_png(p.read_bytes()), padded[:, :25])
    # reference: stbi_write_png returns 0 -> HELIOS_LOG_ERROR (renderer.cpp:651-653); here a non-zero status
    assert engine.helios_write_png_rgba8(str(tmp_path / "no_such_dir" / "x.png").encode(), 25, 20, padded.ctypes.data, 160) != 0
    assert engine.helios_write_png_rgba8(str(p).encode(), 0, 20, padded.ctypes.data, 160) != 0
