// utility/png_writer.h — 8-bit RGBA PNG encoder for Renderer::save_image_to_disk (reference:
// src/engine/gfx/renderer.cpp:651 `stbi_write_png(path, w, h, 4, mapped_ptr, 4 * w)`; stb_image_write is a
// third-party encoder and is not restated: this one emits the same pixels in a standards-conforming file).
// Layout: IHDR (8-bit, colour type 6) + one IDAT holding a zlib stream whose deflate blocks are fixed-Huffman
// coded run-length matches (distance = 4 bytes = one pixel, and distance = one row) over `Sub`-filtered rows,
// + IEND.  Returns false when the file cannot be written.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace helios
{
bool write_png_rgba8(const std::string& path, uint32_t width, uint32_t height, const uint8_t* rgba, size_t row_stride_bytes);
// the encoded file in memory (what write_png_rgba8 writes)
std::vector<uint8_t> encode_png_rgba8(uint32_t width, uint32_t height, const uint8_t* rgba, size_t row_stride_bytes);
} // namespace helios

extern "C" int helios_write_png_rgba8(const char* path, uint32_t width, uint32_t height, const uint8_t* rgba, uint32_t row_stride_bytes);
