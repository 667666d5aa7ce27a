// utility/bc_decode.h — host-side decoders for the block-compressed texture payloads AssetCore images can carry
// (include/common/image.h:9-24 -> VK_FORMAT_BC*_BLOCK in core/resource_manager.cpp:14-28; the reference hands the
// blocks to the Vulkan sampler, which decodes them in hardware — here they are expanded once at load time to
// RGBA8 and uploaded as ordinary texels, so the traversal / shading kernels sample one texel layout).
// Block layouts after the Direct3D / Khronos Data Format specifications: BC1 (with and without 1-bit alpha),
// BC2, BC3, BC4 (one channel), BC5 (two channels), BC7 (all eight modes); BC6H (HDR, all fourteen modes) through decode_bc6h.
#pragma once
#include <cstdint>
#include <vector>

namespace helios
{
// compression: ast::CompressionType value (1 BC1, 2 BC1a, 3 BC2, 4 BC3, 5 BC3n, 6 BC4, 7 BC5, 9 BC7).  `blocks` holds
// ceil(w/4) * ceil(h/4) blocks in row-major order.  out = width * height RGBA8 (BC4: r,0,0,255; BC5: r,g,0,255).
bool decode_bc(int compression, const uint8_t* blocks, size_t n_bytes, uint32_t width, uint32_t height, std::vector<uint8_t>& out_rgba8);
// BC6H (ast::COMPRESSION_BC6): out = width * height * 3 half-float bit patterns (r, g, b); is_signed selects the SFLOAT variant
// (the reference's format table only uses VK_FORMAT_BC6H_UFLOAT_BLOCK, core/resource_manager.cpp:14-28)
bool decode_bc6h(const uint8_t* blocks, size_t n_bytes, uint32_t width, uint32_t height, bool is_signed, std::vector<uint16_t>& out_rgb16f);
// bytes per 4x4 block of a compression type, 0 when it is not block compressed / unknown
uint32_t bc_block_bytes(int compression);
} // namespace helios
