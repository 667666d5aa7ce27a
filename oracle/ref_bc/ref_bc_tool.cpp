// ORACLE — TEST INFRASTRUCTURE ONLY (never linked into the product).
// ref_bc_tool: BC7 and BC6H block decoding by the library the reference's asset pipeline itself uses — nvidia-texture-tools
// (vendored by AssetCore under external/AssetCore/external/nvidia-texture-tools, version file VERSION; the exporter
// external/AssetCore/src/exporter/image_exporter.cpp compresses with it, nvimage/BlockDXT.cpp:635-674 decodes through
// AVPCL::decompress / ZOH::decompress).  Its bc7/ and bc6h/ sources are compiled where they lie (oracle/Makefile target
// ref_bc; three symbols they reference from nvcore / nvmath are stubbed below: an abort hook and the encoder-only PCA fit).
// The engine hands such images to the Vulkan driver (format table src/engine/core/resource_manager.cpp:14-44); the shim decodes
// them on the host (helios_b200/shim/src/bc_decode.cpp), and tests/test_bc67.py holds that decoder to this one bit for bit.
//
//   ref_bc_tool 7|6u|6s blocks.bin out.bin     blocks.bin = n x 16 bytes
//   out.bin: BC7 -> n x 16 pixels x RGBA8; BC6H -> n x 16 pixels x 3 x uint16 (half-float bit patterns)
#include "bc7/tile.h"
#include "bc7/avpcl.h"
#include "bc6h/tile.h"
#include "bc6h/zoh.h"
#include "nvmath/Fitting.h"
#include "nvmath/Vector.inl"
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

int nvAbort(const char* exp, const char* file, int line, const char* func, const char* msg, ...)
{
    (void)exp, (void)file, (void)line, (void)func, (void)msg;
    return 0; // NV_ABORT_IGNORE: the decoders' bookkeeping assertions (bit-pointer checks) are not part of the result
}
namespace nv
{
namespace Fit
{
Vector3 computePrincipalComponent_EigenSolver(int, const Vector3*) { return Vector3(0.0f); } // encoder only
Vector4 computePrincipalComponent_EigenSolver(int, const Vector4*) { return Vector4(0.0f); }
} // namespace Fit
} // namespace nv

int main(int argc, char** argv)
{
    if (argc != 4)
    {
        std::fprintf(stderr, "usage: ref_bc_tool 7|6u|6s blocks.bin out.bin\n");
        return 2;
    }
    FILE* f = std::fopen(argv[2], "rb");
    if (!f) return 1;
    std::vector<char> in;
    char              buf[4096];
    size_t            n;
    while ((n = std::fread(buf, 1, sizeof(buf), f)) > 0) in.insert(in.end(), buf, buf + n);
    std::fclose(f);
    FILE* o = std::fopen(argv[3], "wb");
    if (!o) return 1;
    const size_t blocks = in.size() / 16;
    if (argv[1][0] == '7')
    {
        for (size_t b = 0; b < blocks; b++)
        {
            AVPCL::Tile tile(4, 4);
            const char* blk = in.data() + 16 * b;
            char        shifted[16];
            if ((blk[0] & 1) != 0)
            {
                // mode 0: this nvtt version's read_header (bc7/avpcl_mode0.cpp:242-246) has the getmode() call that consumes
                // the mode bit commented out, so AVPCL::decompress reads every field one bit early.  The rest of its mode-0
                // decoder is sound: hand it the block with the mode bit already removed.
                for (int k = 0; k < 16; k++) shifted[k] = (char)((((unsigned char)blk[k]) >> 1) | (k < 15 ? (((unsigned char)blk[k + 1]) & 1) << 7 : 0));
                AVPCL::decompress_mode0(shifted, tile);
            }
            else
                AVPCL::decompress(blk, tile);
            unsigned char px[64];
            for (int y = 0; y < 4; y++)
                for (int x = 0; x < 4; x++)
                {
                    const nv::Vector4 c = tile.data[y][x];
                    px[(y * 4 + x) * 4 + 0] = (unsigned char)c.x, px[(y * 4 + x) * 4 + 1] = (unsigned char)c.y, px[(y * 4 + x) * 4 + 2] = (unsigned char)c.z, px[(y * 4 + x) * 4 + 3] = (unsigned char)c.w;
                }
            std::fwrite(px, 1, 64, o);
        }
    }
    else
    {
        ZOH::Utils::FORMAT = argv[1][1] == 's' ? ZOH::SIGNED_F16 : ZOH::UNSIGNED_F16;
        for (size_t b = 0; b < blocks; b++)
        {
            ZOH::Tile tile(4, 4);
            ZOH::decompress(in.data() + 16 * b, tile);
            unsigned short px[48];
            for (int y = 0; y < 4; y++)
                for (int x = 0; x < 4; x++)
                {
                    px[(y * 4 + x) * 3 + 0] = ZOH::Tile::float2half(tile.data[y][x].x), px[(y * 4 + x) * 3 + 1] = ZOH::Tile::float2half(tile.data[y][x].y);
                    px[(y * 4 + x) * 3 + 2] = ZOH::Tile::float2half(tile.data[y][x].z);
                }
            std::fwrite(px, 2, 48, o);
        }
    }
    std::fclose(o);
    return 0;
}
