// bc_decode.cpp — see utility/bc_decode.h.
#include <utility/bc_decode.h>

namespace helios
{
namespace
{
    inline void rgb565(uint16_t c, int out[3])
    {
        const int r = (c >> 11) & 31, g = (c >> 5) & 63, b = c & 31;
        out[0] = (r << 3) | (r >> 2), out[1] = (g << 2) | (g >> 4), out[2] = (b << 3) | (b >> 2);
    }
    // 8-byte colour block -> 16 RGBA texels.  `punch_through`: the c0 <= c1 ordering selects the 3-colour +
    // transparent-black mode (BC1); BC2 / BC3 colour blocks always use the 4-colour mode.
    void color_block(const uint8_t* b, bool punch_through, uint8_t out[16][4])
    {
        const uint16_t c0 = (uint16_t)(b[0] | (b[1] << 8)), c1 = (uint16_t)(b[2] | (b[3] << 8));
        int            p[4][4];
        rgb565(c0, p[0]), rgb565(c1, p[1]);
        p[0][3] = p[1][3] = p[2][3] = p[3][3] = 255;
        if (c0 > c1 || !punch_through)
            for (int k = 0; k < 3; k++) p[2][k] = (2 * p[0][k] + p[1][k]) / 3, p[3][k] = (p[0][k] + 2 * p[1][k]) / 3;
        else
        {
            for (int k = 0; k < 3; k++) p[2][k] = (p[0][k] + p[1][k]) / 2, p[3][k] = 0;
            p[3][3] = 0;
        }
        const uint32_t idx = (uint32_t)b[4] | ((uint32_t)b[5] << 8) | ((uint32_t)b[6] << 16) | ((uint32_t)b[7] << 24);
        for (int i = 0; i < 16; i++)
        {
            const int s = (idx >> (2 * i)) & 3;
            for (int k = 0; k < 4; k++) out[i][k] = (uint8_t)p[s][k];
        }
    }
    // 8-byte interpolated single-channel block (BC3 alpha, BC4, BC5 halves) -> 16 values
    void alpha_block(const uint8_t* b, uint8_t out[16])
    {
        int a[8];
        a[0] = b[0], a[1] = b[1];
        if (a[0] > a[1])
            for (int i = 1; i < 7; i++) a[i + 1] = ((7 - i) * a[0] + i * a[1]) / 7;
        else
        {
            for (int i = 1; i < 5; i++) a[i + 1] = ((5 - i) * a[0] + i * a[1]) / 5;
            a[6] = 0, a[7] = 255;
        }
        uint64_t bits = 0;
        for (int i = 0; i < 6; i++) bits |= (uint64_t)b[2 + i] << (8 * i);
        for (int i = 0; i < 16; i++) out[i] = (uint8_t)a[(bits >> (3 * i)) & 7];
    }
} // namespace

uint32_t bc_block_bytes(int compression)
{
    switch (compression)
    {
        case 1:
        case 2:
        case 6: return 8;
        case 3:
        case 4:
        case 5:
        case 7:
        case 8:
        case 9: return 16;
        default: return 0;
    }
}

bool decode_bc(int compression, const uint8_t* blocks, size_t n_bytes, uint32_t width, uint32_t height, std::vector<uint8_t>& out)
{
    const uint32_t bb = bc_block_bytes(compression);
    if (bb == 0 || compression == 8 || compression == 9 || width == 0 || height == 0) return false;
    const uint32_t bw = (width + 3) / 4, bh = (height + 3) / 4;
    if ((uint64_t)bw * bh * bb > n_bytes) return false;
    out.assign((size_t)width * height * 4, 0);
    for (uint32_t by = 0; by < bh; by++)
        for (uint32_t bx = 0; bx < bw; bx++)
        {
            const uint8_t* b = blocks + ((size_t)by * bw + bx) * bb;
            uint8_t        px[16][4];
            if (compression == 1 || compression == 2)
                color_block(b, true, px);
            else if (compression == 3)
            {
                color_block(b + 8, false, px);
                for (int i = 0; i < 16; i++)
                {
                    const int a4 = (b[i / 2] >> (4 * (i & 1))) & 15;
                    px[i][3]     = (uint8_t)(a4 * 17);
                }
            }
            else if (compression == 4 || compression == 5)
            {
                uint8_t a[16];
                color_block(b + 8, false, px);
                alpha_block(b, a);
                for (int i = 0; i < 16; i++) px[i][3] = a[i];
            }
            else
            {
                uint8_t r[16], g[16];
                alpha_block(b, r);
                if (compression == 7) alpha_block(b + 8, g);
                for (int i = 0; i < 16; i++) px[i][0] = r[i], px[i][1] = compression == 7 ? g[i] : 0, px[i][2] = 0, px[i][3] = 255;
            }
            for (int i = 0; i < 16; i++)
            {
                const uint32_t x = bx * 4 + (uint32_t)(i & 3), y = by * 4 + (uint32_t)(i >> 2);
                if (x >= width || y >= height) continue;
                uint8_t* o = &out[((size_t)y * width + x) * 4];
                o[0] = px[i][0], o[1] = px[i][1], o[2] = px[i][2], o[3] = px[i][3];
            }
        }
    return true;
}
} // namespace helios
