// png_writer.cpp — see utility/png_writer.h.  Written from the PNG (ISO/IEC 15948) and DEFLATE / zlib
// (RFC 1951 / 1950) specifications: Sub-filtered scanlines, greedy hash-chain-free LZ77 (one candidate per
// 3-byte hash, 32 KiB window), fixed Huffman codes, CRC-32 per chunk, Adler-32 over the filtered stream.
#include <utility/png_writer.h>
#include <cstdio>
#include <cstring>

namespace helios
{
namespace
{
    struct Crc32
    {
        uint32_t table[256];
        Crc32()
        {
            for (uint32_t n = 0; n < 256; n++)
            {
                uint32_t c = n;
                for (int k = 0; k < 8; k++) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
                table[n] = c;
            }
        }
        uint32_t operator()(const uint8_t* p, size_t n, uint32_t crc = 0) const
        {
            crc = ~crc;
            for (size_t i = 0; i < n; i++) crc = table[(crc ^ p[i]) & 0xFFu] ^ (crc >> 8);
            return ~crc;
        }
    };

    uint32_t adler32(const uint8_t* p, size_t n)
    {
        uint32_t a = 1, b = 0;
        while (n)
        {
            const size_t k = n < 5552 ? n : 5552; // largest run before the sums can overflow 32 bits
            for (size_t i = 0; i < k; i++) a += p[i], b += a;
            a %= 65521u, b %= 65521u, p += k, n -= k;
        }
        return (b << 16) | a;
    }

    void put_be32(std::vector<uint8_t>& o, uint32_t v)
    {
        o.push_back(uint8_t(v >> 24)), o.push_back(uint8_t(v >> 16)), o.push_back(uint8_t(v >> 8)), o.push_back(uint8_t(v));
    }

    void put_chunk(std::vector<uint8_t>& o, const char type[4], const uint8_t* data, size_t n)
    {
        static const Crc32 crc;
        put_be32(o, (uint32_t)n);
        const size_t start = o.size();
        o.insert(o.end(), type, type + 4);
        if (n) o.insert(o.end(), data, data + n);
        put_be32(o, crc(o.data() + start, n + 4));
    }

    // LSB-first bit writer (RFC 1951 3.1.1); Huffman codes are sent most-significant code bit first
    struct BitWriter
    {
        std::vector<uint8_t>& out;
        uint64_t              acc = 0;
        int                   n   = 0;
        explicit BitWriter(std::vector<uint8_t>& o) : out(o) {}
        void bits(uint32_t v, int count)
        {
            acc |= (uint64_t)v << n, n += count;
            while (n >= 8) out.push_back(uint8_t(acc)), acc >>= 8, n -= 8;
        }
        void code(uint32_t c, int len)
        {
            uint32_t r = 0;
            for (int i = 0; i < len; i++) r |= ((c >> i) & 1u) << (len - 1 - i);
            bits(r, len);
        }
        void flush()
        {
            if (n) out.push_back(uint8_t(acc)), acc = 0, n = 0;
        }
    };

    // fixed Huffman literal/length alphabet, RFC 1951 3.2.6
    void put_symbol(BitWriter& w, uint32_t s)
    {
        if (s < 144)
            w.code(0x30u + s, 8);
        else if (s < 256)
            w.code(0x190u + (s - 144), 9);
        else if (s < 280)
            w.code(s - 256, 7);
        else
            w.code(0xC0u + (s - 280), 8);
    }

    const uint16_t LEN_BASE[29]  = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    const uint8_t  LEN_EXTRA[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    const uint16_t DIST_BASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
    const uint8_t  DIST_EXTRA[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

    void put_match(BitWriter& w, uint32_t length, uint32_t distance)
    {
        int li = 28;
        while (LEN_BASE[li] > length) li--;
        put_symbol(w, 257u + (uint32_t)li);
        if (LEN_EXTRA[li]) w.bits(length - LEN_BASE[li], LEN_EXTRA[li]);
        int di = 29;
        while (DIST_BASE[di] > distance) di--;
        w.code((uint32_t)di, 5);
        if (DIST_EXTRA[di]) w.bits(distance - DIST_BASE[di], DIST_EXTRA[di]);
    }

    // zlib stream: one final fixed-Huffman block
    void deflate_fixed(std::vector<uint8_t>& out, const uint8_t* src, size_t n)
    {
        out.push_back(0x78), out.push_back(0x01); // CM 8, 32 KiB window; FCHECK makes the pair a multiple of 31
        BitWriter w(out);
        w.bits(1, 1), w.bits(1, 2); // BFINAL, BTYPE = 01
        const uint32_t        HASH_BITS = 15;
        std::vector<uint32_t> head(1u << HASH_BITS, 0xFFFFFFFFu);
        auto hash3 = [&](size_t i) { return ((uint32_t(src[i]) | uint32_t(src[i + 1]) << 8 | uint32_t(src[i + 2]) << 16) * 0x9E3779B1u) >> (32 - HASH_BITS); };
        size_t i = 0;
        while (i < n)
        {
            uint32_t best_len = 0, best_dist = 0;
            if (i + 3 <= n)
            {
                const uint32_t h = hash3(i);
                const uint32_t c = head[h];
                head[h]          = (uint32_t)i;
                if (c != 0xFFFFFFFFu && i - c <= 32768)
                {
                    const size_t lim = n - i < 258 ? n - i : 258;
                    size_t       l   = 0;
                    while (l < lim && src[c + l] == src[i + l]) l++;
                    if (l >= 3) best_len = (uint32_t)l, best_dist = (uint32_t)(i - c);
                }
            }
            if (best_len)
            {
                put_match(w, best_len, best_dist);
                // index the skipped positions sparsely (every other byte): enough for image rows
                for (size_t k = i + 2; k + 3 <= n && k < i + best_len; k += 2) head[hash3(k)] = (uint32_t)k;
                i += best_len;
            }
            else
                put_symbol(w, src[i++]);
        }
        put_symbol(w, 256);
        w.flush();
        put_be32(out, adler32(src, n));
    }
} // namespace

std::vector<uint8_t> encode_png_rgba8(uint32_t width, uint32_t height, const uint8_t* rgba, size_t row_stride_bytes)
{
    std::vector<uint8_t> png = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    std::vector<uint8_t> ihdr;
    put_be32(ihdr, width), put_be32(ihdr, height);
    ihdr.push_back(8), ihdr.push_back(6), ihdr.push_back(0), ihdr.push_back(0), ihdr.push_back(0);
    put_chunk(png, "IHDR", ihdr.data(), ihdr.size());
    // filter type 1 (Sub): each byte minus the byte one pixel to the left
    const size_t         row = (size_t)width * 4;
    std::vector<uint8_t> filtered((row + 1) * height);
    for (uint32_t y = 0; y < height; y++)
    {
        const uint8_t* s = rgba + (size_t)y * row_stride_bytes;
        uint8_t*       d = filtered.data() + (size_t)y * (row + 1);
        *d++             = 1;
        for (size_t x = 0; x < row; x++) d[x] = uint8_t(s[x] - (x >= 4 ? s[x - 4] : 0));
    }
    std::vector<uint8_t> z;
    z.reserve(filtered.size() / 2 + 64);
    deflate_fixed(z, filtered.data(), filtered.size());
    put_chunk(png, "IDAT", z.data(), z.size());
    put_chunk(png, "IEND", nullptr, 0);
    return png;
}

bool write_png_rgba8(const std::string& path, uint32_t width, uint32_t height, const uint8_t* rgba, size_t row_stride_bytes)
{
    if (width == 0 || height == 0 || rgba == nullptr) return false;
    const std::vector<uint8_t> png = encode_png_rgba8(width, height, rgba, row_stride_bytes);
    FILE*                      f   = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    const bool ok = std::fwrite(png.data(), 1, png.size(), f) == png.size();
    return (std::fclose(f) == 0) && ok;
}
} // namespace helios

extern "C" __attribute__((visibility("default"))) int helios_write_png_rgba8(const char* path, uint32_t width, uint32_t height, const uint8_t* rgba, uint32_t row_stride_bytes)
{
    return helios::write_png_rgba8(path ? path : "", width, height, rgba, row_stride_bytes) ? 0 : 1;
}
