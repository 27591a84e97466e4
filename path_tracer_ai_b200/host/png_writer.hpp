// png_writer.hpp — 8-bit RGB PNG encoder over zlib, standing in for stb_image_write (an absent
// dependency of the reference, src/renderer.cpp:3,19).  Rows are written in the order given.
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>
#include <zlib.h>

namespace b2pt {

inline bool writePngRGB8(const std::string& path, int width, int height, const uint8_t* rgb) {
    std::vector<uint8_t> raw;
    raw.reserve((size_t)height * (1 + (size_t)width * 3));
    for (int y = 0; y < height; ++y) {
        raw.push_back(0);   // filter type: none
        raw.insert(raw.end(), rgb + (size_t)y * width * 3, rgb + (size_t)(y + 1) * width * 3);
    }
    uLongf clen = compressBound((uLong)raw.size());
    std::vector<uint8_t> comp(clen);
    if (compress2(comp.data(), &clen, raw.data(), (uLong)raw.size(), 6) != Z_OK) return false;
    comp.resize(clen);

    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    auto be32 = [](uint8_t* p, uint32_t v) { p[0] = v >> 24; p[1] = v >> 16; p[2] = v >> 8; p[3] = v; };
    auto chunk = [&](const char* type, const uint8_t* data, uint32_t len) {
        uint8_t hdr[8];
        be32(hdr, len);
        hdr[4] = type[0]; hdr[5] = type[1]; hdr[6] = type[2]; hdr[7] = type[3];
        std::fwrite(hdr, 1, 8, f);
        if (len) std::fwrite(data, 1, len, f);
        uLong crc = crc32(0L, hdr + 4, 4);
        if (len) crc = crc32(crc, data, len);
        uint8_t c[4];
        be32(c, (uint32_t)crc);
        std::fwrite(c, 1, 4, f);
    };
    const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    std::fwrite(sig, 1, 8, f);
    uint8_t ihdr[13];
    be32(ihdr, (uint32_t)width);
    be32(ihdr + 4, (uint32_t)height);
    ihdr[8] = 8; ihdr[9] = 2; ihdr[10] = 0; ihdr[11] = 0; ihdr[12] = 0;   // 8-bit, truecolour
    chunk("IHDR", ihdr, 13);
    chunk("IDAT", comp.data(), (uint32_t)comp.size());
    chunk("IEND", nullptr, 0);
    return std::fclose(f) == 0;
}

// Portable float map ("PF", little-endian, 3 channels): the linear float framebuffer without tone mapping.  A PFM stores
// its rows bottom-to-top, which is the order of the reference's frameBuffer (renderer.hpp:81: row 0 = bottom of the view),
// so the buffer is written as it is and any PFM viewer shows the image upright.
inline bool writePfmRGB(const std::string& path, int width, int height, const float* rgb) {
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    std::fprintf(f, "PF\n%d %d\n-1.0\n", width, height);
    const size_t n = (size_t)width * height * 3;
    bool ok = std::fwrite(rgb, sizeof(float), n, f) == n;
    return (std::fclose(f) == 0) && ok;
}

}  // namespace b2pt
