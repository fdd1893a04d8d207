// png_writer.hpp -- RGB8 PNG encoder over zlib (what massiv-io's `writeArray PNG` produces for
// Raytracer.writeImg, src/Raytracer.hs:29-32: 8-bit RGB, non-interlaced).
#pragma once

#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

namespace pngw {

inline void put32(std::vector<uint8_t> &v, uint32_t x)
{
    v.push_back(uint8_t(x >> 24)); v.push_back(uint8_t(x >> 16)); v.push_back(uint8_t(x >> 8)); v.push_back(uint8_t(x));
}

inline void chunk(std::vector<uint8_t> &out, const char type[4], const uint8_t *data, size_t n)
{
    put32(out, (uint32_t)n);
    const size_t start = out.size();
    out.insert(out.end(), type, type + 4);
    if (n) out.insert(out.end(), data, data + n);
    put32(out, (uint32_t)crc32(0L, out.data() + start, (uInt)(n + 4)));
}

// rgb: height rows of width*3 bytes.  Returns "" or an error message.
inline std::string write_rgb8(const std::string &path, const uint8_t *rgb, int width, int height, int level = 6)
{
    if (width <= 0 || height <= 0) return "bad image size";
    const size_t stride = (size_t)width * 3;
    std::vector<uint8_t> raw((stride + 1) * (size_t)height);
    for (int y = 0; y < height; y++) {
        raw[(stride + 1) * y] = 0;  // filter type None
        std::copy(rgb + stride * y, rgb + stride * (y + 1), raw.begin() + (stride + 1) * y + 1);
    }
    uLongf clen = compressBound((uLong)raw.size());
    std::vector<uint8_t> comp(clen);
    if (compress2(comp.data(), &clen, raw.data(), (uLong)raw.size(), level) != Z_OK) return "zlib compress2 failed";
    std::vector<uint8_t> out = { 0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a };
    std::vector<uint8_t> ihdr;
    put32(ihdr, (uint32_t)width); put32(ihdr, (uint32_t)height);
    ihdr.push_back(8);  // bit depth
    ihdr.push_back(2);  // colour type: truecolour
    ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
    chunk(out, "IHDR", ihdr.data(), ihdr.size());
    chunk(out, "IDAT", comp.data(), clen);
    chunk(out, "IEND", nullptr, 0);
    FILE *f = std::fopen(path.c_str(), "wb");
    if (!f) return "cannot open " + path + " for writing";
    const bool ok = std::fwrite(out.data(), 1, out.size(), f) == out.size();
    std::fclose(f);
    return ok ? "" : "short write to " + path;
}

}  // namespace pngw
