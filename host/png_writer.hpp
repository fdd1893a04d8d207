// png_writer.hpp -- RGB8 PNG encoder over zlib (what massiv-io's `writeArray PNG` produces for
// Raytracer.writeImg, src/Raytracer.hs:29-32: 8-bit RGB, non-interlaced).
#pragma once

#include <sched.h>
#include <zlib.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <string>
#include <thread>
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

inline std::string write_file(const std::string &path, const std::vector<uint8_t> &out)
{
    FILE *f = std::fopen(path.c_str(), "wb");
    if (!f) return "cannot open " + path + " for writing";
    const bool ok = std::fwrite(out.data(), 1, out.size(), f) == out.size();
    std::fclose(f);
    return ok ? "" : "short write to " + path;
}

inline void header(std::vector<uint8_t> &out, int width, int height)
{
    const uint8_t sig[8] = { 0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a };
    out.assign(sig, sig + 8);
    std::vector<uint8_t> ihdr;
    put32(ihdr, (uint32_t)width); put32(ihdr, (uint32_t)height);
    ihdr.push_back(8);  // bit depth
    ihdr.push_back(2);  // colour type: truecolour
    ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
    chunk(out, "IHDR", ihdr.data(), ihdr.size());
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
    std::vector<uint8_t> out;
    header(out, width, height);
    chunk(out, "IDAT", comp.data(), clen);
    chunk(out, "IEND", nullptr, 0);
    return write_file(path, out);
}

// Same image, deflated by `threads` workers over bands of rows (the pigz construction: every
// band is an independent raw-deflate stream ended with a sync flush, the last one with the final
// block; one zlib header in front, the combined Adler-32 behind).  PNG encoding is the slow part
// of writeImg for large frames; this makes it scale with host cores.
// CPUs this process may actually use: the scheduler affinity and the cgroup v2 quota, not what the machine has
// (hardware_concurrency() is 128+ on a GPU host whose container may use 16: one deflate thread per "core" then
// means an order of magnitude of oversubscription and quota throttling).
inline int usable_cpus()
{
    int n = (int)std::max(1u, std::thread::hardware_concurrency());
#if defined(__linux__)
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof set, &set) == 0) n = std::min(n, std::max(1, CPU_COUNT(&set)));
    if (FILE *f = std::fopen("/sys/fs/cgroup/cpu.max", "r")) {
        char q[32] = "";
        long long period = 0;
        if (std::fscanf(f, "%31s %lld", q, &period) == 2 && std::strcmp(q, "max") != 0 && period > 0)
            n = std::min(n, std::max(1, (int)((std::atoll(q) + period / 2) / period)));
        std::fclose(f);
    }
#endif
    return n;
}

inline std::string write_rgb8_parallel(const std::string &path, const uint8_t *rgb, int width, int height,
                                       int threads = 0, int level = 6)
{
    if (width <= 0 || height <= 0) return "bad image size";
    if (threads <= 0) threads = usable_cpus();
    const int bands = std::max(1, std::min(threads, height / 16));
    if (bands == 1) return write_rgb8(path, rgb, width, height, level);
    const size_t stride = (size_t)width * 3;
    std::vector<std::vector<uint8_t>> comp(bands);
    std::vector<uLong> adler(bands), rawlen(bands);
    std::vector<int> status(bands, Z_OK);
    auto work = [&](int b) {
        const int y0 = (int)((long long)height * b / bands), y1 = (int)((long long)height * (b + 1) / bands);
        std::vector<uint8_t> raw((stride + 1) * (size_t)(y1 - y0));
        for (int y = y0; y < y1; y++) {
            raw[(stride + 1) * (y - y0)] = 0;
            std::copy(rgb + stride * y, rgb + stride * (y + 1), raw.begin() + (stride + 1) * (y - y0) + 1);
        }
        rawlen[b] = (uLong)raw.size();
        adler[b] = adler32(adler32(0L, Z_NULL, 0), raw.data(), (uInt)raw.size());
        z_stream zs{};
        if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) { status[b] = Z_STREAM_ERROR; return; }
        comp[b].resize(deflateBound(&zs, (uLong)raw.size()) + 16);
        zs.next_in = raw.data(); zs.avail_in = (uInt)raw.size();
        zs.next_out = comp[b].data(); zs.avail_out = (uInt)comp[b].size();
        const int rc = deflate(&zs, b == bands - 1 ? Z_FINISH : Z_SYNC_FLUSH);
        if ((b == bands - 1 && rc != Z_STREAM_END) || (b != bands - 1 && rc != Z_OK) || zs.avail_in != 0) status[b] = Z_BUF_ERROR;
        comp[b].resize(zs.total_out);
        deflateEnd(&zs);
    };
    std::vector<std::thread> pool;
    for (int b = 1; b < bands; b++) pool.emplace_back(work, b);
    work(0);
    for (auto &t : pool) t.join();
    for (int b = 0; b < bands; b++)
        if (status[b] != Z_OK) return "zlib deflate failed";
    std::vector<uint8_t> z = { 0x78, 0x9c };
    uLong ad = adler[0];
    for (int b = 0; b < bands; b++) {
        z.insert(z.end(), comp[b].begin(), comp[b].end());
        if (b > 0) ad = adler32_combine(ad, adler[b], (z_off_t)rawlen[b]);
    }
    put32(z, (uint32_t)ad);
    std::vector<uint8_t> out;
    header(out, width, height);
    chunk(out, "IDAT", z.data(), z.size());
    chunk(out, "IEND", nullptr, 0);
    return write_file(path, out);
}

}  // namespace pngw
