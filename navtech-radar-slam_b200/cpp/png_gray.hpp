// png_gray.hpp -- the one thing cv::imread(path, cv::IMREAD_GRAYSCALE) does for load_radar
// (outlier-robust-radar-odometry/src/radar_utils.cpp:107): decode a grey PNG into 8-bit rows.
// OpenCV is not in this image, zlib is; Navtech "oxford-form" polar scans are 8-bit grey, non-interlaced PNGs, so a
// chunk walk + inflate + the five PNG row filters is the whole job.  Grey and grey+alpha of every bit depth decode to
// what imread returns (16-bit: the high byte, like png_set_strip_16; 1/2/4-bit: expanded to 0..255; alpha dropped).
// Colour and palette images are refused: imread would mix channels with libpng's own weights, nothing on this path
// produces them.  Host-side input handling only -- no arithmetic of the hot path lives here.
#ifndef NRS_PNG_GRAY_HPP
#define NRS_PNG_GRAY_HPP
#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace nrs {

struct GrayImage {
    int rows = 0, cols = 0;
    std::vector<uint8_t> data;                     // rows x cols, row-major, step == cols
};

namespace png_detail {
inline uint32_t be32(const uint8_t *p) { return (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3]; }
inline int paeth(int a, int b, int c)
{
    const int p = a + b - c, pa = p > a ? p - a : a - p, pb = p > b ? p - b : b - p, pc = p > c ? p - c : c - p;
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}
}  // namespace png_detail

inline GrayImage decode_png_gray(const uint8_t *buf, size_t len, const std::string &what = "png")
{
    using namespace png_detail;
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    auto fail = [&](const std::string &m) -> void { throw std::runtime_error(what + ": " + m); };
    if (len < 8 || std::memcmp(buf, sig, 8) != 0) fail("not a PNG file");
    uint32_t W = 0, H = 0;
    int depth = 0, ctype = -1;
    bool have_ihdr = false, have_iend = false;
    std::vector<uint8_t> z;
    size_t pos = 8;
    while (pos + 12 <= len && !have_iend) {
        const uint32_t n = be32(buf + pos);
        const uint8_t *type = buf + pos + 4, *body = buf + pos + 8;
        if (n > len - pos - 12) fail("truncated chunk");
        const uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), type, n + 4);
        if (crc != be32(body + n)) fail("chunk CRC mismatch");
        if (!std::memcmp(type, "IHDR", 4)) {
            if (n != 13) fail("bad IHDR");
            W = be32(body); H = be32(body + 4); depth = body[8]; ctype = body[9];
            if (body[10] != 0 || body[11] != 0) fail("unknown compression / filter method");
            if (body[12] != 0) fail("interlaced PNGs are not supported");
            if (W == 0 || H == 0 || W > (1u << 20) || H > (1u << 20)) fail("unreasonable dimensions");
            if (ctype != 0 && ctype != 4) fail("only grey / grey+alpha PNGs are decoded (colour type " + std::to_string(ctype) + ")");
            if (!(depth == 8 || depth == 16 || (ctype == 0 && (depth == 1 || depth == 2 || depth == 4)))) fail("bad bit depth");
            have_ihdr = true;
        } else if (!std::memcmp(type, "IDAT", 4)) {
            if (!have_ihdr) fail("IDAT before IHDR");
            z.insert(z.end(), body, body + n);
        } else if (!std::memcmp(type, "IEND", 4)) {
            have_iend = true;
        } else if (!(type[0] & 0x20)) {
            if (std::memcmp(type, "PLTE", 4)) fail("unknown critical chunk");
        }
        pos += 12 + (size_t)n;
    }
    if (!have_ihdr || !have_iend) fail("missing IHDR / IEND");
    const int channels = ctype == 4 ? 2 : 1;
    const size_t bits_pp = (size_t)depth * channels;
    const size_t stride = ((size_t)W * bits_pp + 7) / 8;           // bytes of one filtered scanline (without the type byte)
    const size_t bpp = bits_pp >= 8 ? bits_pp / 8 : 1;             // the filters' "corresponding byte" distance
    std::vector<uint8_t> raw((stride + 1) * (size_t)H);
    uLongf out_len = (uLongf)raw.size();
    const int zrc = uncompress(raw.data(), &out_len, z.data(), (uLong)z.size());
    if (zrc != Z_OK || out_len != raw.size()) fail("inflate failed");
    // undo the row filters in place (type byte at the head of every row)
    std::vector<uint8_t> zero(stride, 0);
    for (uint32_t y = 0; y < H; ++y) {
        uint8_t *row = raw.data() + (stride + 1) * (size_t)y + 1;
        const uint8_t *up = y ? row - (stride + 1) : zero.data();
        switch (row[-1]) {
        case 0: break;
        case 1: for (size_t x = bpp; x < stride; ++x) row[x] = (uint8_t)(row[x] + row[x - bpp]); break;
        case 2: for (size_t x = 0; x < stride; ++x) row[x] = (uint8_t)(row[x] + up[x]); break;
        case 3:
            for (size_t x = 0; x < stride; ++x)
                row[x] = (uint8_t)(row[x] + (((x >= bpp ? row[x - bpp] : 0) + up[x]) >> 1));
            break;
        case 4:
            for (size_t x = 0; x < stride; ++x)
                row[x] = (uint8_t)(row[x] + paeth(x >= bpp ? row[x - bpp] : 0, up[x], x >= bpp ? up[x - bpp] : 0));
            break;
        default: fail("unknown row filter");
        }
    }
    GrayImage im;
    im.rows = (int)H; im.cols = (int)W;
    im.data.resize((size_t)W * H);
    for (uint32_t y = 0; y < H; ++y) {
        const uint8_t *row = raw.data() + (stride + 1) * (size_t)y + 1;
        uint8_t *dst = im.data.data() + (size_t)W * y;
        if (depth == 8) {
            for (uint32_t x = 0; x < W; ++x) dst[x] = row[(size_t)x * channels];
        } else if (depth == 16) {
            for (uint32_t x = 0; x < W; ++x) dst[x] = row[(size_t)x * channels * 2];          // high byte
        } else {
            const int maxv = (1 << depth) - 1;
            for (uint32_t x = 0; x < W; ++x) {
                const size_t bit = (size_t)x * depth;
                const int v = (row[bit >> 3] >> (8 - depth - (int)(bit & 7))) & maxv;
                dst[x] = (uint8_t)(v * 255 / maxv);
            }
        }
    }
    return im;
}

inline GrayImage read_png_gray(const std::string &path)
{
    std::FILE *f = std::fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("cannot open " + path);
    std::vector<uint8_t> buf;
    uint8_t chunk[1 << 16];
    size_t n;
    while ((n = std::fread(chunk, 1, sizeof(chunk), f)) > 0) buf.insert(buf.end(), chunk, chunk + n);
    std::fclose(f);
    return decode_png_gray(buf.data(), buf.size(), path);
}

}  // namespace nrs
#endif
