// png_decode.cpp -- PNG -> RGBA8 for .scene textures (inflate by the system zlib, everything else here).
// Conversions that a decoder is free to choose follow stb_image, which the reference uses (scene_shift.cpp:39):
// 16-bit samples keep their high byte, 1/2/4-bit grey is scaled by 255/(2^d - 1), a tRNS colour key gives alpha 0,
// gAMA/sRGB/iCCP are ignored.  Adam7 interlacing is supported.
#include <zlib.h>

#include <cstdlib>
#include <cstring>

#include "image_io.hpp"

namespace spchost {
namespace {

inline uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
inline int paeth(int a, int b, int c) {
    const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

// undo the per-row filters of one (sub-)image in place; rows are `stride` bytes preceded by a filter byte
bool unfilter(uint8_t* raw, size_t avail, int rows, size_t stride, int bpp, std::vector<uint8_t>& out) {
    if ((stride + 1) * (size_t)rows > avail) return false;
    out.assign(stride * rows, 0);
    for (int y = 0; y < rows; y++) {
        const uint8_t* in = raw + (stride + 1) * y;
        const int f = *in++;
        uint8_t* cur = out.data() + stride * y;
        const uint8_t* up = y ? cur - stride : nullptr;
        for (size_t i = 0; i < stride; i++) {
            const int a = i >= (size_t)bpp ? cur[i - bpp] : 0;
            const int b = up ? up[i] : 0;
            const int c = (up && i >= (size_t)bpp) ? up[i - bpp] : 0;
            int v = in[i];
            switch (f) {
                case 0: break;
                case 1: v += a; break;
                case 2: v += b; break;
                case 3: v += (a + b) >> 1; break;
                case 4: v += paeth(a, b, c); break;
                default: return false;
            }
            cur[i] = (uint8_t)v;
        }
    }
    return true;
}

}  // namespace

bool decode_png_rgba8(const uint8_t* data, size_t size, ImageRGBA8& img, std::string& err) {
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (size < 8 || memcmp(data, sig, 8) != 0) {
        err = "not a PNG file";
        return false;
    }
    int W = 0, H = 0, depth = 0, ctype = 0, interlace = 0;
    uint8_t pal[256][4];
    int pal_n = 0;
    bool has_key = false;
    uint16_t key[3] = {0, 0, 0};
    std::vector<uint8_t> z;
    for (size_t p = 8; p + 12 <= size;) {
        const uint32_t len = be32(data + p);
        const uint8_t* tag = data + p + 4;
        const uint8_t* body = data + p + 8;
        if (p + 12 + (size_t)len > size) {
            err = "truncated PNG chunk";
            return false;
        }
        if (!memcmp(tag, "IHDR", 4) && len >= 13) {
            W = (int)be32(body);
            H = (int)be32(body + 4);
            depth = body[8];
            ctype = body[9];
            interlace = body[12];
        } else if (!memcmp(tag, "PLTE", 4)) {
            pal_n = (int)(len / 3);
            if (pal_n > 256) pal_n = 256;
            for (int i = 0; i < pal_n; i++) {
                pal[i][0] = body[3 * i];
                pal[i][1] = body[3 * i + 1];
                pal[i][2] = body[3 * i + 2];
                pal[i][3] = 255;
            }
        } else if (!memcmp(tag, "tRNS", 4)) {
            if (ctype == 3) {
                for (uint32_t i = 0; i < len && (int)i < pal_n; i++) pal[i][3] = body[i];
            } else if (ctype == 0 && len >= 2) {
                has_key = true;
                key[0] = (uint16_t)((body[0] << 8) | body[1]);
            } else if (ctype == 2 && len >= 6) {
                has_key = true;
                for (int k = 0; k < 3; k++) key[k] = (uint16_t)((body[2 * k] << 8) | body[2 * k + 1]);
            }
        } else if (!memcmp(tag, "IDAT", 4)) {
            z.insert(z.end(), body, body + len);
        } else if (!memcmp(tag, "IEND", 4)) {
            break;
        }
        p += 12 + (size_t)len;
    }
    static const int chan_of[7] = {1, 0, 3, 1, 2, 0, 4};
    if (W <= 0 || H <= 0 || ctype > 6 || chan_of[ctype] == 0 || !(depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16)) {
        err = "bad PNG header";
        return false;
    }
    // texture files are untrusted: a header may not ask for more memory than its data can fill (deflate expands at most ~1032 : 1)
    // nor for an image beyond 2^28 pixels (1 GiB of RGBA8)
    if ((uint64_t)W * (uint64_t)H > (1ull << 28) || (uint64_t)W * (uint64_t)H / 8 > (uint64_t)z.size() * 1040 + 1024) {
        err = "PNG header does not match its data (or the image exceeds 2^28 pixels)";
        return false;
    }
    const int chan = chan_of[ctype];
    const int bits_pp = chan * depth;
    const int bpp = bits_pp >= 8 ? bits_pp / 8 : 1;

    // inflate everything (upper bound: every pass adds at most one filter byte per row and 7 bits of padding)
    std::vector<uint8_t> raw(((size_t)W * bits_pp / 8 + 8) * (size_t)H + (size_t)H * 8 + 64);
    uLongf raw_len = (uLongf)raw.size();
    const int zr = uncompress(raw.data(), &raw_len, z.data(), (uLong)z.size());
    if (zr != Z_OK) {
        err = "PNG inflate failed";
        return false;
    }

    img.width = W;
    img.height = H;
    img.rgba.assign((size_t)W * H * 4, 255);
    static const int gscale[9] = {0, 0xff, 0x55, 0, 0x11, 0, 0, 0, 0x01};

    auto emit = [&](const uint8_t* row, int i, int X, int Y) {   // sample i of an unfiltered row -> pixel (X, Y)
        uint16_t s[4] = {0, 0, 0, 0};
        if (depth == 16) {
            for (int k = 0; k < chan; k++) s[k] = (uint16_t)((row[(i * chan + k) * 2] << 8) | row[(i * chan + k) * 2 + 1]);
        } else if (depth == 8) {
            for (int k = 0; k < chan; k++) s[k] = row[i * chan + k];
        } else {
            const int per = 8 / depth, sh = (per - 1 - i % per) * depth;
            s[0] = (uint16_t)((row[i / per] >> sh) & ((1 << depth) - 1));
        }
        uint8_t* o = img.rgba.data() + ((size_t)Y * W + X) * 4;
        bool keyed = false;
        if (has_key) {
            keyed = true;
            for (int k = 0; k < (ctype == 0 ? 1 : 3); k++) keyed = keyed && s[k] == key[k];
        }
        auto to8 = [&](uint16_t v) -> uint8_t { return depth == 16 ? (uint8_t)(v >> 8) : (uint8_t)v; };
        switch (ctype) {
            case 0: {
                const uint8_t g = depth < 8 ? (uint8_t)(s[0] * gscale[depth]) : to8(s[0]);
                o[0] = o[1] = o[2] = g;
                o[3] = keyed ? 0 : 255;
                break;
            }
            case 2:
                o[0] = to8(s[0]); o[1] = to8(s[1]); o[2] = to8(s[2]);
                o[3] = keyed ? 0 : 255;
                break;
            case 3: {
                const int k = s[0] < pal_n ? s[0] : 0;
                memcpy(o, pal[k], 4);
                break;
            }
            case 4:
                o[0] = o[1] = o[2] = to8(s[0]);
                o[3] = to8(s[1]);
                break;
            case 6:
                o[0] = to8(s[0]); o[1] = to8(s[1]); o[2] = to8(s[2]); o[3] = to8(s[3]);
                break;
        }
    };

    std::vector<uint8_t> pix;
    if (!interlace) {
        const size_t stride = ((size_t)W * bits_pp + 7) / 8;
        if (!unfilter(raw.data(), raw_len, H, stride, bpp, pix)) {
            err = "corrupt PNG data";
            return false;
        }
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++) emit(pix.data() + stride * y, x, x, y);
        return true;
    }
    static const int xo[7] = {0, 4, 0, 2, 0, 1, 0}, yo[7] = {0, 0, 4, 0, 2, 0, 1}, xs[7] = {8, 8, 4, 4, 2, 2, 1}, ys[7] = {8, 8, 8, 4, 4, 2, 2};
    size_t off = 0;
    for (int p = 0; p < 7; p++) {
        const int pw = (W - xo[p] + xs[p] - 1) / xs[p], ph = (H - yo[p] + ys[p] - 1) / ys[p];
        if (pw <= 0 || ph <= 0) continue;
        const size_t stride = ((size_t)pw * bits_pp + 7) / 8;
        if (off > raw_len || !unfilter(raw.data() + off, raw_len - off, ph, stride, bpp, pix)) {
            err = "corrupt PNG data";
            return false;
        }
        for (int y = 0; y < ph; y++)
            for (int x = 0; x < pw; x++) emit(pix.data() + stride * y, x, xo[p] + x * xs[p], yo[p] + y * ys[p]);
        off += (stride + 1) * ph;
    }
    return true;
}

// ---- writer: IHDR / IDAT / IEND, colour type 2 (RGB), 8 bit, filter method 0 on every row ---------------------------------------
namespace {
void png_chunk(FILE* f, const char type[4], const uint8_t* data, size_t n) {
    const uint8_t len[4] = {(uint8_t)(n >> 24), (uint8_t)(n >> 16), (uint8_t)(n >> 8), (uint8_t)n};
    fwrite(len, 1, 4, f);
    fwrite(type, 1, 4, f);
    if (n) fwrite(data, 1, n, f);
    uLong crc = crc32(0L, reinterpret_cast<const Bytef*>(type), 4);
    if (n) crc = crc32(crc, data, (uInt)n);
    const uint8_t c[4] = {(uint8_t)(crc >> 24), (uint8_t)(crc >> 16), (uint8_t)(crc >> 8), (uint8_t)crc};
    fwrite(c, 1, 4, f);
}
}  // namespace

bool write_png_from_uchar4(const std::string& path, const uint32_t* frame, int width, int height) {
    if (!frame || width <= 0 || height <= 0) return false;
    std::vector<uint8_t> raw((size_t)height * (1 + (size_t)width * 3));
    size_t o = 0;
    for (int y = height - 1; y >= 0; y--) {   // the render's row 0 is the bottom row (image_io.hpp)
        raw[o++] = 0;                          // filter type: none
        const uint8_t* s = reinterpret_cast<const uint8_t*>(frame + (size_t)y * width);
        for (int x = 0; x < width; x++) {
            raw[o++] = s[4 * x];
            raw[o++] = s[4 * x + 1];
            raw[o++] = s[4 * x + 2];
        }
    }
    uLongf zn = compressBound((uLong)raw.size());
    std::vector<uint8_t> z(zn);
    if (compress2(z.data(), &zn, raw.data(), (uLong)raw.size(), 6) != Z_OK) return false;
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) return false;
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    fwrite(sig, 1, 8, f);
    const uint8_t ihdr[13] = {(uint8_t)(width >> 24), (uint8_t)(width >> 16), (uint8_t)(width >> 8), (uint8_t)width,
                              (uint8_t)(height >> 24), (uint8_t)(height >> 16), (uint8_t)(height >> 8), (uint8_t)height, 8, 2, 0, 0, 0};
    png_chunk(f, "IHDR", ihdr, sizeof(ihdr));
    png_chunk(f, "IDAT", z.data(), zn);
    png_chunk(f, "IEND", nullptr, 0);
    const bool ok = !ferror(f);
    return fclose(f) == 0 && ok;      // (a full disk shows up at the flush)
}

}  // namespace spchost
