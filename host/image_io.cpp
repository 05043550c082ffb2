// image_io.cpp -- see image_io.hpp
#include "image_io.hpp"

#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace spchost {

static bool read_file(const std::string& path, std::vector<uint8_t>& out) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    const long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    if (n < 0 || n > (1L << 31)) {   // not a regular file (fopen succeeds on a directory, whose "size" is LONG_MAX), or no texture
        fclose(f);
        return false;
    }
    out.resize((size_t)n);
    const size_t got = out.empty() ? 0 : fread(out.data(), 1, out.size(), f);
    fclose(f);
    return got == out.size();
}

static bool decode_cache(const std::vector<uint8_t>& d, ImageRGBA8& img) {
    if (d.size() < 16 || memcmp(d.data(), "SPCRGBA8", 8) != 0) return false;
    int32_t w, h;
    memcpy(&w, d.data() + 8, 4);
    memcpy(&h, d.data() + 12, 4);
    if (w <= 0 || h <= 0 || d.size() != 16 + (size_t)w * h * 4) return false;
    img.width = w;
    img.height = h;
    img.rgba.assign(d.begin() + 16, d.end());
    return true;
}

// binary PGM (P5) / PPM (P6), maxval <= 255
static bool decode_pnm(const std::vector<uint8_t>& d, ImageRGBA8& img) {
    if (d.size() < 7 || d[0] != 'P' || (d[1] != '5' && d[1] != '6')) return false;
    const int chan = d[1] == '6' ? 3 : 1;
    size_t p = 2;
    int vals[3];
    for (int k = 0; k < 3; k++) {
        for (;;) {
            while (p < d.size() && isspace(d[p])) p++;
            if (p < d.size() && d[p] == '#') {
                while (p < d.size() && d[p] != '\n') p++;
                continue;
            }
            break;
        }
        int v = 0, digits = 0;
        while (p < d.size() && isdigit(d[p])) {
            v = v * 10 + (d[p++] - '0');
            digits++;
        }
        if (!digits) return false;
        vals[k] = v;
    }
    p++;   // the single whitespace byte after maxval
    const int w = vals[0], h = vals[1];
    if (w <= 0 || h <= 0 || vals[2] <= 0 || vals[2] > 255 || p + (size_t)w * h * chan > d.size()) return false;
    img.width = w;
    img.height = h;
    img.rgba.resize((size_t)w * h * 4);
    for (size_t i = 0; i < (size_t)w * h; i++) {
        const uint8_t* s = d.data() + p + i * chan;
        img.rgba[4 * i] = s[0];
        img.rgba[4 * i + 1] = s[chan == 3 ? 1 : 0];
        img.rgba[4 * i + 2] = s[chan == 3 ? 2 : 0];
        img.rgba[4 * i + 3] = 255;
    }
    return true;
}

bool load_image_rgba8(const std::string& path, ImageRGBA8& out, std::string& err) {
    std::vector<uint8_t> d;
    if (read_file(path + ".rgba8", d) && decode_cache(d, out)) return true;   // pre-decoded cache next to the file wins
    if (!read_file(path, d)) {
        err = "cannot read image " + path;
        return false;
    }
    if (decode_cache(d, out)) return true;
    if (d.size() >= 2 && d[0] == 0xff && d[1] == 0xd8) return decode_jpeg_rgba8(d.data(), d.size(), out, err);
    if (d.size() >= 8 && d[0] == 0x89 && d[1] == 'P') return decode_png_rgba8(d.data(), d.size(), out, err);
    if (decode_pnm(d, out)) return true;
    err = "unsupported image format: " + path;
    return false;
}

bool write_rgba8_cache(const std::string& path, const ImageRGBA8& img) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) return false;
    const int32_t wh[2] = {img.width, img.height};
    bool ok = fwrite("SPCRGBA8", 1, 8, f) == 8 && fwrite(wh, 4, 2, f) == 2 && fwrite(img.rgba.data(), 1, img.rgba.size(), f) == img.rgba.size();
    fclose(f);
    return ok;
}

bool write_ppm_from_uchar4(const std::string& path, const uint32_t* frame, int width, int height) {
    if (!frame || width <= 0 || height <= 0) return false;
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) return false;
    fprintf(f, "P6\n%d %d\n255\n", width, height);
    std::vector<uint8_t> row((size_t)width * 3);
    for (int y = height - 1; y >= 0; y--) {
        const uint8_t* s = reinterpret_cast<const uint8_t*>(frame + (size_t)y * width);
        for (int x = 0; x < width; x++) {
            row[3 * x] = s[4 * x];
            row[3 * x + 1] = s[4 * x + 1];
            row[3 * x + 2] = s[4 * x + 2];
        }
        fwrite(row.data(), 1, row.size(), f);
    }
    const bool ok = !ferror(f);
    return fclose(f) == 0 && ok;
}

bool write_pfm_from_float4(const std::string& path, const float* accum4, int width, int height) {
    if (!accum4 || width <= 0 || height <= 0) return false;
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) return false;
    fprintf(f, "PF\n%d %d\n-1.0\n", width, height);
    std::vector<float> row((size_t)width * 3);
    for (int y = 0; y < height; y++) {
        const float* s = accum4 + (size_t)y * width * 4;
        for (int x = 0; x < width; x++) {
            row[3 * x] = s[4 * x];
            row[3 * x + 1] = s[4 * x + 1];
            row[3 * x + 2] = s[4 * x + 2];
        }
        fwrite(row.data(), sizeof(float), row.size(), f);
    }
    const bool ok = !ferror(f);
    return fclose(f) == 0 && ok;
}

bool read_pfm_rgb(const std::string& path, std::vector<float>& rgb, int& width, int& height, std::string& err) {
    std::vector<uint8_t> d;
    if (!read_file(path, d)) {
        err = "cannot read " + path;
        return false;
    }
    // header: "PF\n<w> <h>\n<scale>\n" (any whitespace between the tokens, one whitespace byte after the scale)
    size_t p = 0;
    auto token = [&](std::string& t) {
        while (p < d.size() && isspace(d[p])) p++;
        t.clear();
        while (p < d.size() && !isspace(d[p]) && t.size() < 32) t.push_back((char)d[p++]);
        return !t.empty();
    };
    std::string magic, sw, sh, ss;
    if (!token(magic) || magic != "PF" || !token(sw) || !token(sh) || !token(ss)) {
        err = "not a colour PFM file: " + path;
        return false;
    }
    p++;
    const long w = strtol(sw.c_str(), nullptr, 10), h = strtol(sh.c_str(), nullptr, 10);
    const double scale = strtod(ss.c_str(), nullptr);
    if (w <= 0 || h <= 0 || w > (1 << 16) || h > (1 << 16) || scale == 0.0 || p + (size_t)w * h * 12 > d.size()) {
        err = "bad PFM header or truncated data: " + path;
        return false;
    }
    width = (int)w;
    height = (int)h;
    rgb.resize((size_t)w * h * 3);
    const bool big_endian = scale > 0.0;
    for (size_t i = 0; i < rgb.size(); i++) {
        uint8_t b[4];
        memcpy(b, d.data() + p + 4 * i, 4);
        if (big_endian) {
            const uint8_t t0 = b[0], t1 = b[1];
            b[0] = b[3]; b[1] = b[2]; b[2] = t1; b[3] = t0;
        }
        memcpy(&rgb[i], b, 4);
    }
    return true;
}

double rel_mse(const std::vector<float>& img, const std::vector<float>& ref, size_t* skipped) {
    double sum = 0.0;
    size_t n = 0, bad = 0;
    const size_t m = img.size() < ref.size() ? img.size() : ref.size();
    for (size_t i = 0; i < m; i++) {
        const double r = ref[i], e = (double)img[i] - r;
        const double t = e * e / (r * r + 1e-2);
        if (std::isfinite(t)) {
            sum += t;
            n++;
        } else bad++;
    }
    if (skipped) *skipped = bad;
    return n ? sum / (double)n : 0.0;
}

}  // namespace spchost
