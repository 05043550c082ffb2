// jpeg_decode.cpp -- baseline and progressive (SOF0/SOF1/SOF2, 8-bit, Huffman) JPEG -> RGBA8 for the textures of a .scene file.
//
// The reference decodes textures with stb_image 2.x (scene_shift.cpp:39-40, stbi_load(..., STBI_rgb_alpha)) and
// uploads the bytes unchanged, so texel values are part of the shading parity contract.  The entropy decoding is
// fixed by ITU-T T.81; the three lossy-stage choices that differ between decoders are taken as stb_image makes them
// (observed in stb_image.h of the reference tree, lines cited) so that every byte agrees:
//   * inverse DCT: the 12-bit fixed-point "islow" butterfly, column pass keeps 2 extra bits (>>10), row pass >>17
//     with the +128 level shift folded into the rounding bias (stb_image.h:2392-2477); coefficients are de-quantised
//     into 16-bit storage first (:2213-2230);
//   * chroma up-sampling: centred bilinear 3:1 taps, (3a+b+2)>>2 and (3t0+t1+8)>>4 (:3411-3474), edges replicated;
//   * YCbCr -> RGB: 20-bit fixed point with the Cb term of G masked to its upper 16 bits (:3605-3630).
// Progressive files (SOF2: spectral selection + successive approximation, T.81 annex G) keep the whole image as 16-bit coefficients
// across their scans and are de-quantised and inverse-transformed once at the end, block by block, in 16-bit storage as stb_image
// does (:2084-2233 the two block decoders, :3208-3226 the final pass) -- the same three lossy stages, so the bytes agree as well.
// tests/test_host_loader.py pins the decoder against the reference's own stb_image compiled from the reference tree.
// Arithmetic-coded, lossless and hierarchical files are rejected.
#include <cstring>
#include <memory>

#include "image_io.hpp"

namespace spchost {
namespace {

struct HuffTable {
    bool     defined = false;
    uint8_t  bits[17] = {0};      // number of codes of each length
    uint8_t  vals[256] = {0};
    int      mincode[18], maxcode[18], valptr[18];
    HuffTable() { build(); }      // a table no DHT segment defined has no codes: every look-up fails ("bad huffman code")
    void build() {
        int code = 0, k = 0;
        for (int len = 1; len <= 16; len++) {
            valptr[len] = k;
            mincode[len] = code;
            code += bits[len];
            k += bits[len];
            maxcode[len] = bits[len] ? code - 1 : -1;
            code <<= 1;
        }
        maxcode[17] = 0x7fffffff;
    }
};

struct Component {
    int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0;
    int x = 0, y = 0;      // size in samples
    int w2 = 0, h2 = 0;    // allocated size: whole MCUs
    int dc_pred = 0;
    std::vector<uint8_t> data;
    std::vector<int16_t> coeff;   // progressive only: 64 coefficients per block, blocks in raster order, coeff_w blocks per row
    int coeff_w = 0;
};

struct Decoder {
    const uint8_t* p;
    const uint8_t* begin;
    const uint8_t* end;
    uint32_t bitbuf = 0;
    int      bitcnt = 0;
    int      marker = -1;     // marker met inside entropy data (0xD0.. etc.), -1 = none
    bool     nomore = false;

    uint16_t dequant[4][64] = {};
    HuffTable dc[4], ac[4];
    Component comp[4];
    int ncomp = 0, width = 0, height = 0, hmax = 1, vmax = 1, mcu_w = 0, mcu_h = 0, mcus_x = 0, mcus_y = 0;
    int restart_interval = 0;
    bool jfif = false;
    int  adobe_transform = -1, rgb_ids = 0;
    bool progressive = false;
    int  spec_start = 0, spec_end = 63, succ_high = 0, succ_low = 0;   // the current scan's Ss, Se, Ah, Al
    int  eob_run = 0;                                                   // progressive AC scans: blocks still covered by an end-of-band run
    std::string error;

    bool fail(const char* msg) {
        error = msg;
        return false;
    }
    int get8() { return p < end ? *p++ : 0; }
    int get16() {
        const int a = get8();
        return (a << 8) | get8();
    }

    // ---- entropy-coded segment bit reader: 0xFF00 is a stuffed 0xFF, any other marker ends the data (zeros follow)
    void fill() {
        while (bitcnt <= 24) {
            int b = nomore ? 0 : get8();
            if (b == 0xff && !nomore) {
                int c = get8();
                while (c == 0xff) c = get8();
                if (c != 0) {
                    marker = c;
                    nomore = true;
                    b = 0;
                }
            }
            bitbuf |= (uint32_t)b << (24 - bitcnt);
            bitcnt += 8;
        }
    }
    int bits(int n) {
        if (n == 0) return 0;
        if (bitcnt < n) fill();
        const int v = (int)(bitbuf >> (32 - n));
        bitbuf <<= n;
        bitcnt -= n;
        return v;
    }
    int huff(const HuffTable& t) {
        if (bitcnt < 16) fill();
        int code = 0;
        for (int len = 1; len <= 16; len++) {
            code = (code << 1) | (int)(bitbuf >> 31);
            bitbuf <<= 1;
            bitcnt--;
            if (t.maxcode[len] >= 0 && code <= t.maxcode[len] && code >= t.mincode[len]) return t.vals[t.valptr[len] + code - t.mincode[len]];
        }
        return -1;
    }
    static int extend(int v, int n) { return n == 0 ? 0 : (v < (1 << (n - 1)) ? v - (1 << n) + 1 : v); }
    void reset_entropy() {
        bitbuf = 0;
        bitcnt = 0;
        nomore = false;
        marker = -1;
        eob_run = 0;
        for (int i = 0; i < 4; i++) comp[i].dc_pred = 0;
    }
};

const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                             41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                             30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

inline uint8_t clamp8(int x) { return (uint8_t)(x < 0 ? 0 : (x > 255 ? 255 : x)); }

constexpr int fx(double x) { return (int)(x * 4096 + 0.5); }

// one 8-point pass of the islow inverse DCT; e[] = even part (4 values, x0..x3), o[] = odd part (t0..t3)
inline void idct8(int s0, int s1, int s2, int s3, int s4, int s5, int s6, int s7, int e[4], int o[4]) {
    int p1 = (s2 + s6) * fx(0.5411961f);
    const int t2 = p1 + s6 * fx(-1.847759065f);
    const int t3 = p1 + s2 * fx(0.765366865f);
    const int t0 = (s0 + s4) * 4096;
    const int t1 = (s0 - s4) * 4096;
    e[0] = t0 + t3;
    e[3] = t0 - t3;
    e[1] = t1 + t2;
    e[2] = t1 - t2;
    int a0 = s7, a1 = s5, a2 = s3, a3 = s1;
    int p3 = a0 + a2, p4 = a1 + a3;
    p1 = a0 + a3;
    int p2 = a1 + a2;
    const int p5 = (p3 + p4) * fx(1.175875602f);
    a0 = a0 * fx(0.298631336f);
    a1 = a1 * fx(2.053119869f);
    a2 = a2 * fx(3.072711026f);
    a3 = a3 * fx(1.501321110f);
    p1 = p5 + p1 * fx(-0.899976223f);
    p2 = p5 + p2 * fx(-2.562915447f);
    p3 = p3 * fx(-1.961570560f);
    p4 = p4 * fx(-0.390180644f);
    o[3] = a3 + p1 + p4;
    o[2] = a2 + p2 + p3;
    o[1] = a1 + p2 + p4;
    o[0] = a0 + p1 + p3;
}

void idct_block(uint8_t* out, int stride, const int16_t d[64]) {
    int tmp[64];
    for (int c = 0; c < 8; c++) {
        const int16_t* s = d + c;
        int* v = tmp + c;
        if (!(s[8] | s[16] | s[24] | s[32] | s[40] | s[48] | s[56])) {
            const int dc = s[0] * 4;
            for (int r = 0; r < 8; r++) v[8 * r] = dc;
            continue;
        }
        int e[4], o[4];
        idct8(s[0], s[8], s[16], s[24], s[32], s[40], s[48], s[56], e, o);
        for (int k = 0; k < 4; k++) {
            const int x = e[k] + 512;
            v[8 * k] = (x + o[3 - k]) >> 10;
            v[8 * (7 - k)] = (x - o[3 - k]) >> 10;
        }
    }
    for (int r = 0; r < 8; r++) {
        const int* v = tmp + 8 * r;
        uint8_t* q = out + (size_t)r * stride;
        int e[4], o[4];
        idct8(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], e, o);
        for (int k = 0; k < 4; k++) {
            const int x = e[k] + 65536 + (128 << 17);
            q[k] = clamp8((x + o[3 - k]) >> 17);
            q[7 - k] = clamp8((x - o[3 - k]) >> 17);
        }
    }
}

bool decode_block(Decoder& z, Component& c, int16_t blk[64]) {
    memset(blk, 0, 64 * sizeof(int16_t));
    const HuffTable& hd = z.dc[c.td];
    const HuffTable& ha = z.ac[c.ta];
    const uint16_t* dq = z.dequant[c.tq];
    const int t = z.huff(hd);
    if (t < 0 || t > 15) return z.fail("bad huffman code");
    const int diff = t ? Decoder::extend(z.bits(t), t) : 0;
    c.dc_pred += diff;
    blk[0] = (int16_t)(c.dc_pred * dq[0]);
    for (int k = 1; k < 64;) {
        const int rs = z.huff(ha);
        if (rs < 0) return z.fail("bad huffman code");
        const int r = rs >> 4, s = rs & 15;
        if (s == 0) {
            if (rs != 0xf0) break;   // end of block
            k += 16;
        } else {
            k += r;
            if (k > 63) return z.fail("bad coefficient index");
            const int zig = kZigzag[k++];
            blk[zig] = (int16_t)(Decoder::extend(z.bits(s), s) * dq[zig]);
        }
    }
    return true;
}

bool decode_scan(Decoder& z, const int* order, int n_scan) {
    z.reset_entropy();
    int todo = z.restart_interval ? z.restart_interval : 0x7fffffff;
    int16_t blk[64];
    auto restart = [&]() -> bool {   // false = stop decoding this scan (no RST marker where one is due)
        if (z.bitcnt < 24) z.fill();
        if (!(z.marker >= 0xd0 && z.marker <= 0xd7)) return false;
        z.reset_entropy();
        todo = z.restart_interval ? z.restart_interval : 0x7fffffff;
        return true;
    };
    if (n_scan == 1) {   // non-interleaved: the component's own blocks in raster order
        Component& c = z.comp[order[0]];
        const int bw = (c.x + 7) >> 3, bh = (c.y + 7) >> 3;
        for (int j = 0; j < bh; j++)
            for (int i = 0; i < bw; i++) {
                if (!decode_block(z, c, blk)) return false;
                idct_block(c.data.data() + (size_t)c.w2 * j * 8 + i * 8, c.w2, blk);
                if (--todo <= 0 && !restart()) return true;
            }
        return true;
    }
    for (int my = 0; my < z.mcus_y; my++)
        for (int mx = 0; mx < z.mcus_x; mx++) {
            for (int k = 0; k < n_scan; k++) {
                Component& c = z.comp[order[k]];
                for (int y = 0; y < c.v; y++)
                    for (int x = 0; x < c.h; x++) {
                        if (!decode_block(z, c, blk)) return false;
                        const int px = (mx * c.h + x) * 8, py = (my * c.v + y) * 8;
                        idct_block(c.data.data() + (size_t)c.w2 * py + px, c.w2, blk);
                    }
            }
            if (--todo <= 0 && !restart()) return true;
        }
    return true;
}

// ---- progressive scans (T.81 annex G) ----------------------------------------------------------------------------------------
// zig-zag index with the 15 entries a corrupt run can overshoot by mapped onto the last coefficient (as stb_image's table)
inline int zigzag_at(int k) { return k < 64 ? kZigzag[k] : 63; }

// DC scan: first pass stores the predicted DC value scaled by 2^Al, refinement passes add one bit
bool prog_dc(Decoder& z, Component& c, int16_t* blk) {
    if (z.spec_end != 0) return z.fail("progressive scan mixes DC and AC");
    if (z.succ_high == 0) {
        memset(blk, 0, 64 * sizeof(int16_t));
        const int t = z.huff(z.dc[c.td]);
        if (t < 0 || t > 15) return z.fail("bad huffman code");
        const int diff = t ? Decoder::extend(z.bits(t), t) : 0;
        c.dc_pred += diff;
        blk[0] = (int16_t)(c.dc_pred * (1 << z.succ_low));
    } else if (z.bits(1)) {
        blk[0] = (int16_t)(blk[0] + (int16_t)(1 << z.succ_low));
    }
    return true;
}

// one correction bit for an already non-zero coefficient (G.1.2.3): moves it away from zero by `bit` unless that bit is set
inline void refine(Decoder& z, int16_t& v, int bit) {
    if (z.bits(1) && (v & bit) == 0) v = (int16_t)(v > 0 ? v + bit : v - bit);
}

// AC scan over the band [Ss, Se] of one block
bool prog_ac(Decoder& z, Component& c, int16_t* blk) {
    if (z.spec_start == 0) return z.fail("progressive scan mixes DC and AC");
    const HuffTable& ha = z.ac[c.ta];
    if (z.succ_high == 0) {            // first pass of the band
        if (z.eob_run) {
            --z.eob_run;
            return true;
        }
        int k = z.spec_start;
        do {
            const int rs = z.huff(ha);
            if (rs < 0) return z.fail("bad huffman code");
            const int r = rs >> 4, sz = rs & 15;
            if (sz == 0) {
                if (r < 15) {          // end of band for this block and the next 2^r + extra - 1 blocks
                    z.eob_run = 1 << r;
                    if (r) z.eob_run += z.bits(r);
                    --z.eob_run;
                    break;
                }
                k += 16;
            } else {
                k += r;
                const int zig = zigzag_at(k++);
                blk[zig] = (int16_t)(Decoder::extend(z.bits(sz), sz) * (1 << z.succ_low));
            }
        } while (k <= z.spec_end);
        return true;
    }
    // refinement pass: every non-zero coefficient met gets a correction bit; new coefficients are +-2^Al after `r` zero ones
    const int bit = 1 << z.succ_low;
    if (z.eob_run) {
        --z.eob_run;
        for (int k = z.spec_start; k <= z.spec_end; k++) {
            int16_t& v = blk[kZigzag[k]];
            if (v != 0) refine(z, v, bit);
        }
        return true;
    }
    int k = z.spec_start;
    do {
        const int rs = z.huff(ha);
        if (rs < 0) return z.fail("bad huffman code");
        int r = rs >> 4, sz = rs & 15, val = 0;
        if (sz == 0) {
            if (r < 15) {
                z.eob_run = (1 << r) - 1;
                if (r) z.eob_run += z.bits(r);
                r = 64;                // run to the end of the band, correcting what is there
            }                          // (r == 15: sixteen zero coefficients -- a run of 15 followed by "placing" a zero)
        } else {
            if (sz != 1) return z.fail("bad huffman code");
            val = z.bits(1) ? bit : -bit;
        }
        while (k <= z.spec_end) {
            int16_t& v = blk[kZigzag[k++]];
            if (v != 0) refine(z, v, bit);
            else if (r == 0) {
                v = (int16_t)val;
                break;
            } else --r;
        }
    } while (k <= z.spec_end);
    return true;
}

bool decode_scan_progressive(Decoder& z, const int* order, int n_scan) {
    z.reset_entropy();
    int todo = z.restart_interval ? z.restart_interval : 0x7fffffff;
    auto restart = [&]() -> bool {
        if (z.bitcnt < 24) z.fill();
        if (!(z.marker >= 0xd0 && z.marker <= 0xd7)) return false;
        z.reset_entropy();
        todo = z.restart_interval ? z.restart_interval : 0x7fffffff;
        return true;
    };
    if (n_scan == 1) {
        Component& c = z.comp[order[0]];
        const int bw = (c.x + 7) >> 3, bh = (c.y + 7) >> 3;
        for (int j = 0; j < bh; j++)
            for (int i = 0; i < bw; i++) {
                int16_t* blk = c.coeff.data() + 64 * ((size_t)i + (size_t)j * c.coeff_w);
                if (!(z.spec_start == 0 ? prog_dc(z, c, blk) : prog_ac(z, c, blk))) return false;
                if (--todo <= 0 && !restart()) return true;
            }
        return true;
    }
    for (int my = 0; my < z.mcus_y; my++)       // interleaved scans of a progressive file carry DC coefficients only
        for (int mx = 0; mx < z.mcus_x; mx++) {
            for (int k = 0; k < n_scan; k++) {
                Component& c = z.comp[order[k]];
                for (int y = 0; y < c.v; y++)
                    for (int x = 0; x < c.h; x++) {
                        const size_t bx = (size_t)mx * c.h + x, by = (size_t)my * c.v + y;
                        if (!prog_dc(z, c, c.coeff.data() + 64 * (bx + by * c.coeff_w))) return false;
                    }
            }
            if (--todo <= 0 && !restart()) return true;
        }
    return true;
}

// after the last scan: de-quantise in 16-bit storage and inverse-transform every block the image covers
void finish_progressive(Decoder& z) {
    for (int n = 0; n < z.ncomp; n++) {
        Component& c = z.comp[n];
        const uint16_t* dq = z.dequant[c.tq];
        const int bw = (c.x + 7) >> 3, bh = (c.y + 7) >> 3;
        for (int j = 0; j < bh; j++)
            for (int i = 0; i < bw; i++) {
                int16_t* blk = c.coeff.data() + 64 * ((size_t)i + (size_t)j * c.coeff_w);
                for (int k = 0; k < 64; k++) blk[k] = (int16_t)(blk[k] * dq[k]);
                idct_block(c.data.data() + (size_t)c.w2 * j * 8 + i * 8, c.w2, blk);
            }
    }
}

bool read_tables_and_frame(Decoder& z, int m) {
    const int len = z.get16() - 2;
    if (len < 0 || z.p + len > z.end) return z.fail("truncated segment");
    const uint8_t* seg_end = z.p + len;
    switch (m) {
        case 0xdb:   // DQT
            while (z.p < seg_end) {
                const int q = z.get8(), prec = q >> 4, t = q & 15;
                if (t > 3 || prec > 1) return z.fail("bad DQT");
                for (int i = 0; i < 64; i++) z.dequant[t][kZigzag[i]] = (uint16_t)(prec ? z.get16() : z.get8());
            }
            break;
        case 0xc4:   // DHT
            while (z.p < seg_end) {
                const int q = z.get8(), cls = q >> 4, t = q & 15;
                if (cls > 1 || t > 3) return z.fail("bad DHT");
                HuffTable& h = cls ? z.ac[t] : z.dc[t];
                int n = 0;
                for (int i = 1; i <= 16; i++) n += (h.bits[i] = (uint8_t)z.get8());
                if (n > 256) return z.fail("bad DHT");
                for (int i = 0; i < n; i++) h.vals[i] = (uint8_t)z.get8();
                h.build();
                h.defined = true;
            }
            break;
        case 0xdd:   // DRI
            z.restart_interval = z.get16();
            break;
        case 0xe0:   // APP0: "JFIF\0"
            if (len >= 5 && memcmp(z.p, "JFIF\0", 5) == 0) z.jfif = true;
            break;
        case 0xee:   // APP14: "Adobe\0" ... transform byte
            if (len >= 12 && memcmp(z.p, "Adobe\0", 6) == 0) z.adobe_transform = z.p[11];
            break;
        case 0xc0: case 0xc1: case 0xc2: {   // SOF0/1/2
            z.progressive = m == 0xc2;
            if (z.get8() != 8) return z.fail("only 8-bit JPEG is supported");
            z.height = z.get16();
            z.width = z.get16();
            z.ncomp = z.get8();
            if (z.width <= 0 || z.height <= 0) return z.fail("bad JPEG size");
            // untrusted files: no more than 2^28 pixels, and no frame the file cannot fill (every block costs at least one bit per scan)
            if ((uint64_t)z.width * (uint64_t)z.height > (1ull << 28)) return z.fail("JPEG exceeds 2^28 pixels");
            if ((uint64_t)z.width * (uint64_t)z.height > (uint64_t)(z.end - z.begin) * 1024 + (1u << 20)) return z.fail("JPEG frame size does not match its data");
            if (z.ncomp != 1 && z.ncomp != 3) return z.fail("only 1- and 3-component JPEG is supported");
            z.rgb_ids = 0;
            static const char rgb[3] = {'R', 'G', 'B'};
            for (int i = 0; i < z.ncomp; i++) {
                Component& c = z.comp[i];
                c.id = z.get8();
                if (z.ncomp == 3 && c.id == rgb[i]) z.rgb_ids++;
                const int q = z.get8();
                c.h = q >> 4;
                c.v = q & 15;
                c.tq = z.get8();
                if (c.h < 1 || c.h > 4 || c.v < 1 || c.v > 4 || c.tq > 3) return z.fail("bad SOF component");
                if (c.h > z.hmax) z.hmax = c.h;
                if (c.v > z.vmax) z.vmax = c.v;
            }
            // the resampler works with integer ratios hmax / h, vmax / v: a frame whose factors do not divide (e.g. 3 and 2)
            // would make it read past the component rows (stb_image has the same weakness; textures are untrusted files)
            for (int i = 0; i < z.ncomp; i++)
                if (z.hmax % z.comp[i].h != 0 || z.vmax % z.comp[i].v != 0) return z.fail("unsupported sampling factors (not integer ratios)");
            z.mcu_w = z.hmax * 8;
            z.mcu_h = z.vmax * 8;
            z.mcus_x = (z.width + z.mcu_w - 1) / z.mcu_w;
            z.mcus_y = (z.height + z.mcu_h - 1) / z.mcu_h;
            for (int i = 0; i < z.ncomp; i++) {
                Component& c = z.comp[i];
                c.x = (z.width * c.h + z.hmax - 1) / z.hmax;
                c.y = (z.height * c.v + z.vmax - 1) / z.vmax;
                c.w2 = z.mcus_x * c.h * 8;
                c.h2 = z.mcus_y * c.v * 8;
                c.data.assign((size_t)c.w2 * c.h2, 0);
                if (z.progressive) {
                    c.coeff_w = c.w2 / 8;
                    c.coeff.assign((size_t)c.w2 * c.h2, 0);
                }
            }
            break;
        }
        default:
            break;
    }
    z.p = seg_end;
    return true;
}

// up-sample one component row to full width (stb_image.h:3402-3474, 3592-3603)
const uint8_t* upsample_row(std::vector<uint8_t>& out, const uint8_t* near, const uint8_t* far, int w, int hs, int vs) {
    if (hs == 1 && vs == 1) return near;
    if (hs == 1 && vs == 2) {
        for (int i = 0; i < w; i++) out[i] = (uint8_t)((3 * near[i] + far[i] + 2) >> 2);
        return out.data();
    }
    if (hs == 2 && vs == 1) {
        if (w == 1) {
            out[0] = out[1] = near[0];
            return out.data();
        }
        out[0] = near[0];
        out[1] = (uint8_t)((near[0] * 3 + near[1] + 2) >> 2);
        int i = 1;
        for (; i < w - 1; i++) {
            const int n = 3 * near[i] + 2;
            out[2 * i] = (uint8_t)((n + near[i - 1]) >> 2);
            out[2 * i + 1] = (uint8_t)((n + near[i + 1]) >> 2);
        }
        out[2 * i] = (uint8_t)((near[w - 2] * 3 + near[w - 1] + 2) >> 2);
        out[2 * i + 1] = near[w - 1];
        return out.data();
    }
    if (hs == 2 && vs == 2) {
        if (w == 1) {
            out[0] = out[1] = (uint8_t)((3 * near[0] + far[0] + 2) >> 2);
            return out.data();
        }
        int t1 = 3 * near[0] + far[0];
        out[0] = (uint8_t)((t1 + 2) >> 2);
        for (int i = 1; i < w; i++) {
            const int t0 = t1;
            t1 = 3 * near[i] + far[i];
            out[2 * i - 1] = (uint8_t)((3 * t0 + t1 + 8) >> 4);
            out[2 * i] = (uint8_t)((3 * t1 + t0 + 8) >> 4);
        }
        out[2 * w - 1] = (uint8_t)((t1 + 2) >> 2);
        return out.data();
    }
    for (int i = 0; i < w; i++)   // other ratios: nearest neighbour
        for (int j = 0; j < hs; j++) out[i * hs + j] = near[i];
    return out.data();
}

constexpr int f2f20(float x) { return ((int)(x * 4096.0f + 0.5f)) << 8; }

}  // namespace

bool decode_jpeg_rgba8(const uint8_t* data, size_t size, ImageRGBA8& img, std::string& err) {
    auto zp = std::make_unique<Decoder>();
    Decoder& z = *zp;
    z.p = z.begin = data;
    z.end = data + size;
    if (size < 4 || z.get8() != 0xff || z.get8() != 0xd8) {
        err = "not a JPEG file";
        return false;
    }
    bool have_frame = false, done = false;
    int pending = -1;
    while (!done) {
        int m = pending;
        pending = -1;
        if (m < 0) {
            int b = z.get8();
            if (z.p >= z.end) break;
            if (b != 0xff) continue;   // skip fill / junk between segments
            m = z.get8();
            while (m == 0xff) m = z.get8();
            if (m == 0) continue;
        }
        if (m == 0xd9) break;   // EOI
        if (m >= 0xd0 && m <= 0xd7) continue;
        if (m == 0xda) {   // SOS
            if (!have_frame) { err = "SOS before SOF"; return false; }
            z.get16();
            const int n_scan = z.get8();
            if (n_scan < 1 || n_scan > z.ncomp) { err = "bad SOS"; return false; }
            int order[4];
            for (int i = 0; i < n_scan; i++) {
                const int id = z.get8(), q = z.get8();
                int which = -1;
                for (int k = 0; k < z.ncomp; k++)
                    if (z.comp[k].id == id) which = k;
                if (which < 0) { err = "bad SOS component"; return false; }
                z.comp[which].td = q >> 4;
                z.comp[which].ta = q & 15;
                if (z.comp[which].td > 3 || z.comp[which].ta > 3) { err = "bad SOS table"; return false; }
                order[i] = which;
            }
            z.spec_start = z.get8();
            z.spec_end = z.get8();
            const int approx = z.get8();
            z.succ_high = approx >> 4;
            z.succ_low = approx & 15;
            if (z.progressive) {
                if (z.spec_start > 63 || z.spec_end > 63 || z.spec_start > z.spec_end || z.succ_high > 13 || z.succ_low > 13) { err = "bad SOS"; return false; }
                if (!decode_scan_progressive(z, order, n_scan)) { err = z.error; return false; }
            } else {
                if (z.spec_start != 0 || z.succ_high != 0 || z.succ_low != 0) { err = "bad SOS"; return false; }
                if (!decode_scan(z, order, n_scan)) { err = z.error; return false; }
            }
            if (z.marker >= 0) pending = z.marker;
            z.marker = -1;
            continue;
        }
        if (m == 0xc3 || (m >= 0xc5 && m <= 0xcf && m != 0xc8 && m != 0xcc)) {
            err = "unsupported JPEG coding process (lossless / arithmetic / hierarchical)";
            return false;
        }
        if (!read_tables_and_frame(z, m)) { err = z.error; return false; }
        if (m == 0xc0 || m == 0xc1 || m == 0xc2) have_frame = true;
    }
    if (!have_frame) {
        err = "no frame in JPEG";
        return false;
    }
    if (z.progressive) finish_progressive(z);

    // ---- up-sample and colour-convert row by row (stb_image.h:3840-3905) -----------------------------------
    const int W = z.width, H = z.height;
    img.width = W;
    img.height = H;
    img.rgba.assign((size_t)W * H * 4, 255);
    const bool is_rgb = z.ncomp == 3 && (z.rgb_ids == 3 || (z.adobe_transform == 0 && !z.jfif));
    struct Resample { int hs, vs, ystep, w_lores, ypos; const uint8_t *line0, *line1; std::vector<uint8_t> buf; } rs[3];
    for (int k = 0; k < z.ncomp; k++) {
        Resample& r = rs[k];
        r.hs = z.hmax / z.comp[k].h;
        r.vs = z.vmax / z.comp[k].v;
        r.ystep = r.vs >> 1;
        r.w_lores = (W + r.hs - 1) / r.hs;
        r.ypos = 0;
        r.line0 = r.line1 = z.comp[k].data.data();
        r.buf.assign((size_t)W + 8, 0);
    }
    const uint8_t* rows[3] = {nullptr, nullptr, nullptr};
    for (int j = 0; j < H; j++) {
        uint8_t* out = img.rgba.data() + (size_t)j * W * 4;
        for (int k = 0; k < z.ncomp; k++) {
            Resample& r = rs[k];
            const bool bottom = r.ystep >= (r.vs >> 1);
            rows[k] = upsample_row(r.buf, bottom ? r.line1 : r.line0, bottom ? r.line0 : r.line1, r.w_lores, r.hs, r.vs);
            if (++r.ystep >= r.vs) {
                r.ystep = 0;
                r.line0 = r.line1;
                if (++r.ypos < z.comp[k].y) r.line1 += z.comp[k].w2;
            }
        }
        if (z.ncomp == 1) {
            for (int i = 0; i < W; i++) out[4 * i] = out[4 * i + 1] = out[4 * i + 2] = rows[0][i];
        } else if (is_rgb) {
            for (int i = 0; i < W; i++) {
                out[4 * i] = rows[0][i];
                out[4 * i + 1] = rows[1][i];
                out[4 * i + 2] = rows[2][i];
            }
        } else {
            for (int i = 0; i < W; i++) {
                const int yf = (rows[0][i] << 20) + (1 << 19);
                const int cb = rows[1][i] - 128, cr = rows[2][i] - 128;
                const int r = (yf + cr * f2f20(1.40200f)) >> 20;
                const int g = (int)(yf + cr * -f2f20(0.71414f) + (int)((unsigned)(cb * -f2f20(0.34414f)) & 0xffff0000u)) >> 20;
                const int b = (yf + cb * f2f20(1.77200f)) >> 20;
                out[4 * i] = clamp8(r);
                out[4 * i + 1] = clamp8(g);
                out[4 * i + 2] = clamp8(b);
            }
        }
    }
    return true;
}

}  // namespace spchost
