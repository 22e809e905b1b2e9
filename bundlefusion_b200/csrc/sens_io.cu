// sens_io.cu -- `.sens` reader / writer, host code only.  Implements include/bf_sens.h (first half).
//
// Behavioural sources (external/mLib = /root/reference/external/mLib/include; FL/ = /root/reference/FriedLiver/Source/):
//   on-disk layout, version 4: ml::SensorData::loadFromFile (ext-depthcamera/sensorData.h:1187-1227), CalibrationData (:248-256), RGBDFrame (:676-700),
//   IMUFrame (:739-756); compression enums (:289-300); what the frame loop gets: SensorDataReader::processDepth (FL/SensorDataReader.cpp:100-117).
// mLib decodes pixel data with a vendored stb_image; this file has its own decoders: zlib streams through the system's zlib, PNG (inflate + the five
// scan-line filters of the PNG specification) and baseline JPEG (ITU-T T.81: Huffman-coded sequential DCT; IDCT evaluated from the definition in
// float; chroma up-sampling by the triangle filter libjpeg and stb_image both default to; JFIF YCbCr -> RGB).  Nothing here touches the GPU.
#include <zlib.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/bf_sens.h"

#define BF_API extern "C" __attribute__((visibility("default")))

namespace {

// ---- zlib ---------------------------------------------------------------------------------------------------------------------------------
int inflate_all(const uint8_t* src, size_t n, std::vector<uint8_t>& out, size_t expected) {
    out.resize(expected ? expected : n * 4 + 64);
    z_stream z; memset(&z, 0, sizeof(z));
    if (inflateInit(&z) != Z_OK) return BF_SENS_ERR_FORMAT;
    z.next_in = const_cast<Bytef*>(src); z.avail_in = (uInt)n;
    size_t have = 0;
    for (;;) {
        z.next_out = out.data() + have; z.avail_out = (uInt)(out.size() - have);
        const int rc = inflate(&z, Z_NO_FLUSH);
        have = out.size() - z.avail_out;
        if (rc == Z_STREAM_END) break;
        if (rc != Z_OK) { inflateEnd(&z); return BF_SENS_ERR_FORMAT; }
        if (z.avail_out == 0) out.resize(out.size() * 2);
        else if (z.avail_in == 0) { inflateEnd(&z); return BF_SENS_ERR_FORMAT; }          // truncated stream
    }
    inflateEnd(&z);
    out.resize(have);
    return BF_SENS_OK;
}

// ---- PNG (ISO/IEC 15948): 8-bit grey, grey + alpha, RGB, RGBA, non-interlaced -----------------------------------------------------------------
uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
int paeth(int a, int b, int c) { const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c); return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c); }

int decode_png(const uint8_t* d, size_t n, uint8_t* rgb, uint32_t* W, uint32_t* H) {
    static const uint8_t sig[8] = { 0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A };
    if (n < 8 + 25 || memcmp(d, sig, 8) != 0) return BF_SENS_ERR_FORMAT;
    size_t at = 8;
    uint32_t w = 0, h = 0; int depth = 0, type = -1, interlace = 0;
    std::vector<uint8_t> idat;
    uint8_t palette[256 * 3]; unsigned paletteLen = 0;
    while (at + 12 <= n) {
        const uint32_t len = be32(d + at); const uint8_t* tag = d + at + 4; const uint8_t* body = d + at + 8;
        if (at + 12 + (size_t)len > n) return BF_SENS_ERR_FORMAT;
        if (!memcmp(tag, "IHDR", 4)) { if (len < 13) return BF_SENS_ERR_FORMAT; w = be32(body); h = be32(body + 4); depth = body[8]; type = body[9]; interlace = body[12]; }
        else if (!memcmp(tag, "PLTE", 4)) { if (len > 768 || len % 3) return BF_SENS_ERR_FORMAT; memcpy(palette, body, len); paletteLen = len / 3; }
        else if (!memcmp(tag, "IDAT", 4)) idat.insert(idat.end(), body, body + len);
        else if (!memcmp(tag, "IEND", 4)) break;
        at += 12 + (size_t)len;
    }
    if (w == 0 || h == 0) return BF_SENS_ERR_FORMAT;
    if (w > (1u << 24) || h > (1u << 24) || (1u << 30) / w / 4 < h) return BF_SENS_ERR_FORMAT;          // "too large" in the reference's decoder: a corrupt header must not drive an allocation
    // what the reference's decoder (stb_image v2.08) reads: 8 bits per channel for every colour type, 1 / 2 / 4 bits for grey and palette images, Adam7 interlacing; not 16 bits
    const bool small = depth == 1 || depth == 2 || depth == 4;
    if (!(type == 0 || type == 2 || type == 3 || type == 4 || type == 6) || interlace > 1) return BF_SENS_ERR_UNSUPPORTED;
    if (!(depth == 8 || (small && (type == 0 || type == 3)))) return BF_SENS_ERR_UNSUPPORTED;
    if (type == 3 && paletteLen == 0) return BF_SENS_ERR_FORMAT;
    *W = w; *H = h;
    if (!rgb) return BF_SENS_OK;
    const int ch = (type == 0 || type == 3) ? 1 : (type == 4 ? 2 : (type == 2 ? 3 : 4));
    const int bitsPerPixel = ch * depth, fbpp = bitsPerPixel >= 8 ? bitsPerPixel / 8 : 1;       // the filters work on bytes; below 8 bits a "pixel" is the previous byte
    // Adam7: seven reduced images, each filtered and stored like an image of its own (PNG 1.2, section 8.2); not interlaced: one pass over everything
    static const int px0[7] = { 0, 4, 0, 2, 0, 1, 0 }, py0[7] = { 0, 0, 4, 0, 2, 0, 1 }, pdx[7] = { 8, 8, 4, 4, 2, 2, 1 }, pdy[7] = { 8, 8, 8, 4, 4, 2, 2 };
    size_t need = 0;
    for (int p = 0; p < (interlace ? 7 : 1); ++p) {
        const uint32_t pw = interlace ? (w + pdx[p] - 1 - px0[p]) / pdx[p] : w, ph = interlace ? (h + pdy[p] - 1 - py0[p]) / pdy[p] : h;
        if (pw && ph) need += (((size_t)pw * bitsPerPixel + 7) / 8 + 1) * ph;
    }
    std::vector<uint8_t> raw;
    const int rc = inflate_all(idat.data(), idat.size(), raw, need);
    if (rc) return rc;
    if (raw.size() < need) return BF_SENS_ERR_FORMAT;
    const int greyScale = depth == 1 ? 255 : (depth == 2 ? 85 : (depth == 4 ? 17 : 1));
    const uint8_t* s = raw.data();
    for (int p = 0; p < (interlace ? 7 : 1); ++p) {
        const uint32_t pw = interlace ? (w + pdx[p] - 1 - px0[p]) / pdx[p] : w, ph = interlace ? (h + pdy[p] - 1 - py0[p]) / pdy[p] : h;
        if (!pw || !ph) continue;
        const size_t stride = ((size_t)pw * bitsPerPixel + 7) / 8;
        std::vector<uint8_t> prev(stride, 0), cur(stride);
        for (uint32_t y = 0; y < ph; ++y) {
            const int f = s[0]; ++s;
            for (size_t i = 0; i < stride; ++i) {
                const int a = i >= (size_t)fbpp ? cur[i - fbpp] : 0, bb = prev[i], c = i >= (size_t)fbpp ? prev[i - fbpp] : 0;
                int v = s[i];
                switch (f) { case 0: break; case 1: v += a; break; case 2: v += bb; break; case 3: v += (a + bb) >> 1; break; case 4: v += paeth(a, bb, c); break; default: return BF_SENS_ERR_FORMAT; }
                cur[i] = (uint8_t)v;
            }
            s += stride;
            const uint32_t oy = interlace ? (uint32_t)py0[p] + y * (uint32_t)pdy[p] : y;
            for (uint32_t x = 0; x < pw; ++x) {
                const uint32_t ox = interlace ? (uint32_t)px0[p] + x * (uint32_t)pdx[p] : x;
                uint8_t* o = rgb + ((size_t)oy * w + ox) * 3;
                if (ch == 1) {
                    unsigned v = depth == 8 ? cur[x] : (cur[((size_t)x * depth) >> 3] >> (8 - depth - (int)(((size_t)x * depth) & 7))) & ((1u << depth) - 1u);      // samples are packed from the high bit down
                    if (type == 3) {
                        if (v >= paletteLen) return BF_SENS_ERR_FORMAT;
                        o[0] = palette[3 * v]; o[1] = palette[3 * v + 1]; o[2] = palette[3 * v + 2];
                    } else o[0] = o[1] = o[2] = (uint8_t)(v * greyScale);
                } else {
                    const uint8_t* px = &cur[(size_t)x * ch];
                    if (ch == 2) { o[0] = o[1] = o[2] = px[0]; } else { o[0] = px[0]; o[1] = px[1]; o[2] = px[2]; }
                }
            }
            prev.swap(cur);
        }
    }
    return BF_SENS_OK;
}

// ---- baseline JPEG (ITU-T T.81) ---------------------------------------------------------------------------------------------------------------
struct Huff { uint8_t bits[17]; uint8_t vals[256]; int mincode[17], maxcode[18], valptr[17]; bool set = false; };
void huff_build(Huff& h) {                          // T.81 Annex C / F.2.2.3
    int code = 0, k = 0;
    for (int l = 1; l <= 16; ++l) {
        h.valptr[l] = k; h.mincode[l] = code;
        code += h.bits[l]; k += h.bits[l];
        h.maxcode[l] = h.bits[l] ? code - 1 : -1;
        code <<= 1;
    }
    h.maxcode[17] = 0x7FFFFFFF;
    h.set = true;
}
struct BitReader {
    const uint8_t* p; const uint8_t* end; uint32_t acc = 0; int cnt = 0; bool hitMarker = false;
    void fill() {
        while (cnt <= 24) {
            int b = 0;
            if (!hitMarker && p < end) {
                b = *p;
                if (b == 0xFF) {
                    if (p + 1 < end && p[1] == 0x00) p += 2;          // stuffed byte
                    else { hitMarker = true; b = 0; }                  // a marker: feed zeros until the caller deals with it
                } else ++p;
            }
            acc |= (uint32_t)b << (24 - cnt); cnt += 8;
        }
    }
    int bit() { if (cnt == 0) fill(); const int v = (int)(acc >> 31); acc <<= 1; --cnt; return v; }
    int bits(int n) { if (n == 0) return 0; if (cnt < n) fill(); const int v = (int)(acc >> (32 - n)); acc <<= n; cnt -= n; return v; }
    void reset() { acc = 0; cnt = 0; hitMarker = false; }
};
int huff_decode(BitReader& br, const Huff& h) {
    int code = 0;
    for (int l = 1; l <= 16; ++l) {
        code = (code << 1) | br.bit();
        if (h.maxcode[l] >= 0 && code <= h.maxcode[l] && code >= h.mincode[l]) return h.vals[h.valptr[l] + code - h.mincode[l]];
    }
    return -1;
}
int extend(int v, int t) { return (t && v < (1 << (t - 1))) ? v - (1 << t) + 1 : v; }          // T.81 F.2.2.1
const uint8_t kZigzag[64] = { 0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                              35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63 };
// The inverse DCT of the reference's decoder (stb_image v2.08, vendored by mLib: a fixed-point rendering of the Loeffler-Ligtenberg-Moschytz flow graph, 12
// fractional bits, two extra bits kept between the passes), restated so that decoded pixels are the reference's, bit for bit: decoders are free to differ in the
// last level here, and SIFT would see the difference.  Input: de-quantised coefficients as 16-bit integers in natural order.
#define BF_F2F(x) ((int)(((x) * 4096 + 0.5)))
struct Lines { int x0, x1, x2, x3, t0, t1, t2, t3; };
static inline Lines idct_1d(int s0, int s1, int s2, int s3, int s4, int s5, int s6, int s7) {
    int p2 = s2, p3 = s6;
    int p1 = (p2 + p3) * BF_F2F(0.5411961f);
    int t2 = p1 + p3 * BF_F2F(-1.847759065f), t3 = p1 + p2 * BF_F2F(0.765366865f);
    p2 = s0; p3 = s4;
    int t0 = (p2 + p3) << 12, t1 = (p2 - p3) << 12;
    Lines L;
    L.x0 = t0 + t3; L.x3 = t0 - t3; L.x1 = t1 + t2; L.x2 = t1 - t2;
    t0 = s7; t1 = s5; t2 = s3; t3 = s1;
    p3 = t0 + t2; int p4 = t1 + t3; p1 = t0 + t3; p2 = t1 + t2;
    const int p5 = (p3 + p4) * BF_F2F(1.175875602f);
    t0 = t0 * BF_F2F(0.298631336f); t1 = t1 * BF_F2F(2.053119869f); t2 = t2 * BF_F2F(3.072711026f); t3 = t3 * BF_F2F(1.501321110f);
    p1 = p5 + p1 * BF_F2F(-0.899976223f); p2 = p5 + p2 * BF_F2F(-2.562915447f); p3 = p3 * BF_F2F(-1.961570560f); p4 = p4 * BF_F2F(-0.390180644f);
    L.t3 = t3 + p1 + p4; L.t2 = t2 + p2 + p3; L.t1 = t1 + p2 + p4; L.t0 = t0 + p1 + p3;
    return L;
}
static inline uint8_t clamp_u8(int x) { return (uint8_t)(x < 0 ? 0 : (x > 255 ? 255 : x)); }
void idct_block(const short d[64], uint8_t* out, size_t stride) {
    int val[64];
    for (int i = 0; i < 8; ++i) {                    // columns; a column whose AC terms are all zero is its DC term, scaled like the others
        const short* c = d + i; int* v = val + i;
        if (c[8] == 0 && c[16] == 0 && c[24] == 0 && c[32] == 0 && c[40] == 0 && c[48] == 0 && c[56] == 0) { const int dc = c[0] << 2; for (int k = 0; k < 8; ++k) v[8 * k] = dc; continue; }
        Lines L = idct_1d(c[0], c[8], c[16], c[24], c[32], c[40], c[48], c[56]);
        L.x0 += 512; L.x1 += 512; L.x2 += 512; L.x3 += 512;                                   // 12 bits down to 2: round at bit 10
        v[0] = (L.x0 + L.t3) >> 10; v[56] = (L.x0 - L.t3) >> 10; v[8] = (L.x1 + L.t2) >> 10; v[48] = (L.x1 - L.t2) >> 10;
        v[16] = (L.x2 + L.t1) >> 10; v[40] = (L.x2 - L.t1) >> 10; v[24] = (L.x3 + L.t0) >> 10; v[32] = (L.x3 - L.t0) >> 10;
    }
    for (int i = 0; i < 8; ++i) {                    // rows: 12 + 2 + 3 bits to remove, level shift by 128 folded into the rounding constant
        const int* v = val + 8 * i; uint8_t* o = out + (size_t)i * stride;
        Lines L = idct_1d(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]);
        const int bias = 65536 + (128 << 17);
        L.x0 += bias; L.x1 += bias; L.x2 += bias; L.x3 += bias;
        o[0] = clamp_u8((L.x0 + L.t3) >> 17); o[7] = clamp_u8((L.x0 - L.t3) >> 17); o[1] = clamp_u8((L.x1 + L.t2) >> 17); o[6] = clamp_u8((L.x1 - L.t2) >> 17);
        o[2] = clamp_u8((L.x2 + L.t1) >> 17); o[5] = clamp_u8((L.x2 - L.t1) >> 17); o[3] = clamp_u8((L.x3 + L.t0) >> 17); o[4] = clamp_u8((L.x3 - L.t0) >> 17);
    }
}
struct Comp { int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0, pred = 0; int bw = 0, bh = 0; size_t stride = 0; std::vector<uint8_t> plane; std::vector<short> coeff; };     // coeff: progressive only, [bh][bw][64]

// One block of a progressive scan (T.81 G.1.2; the reference's decoder: stb_image v2.08 stbi__jpeg_decode_block_prog_dc / _prog_ac).  Coefficients accumulate in
// `data` (natural order, not yet de-quantised) over the scans; eobRun is the scan's count of further blocks that are all end-of-band.
int prog_block(BitReader& br, const Huff* hdc, const Huff* hac, Comp* c, short* data, int ss, int se, int ah, int al, int& eobRun) {
    if (ss == 0) {                                                                     // DC: first scan carries the value (shifted), later ones one more bit each
        if (se != 0) return BF_SENS_ERR_FORMAT;
        if (ah == 0) {
            const int t = huff_decode(br, *hdc);
            if (t < 0 || t > 15) return BF_SENS_ERR_FORMAT;
            c->pred += extend(br.bits(t), t);
            data[0] = (short)(c->pred << al);
        } else if (br.bit()) data[0] += (short)(1 << al);
        return BF_SENS_OK;
    }
    if (ah == 0) {                                                                     // AC, first pass over the band [ss, se]
        if (eobRun) { --eobRun; return BF_SENS_OK; }
        int k = ss;
        do {
            const int rs = huff_decode(br, *hac);
            if (rs < 0) return BF_SENS_ERR_FORMAT;
            const int sz = rs & 15, r = rs >> 4;
            if (sz == 0) {
                if (r < 15) { eobRun = 1 << r; if (r) eobRun += br.bits(r); --eobRun; break; }
                k += 16;
            } else {
                k += r;
                if (k > 63) return BF_SENS_ERR_FORMAT;
                data[kZigzag[k++]] = (short)(extend(br.bits(sz), sz) << al);
            }
        } while (k <= se);
        return BF_SENS_OK;
    }
    const short bit = (short)(1 << al);                                                // AC refinement: one more bit for known coefficients, new ones of magnitude 1
    auto refine = [&](short* q) { if (br.bit() && (*q & bit) == 0) { if (*q > 0) *q += bit; else *q -= bit; } };
    if (eobRun) {
        --eobRun;
        for (int k = ss; k <= se; ++k) { short* q = &data[kZigzag[k]]; if (*q != 0) refine(q); }
        return BF_SENS_OK;
    }
    int k = ss;
    do {
        const int rs = huff_decode(br, *hac);
        if (rs < 0) return BF_SENS_ERR_FORMAT;
        int sz = rs & 15, r = rs >> 4;
        if (sz == 0) {
            if (r < 15) { eobRun = (1 << r) - 1; if (r) eobRun += br.bits(r); r = 64; }      // the rest of this block: only refinements
        } else {
            if (sz != 1) return BF_SENS_ERR_FORMAT;
            sz = br.bit() ? bit : -bit;
        }
        while (k <= se) {                                                              // skip r zero-history coefficients, refining the non-zero ones passed
            short* q = &data[kZigzag[k++]];
            if (*q != 0) refine(q);
            else { if (r == 0) { *q = (short)sz; break; } --r; }
        }
    } while (k <= se);
    return BF_SENS_OK;
}


int decode_jpeg(const uint8_t* d, size_t n, uint8_t* rgb, uint32_t* W, uint32_t* H) {
    if (n < 4 || d[0] != 0xFF || d[1] != 0xD8) return BF_SENS_ERR_FORMAT;
    uint16_t qt[4][64]; bool qset[4] = { false, false, false, false };
    Huff hdc[4], hac[4];
    std::vector<Comp> comps;
    int width = 0, height = 0, hmax = 1, vmax = 1, restart = 0;
    bool haveFrame = false, decoded = false, progressive = false;
    size_t at = 2;
    while (at + 4 <= n) {
        if (d[at] != 0xFF) { ++at; continue; }
        const int m = d[at + 1];
        if (m == 0xFF) { ++at; continue; }
        if (m == 0x00) { at += 2; continue; }                                         // a stuffed data byte, not a marker
        at += 2;
        if (m == 0xD9) break;
        if (m == 0x01 || (m >= 0xD0 && m <= 0xD7)) continue;
        if (at + 2 > n) return BF_SENS_ERR_FORMAT;
        const size_t len = ((size_t)d[at] << 8) | d[at + 1];
        if (len < 2 || at + len > n) return BF_SENS_ERR_FORMAT;
        const uint8_t* s = d + at + 2; const size_t sl = len - 2;
        if (m == 0xDB) {                                                              // DQT
            size_t i = 0;
            while (i < sl) {
                const int pq = s[i] >> 4, tq = s[i] & 15; ++i;
                if (tq > 3 || i + (pq ? 128 : 64) > sl) return BF_SENS_ERR_FORMAT;
                for (int k = 0; k < 64; ++k) { qt[tq][k] = pq ? (uint16_t)((s[i] << 8) | s[i + 1]) : s[i]; i += pq ? 2 : 1; }
                qset[tq] = true;
            }
        } else if (m == 0xC4) {                                                       // DHT
            size_t i = 0;
            while (i + 17 <= sl) {
                const int tc = s[i] >> 4, th = s[i] & 15; ++i;
                if (th > 3 || tc > 1) return BF_SENS_ERR_FORMAT;
                Huff& h = tc ? hac[th] : hdc[th];
                int total = 0; h.bits[0] = 0;
                for (int l = 1; l <= 16; ++l) { h.bits[l] = s[i + l - 1]; total += h.bits[l]; }
                i += 16;
                if (total > 256 || i + total > sl) return BF_SENS_ERR_FORMAT;
                memcpy(h.vals, s + i, total); i += total;
                huff_build(h);
            }
        } else if (m == 0xC0 || m == 0xC1 || m == 0xC2) {                             // SOF0 / SOF1: sequential DCT; SOF2: progressive DCT; Huffman
            progressive = m == 0xC2;
            if (sl < 6 || s[0] != 8) return BF_SENS_ERR_UNSUPPORTED;
            height = (s[1] << 8) | s[2]; width = (s[3] << 8) | s[4];
            const int nf = s[5];
            if (width == 0 || height == 0 || !(nf == 1 || nf == 3) || sl < (size_t)(6 + 3 * nf)) return BF_SENS_ERR_UNSUPPORTED;
            if ((1 << 30) / width / nf < height) return BF_SENS_ERR_FORMAT;                 // "too large" in the reference's decoder
            comps.resize(nf);
            for (int c = 0; c < nf; ++c) {
                comps[c].id = s[6 + 3 * c]; comps[c].h = s[7 + 3 * c] >> 4; comps[c].v = s[7 + 3 * c] & 15; comps[c].tq = s[8 + 3 * c];
                if (comps[c].h < 1 || comps[c].h > 2 || comps[c].v < 1 || comps[c].v > 2 || comps[c].tq > 3) return BF_SENS_ERR_UNSUPPORTED;
                hmax = comps[c].h > hmax ? comps[c].h : hmax; vmax = comps[c].v > vmax ? comps[c].v : vmax;
            }
            const int mcux = (width + 8 * hmax - 1) / (8 * hmax), mcuy = (height + 8 * vmax - 1) / (8 * vmax);
            for (Comp& c : comps) { c.bw = mcux * c.h; c.bh = mcuy * c.v; c.stride = (size_t)c.bw * 8; c.plane.assign(c.stride * c.bh * 8, 0); }
            haveFrame = true;
            *W = (uint32_t)width; *H = (uint32_t)height;
            if (!rgb) return BF_SENS_OK;
            if (progressive) for (Comp& c : comps) c.coeff.assign((size_t)c.bw * c.bh * 64, 0);
        } else if (m >= 0xC3 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC) return BF_SENS_ERR_UNSUPPORTED;    // lossless, hierarchical, arithmetic
        else if (m == 0xDD) { if (sl < 2) return BF_SENS_ERR_FORMAT; restart = (s[0] << 8) | s[1]; }
        else if (m == 0xDA) {                                                         // SOS + entropy-coded segment
            if (!haveFrame || sl < 1) return BF_SENS_ERR_FORMAT;
            const int ns = s[0];
            if (ns < 1 || ns > (int)comps.size() || sl < (size_t)(1 + 2 * ns + 3)) return BF_SENS_ERR_FORMAT;
            std::vector<Comp*> sc;
            for (int k = 0; k < ns; ++k) {
                Comp* c = nullptr;
                for (Comp& q : comps) if (q.id == s[1 + 2 * k]) c = &q;
                if (!c) return BF_SENS_ERR_FORMAT;
                c->td = s[2 + 2 * k] >> 4; c->ta = s[2 + 2 * k] & 15; c->pred = 0;
                if (c->td > 3 || c->ta > 3 || !qset[c->tq]) return BF_SENS_ERR_FORMAT;
                if (!progressive && (!hdc[c->td].set || !hac[c->ta].set)) return BF_SENS_ERR_FORMAT;
                sc.push_back(c);
            }
            const int ss = s[1 + 2 * ns], se = s[2 + 2 * ns], ah = s[3 + 2 * ns] >> 4, al = s[3 + 2 * ns] & 15;
            if (!progressive && (ss != 0 || se != 63)) return BF_SENS_ERR_UNSUPPORTED;
            if (progressive) {
                if (ss > 63 || se > 63 || ss > se || ah > 13 || al > 13 || (ss != 0 && ns != 1)) return BF_SENS_ERR_FORMAT;
                for (Comp* c : sc) if (ss == 0 ? (ah == 0 && !hdc[c->td].set) : !hac[c->ta].set) return BF_SENS_ERR_FORMAT;
            }
            int eobRun = 0;
            BitReader br; br.p = d + at + len; br.end = d + n;
            // an interleaved scan walks MCUs of h x v blocks per component; a single-component scan walks that component's own blocks (T.81 A.2.2 / A.2.3)
            const bool inter = ns > 1;
            const int mcux = inter ? (width + 8 * hmax - 1) / (8 * hmax) : ((width * sc[0]->h + hmax - 1) / hmax + 7) / 8;
            const int mcuy = inter ? (height + 8 * vmax - 1) / (8 * vmax) : ((height * sc[0]->v + vmax - 1) / vmax + 7) / 8;
            int untilRestart = restart, nextRst = 0;
            for (int my = 0; my < mcuy; ++my)
                for (int mx = 0; mx < mcux; ++mx) {
                    if (restart && untilRestart == 0) {                               // RSTn: byte-align, skip the marker, reset the predictors
                        br.reset();
                        while (br.p + 1 < br.end && !(br.p[0] == 0xFF && br.p[1] >= 0xD0 && br.p[1] <= 0xD7)) ++br.p;
                        if (br.p + 1 < br.end) br.p += 2;
                        nextRst = (nextRst + 1) & 7;
                        for (Comp* c : sc) c->pred = 0;
                        eobRun = 0;
                        untilRestart = restart;
                    }
                    for (Comp* c : sc) {
                        const int nbx = inter ? c->h : 1, nby = inter ? c->v : 1;
                        for (int by = 0; by < nby; ++by)
                            for (int bx = 0; bx < nbx; ++bx) {
                                if (progressive) {
                                    const int gx = (inter ? mx * c->h : mx) + bx, gy = (inter ? my * c->v : my) + by;
                                    if (gx >= c->bw || gy >= c->bh) return BF_SENS_ERR_FORMAT;
                                    const int rc = prog_block(br, &hdc[c->td], &hac[c->ta], c, c->coeff.data() + ((size_t)gy * c->bw + gx) * 64, ss, se, ah, al, eobRun);
                                    if (rc) return rc;
                                    continue;
                                }
                                short blk[64]; for (int k = 0; k < 64; ++k) blk[k] = 0;
                                const int t = huff_decode(br, hdc[c->td]);
                                if (t < 0 || t > 11) return BF_SENS_ERR_FORMAT;
                                c->pred += extend(br.bits(t), t);
                                blk[0] = (short)(c->pred * (int)qt[c->tq][0]);
                                for (int k = 1; k < 64;) {
                                    const int rs = huff_decode(br, hac[c->ta]);
                                    if (rs < 0) return BF_SENS_ERR_FORMAT;
                                    const int r = rs >> 4, sz = rs & 15;
                                    if (sz == 0) { if (r == 15) { k += 16; continue; } break; }
                                    k += r;
                                    if (k > 63) return BF_SENS_ERR_FORMAT;
                                    blk[kZigzag[k]] = (short)(extend(br.bits(sz), sz) * (int)qt[c->tq][k]);
                                    ++k;
                                }
                                const int gx = (inter ? mx * c->h : mx) + bx, gy = (inter ? my * c->v : my) + by;
                                if (gx < c->bw && gy < c->bh) idct_block(blk, c->plane.data() + (size_t)gy * 8 * c->stride + (size_t)gx * 8, c->stride);
                            }
                    }
                    if (restart) --untilRestart;
                }
            decoded = true;
            at = (size_t)(br.p - d);
            continue;
        }
        at += len;
    }
    if (!haveFrame || !decoded) return BF_SENS_ERR_FORMAT;
    if (progressive)                                                                   // all scans are in: de-quantise and transform the blocks under the image
        for (Comp& c : comps) {
            const int nbx = ((width * c.h + hmax - 1) / hmax + 7) / 8, nby = ((height * c.v + vmax - 1) / vmax + 7) / 8;
            for (int gy = 0; gy < nby; ++gy)
                for (int gx = 0; gx < nbx; ++gx) {
                    short* blk = c.coeff.data() + ((size_t)gy * c.bw + gx) * 64;
                    for (int k = 0; k < 64; ++k) blk[kZigzag[k]] = (short)(blk[kZigzag[k]] * (int)qt[c.tq][k]);
                    idct_block(blk, c.plane.data() + (size_t)gy * 8 * c.stride + (size_t)gx * 8, c.stride);
                }
        }
    // ---- up-sample and convert, row by row, with the reference decoder's arithmetic (stb_image v2.08: load_jpeg_image and its resample_row_* / YCbCr kernels):
    // vertically 3/4 of the nearer chroma row + 1/4 of the farther one, horizontally the same triangle on those sums; fixed-point JFIF conversion ----
    std::vector<std::vector<uint8_t>> line(comps.size());
    for (size_t ci = 0; ci < comps.size(); ++ci) line[ci].assign((size_t)width + 8, 0);
    for (int y = 0; y < height; ++y) {
        const uint8_t* row[3] = { nullptr, nullptr, nullptr };
        for (size_t ci = 0; ci < comps.size(); ++ci) {
            Comp& c = comps[ci];
            const int hs = hmax / c.h, vs = vmax / c.v;
            const int lores = (width + hs - 1) / hs;                                          // samples of this component under the image's width
            const int ch = (height * c.v + vmax - 1) / vmax;                                  // its rows under the image's height
            int iy = y / vs, fy = iy;
            if (vs == 2) { fy = (y & 1) ? (iy + 1 < ch ? iy + 1 : iy) : (iy > 0 ? iy - 1 : iy); }
            const uint8_t* nr = c.plane.data() + (size_t)iy * c.stride; const uint8_t* fr = c.plane.data() + (size_t)fy * c.stride;
            uint8_t* o = line[ci].data();
            if (hs == 1 && vs == 1) { row[ci] = nr; continue; }
            if (hs == 1) { for (int i = 0; i < lores; ++i) o[i] = (uint8_t)((3 * nr[i] + fr[i] + 2) >> 2); }
            else if (vs == 1) {                                                               // two samples per input sample along the row
                if (lores == 1) o[0] = o[1] = nr[0];
                else {
                    o[0] = nr[0]; o[1] = (uint8_t)((nr[0] * 3 + nr[1] + 2) >> 2);
                    int i = 1;
                    for (; i < lores - 1; ++i) { const int n3 = 3 * nr[i] + 2; o[2 * i] = (uint8_t)((n3 + nr[i - 1]) >> 2); o[2 * i + 1] = (uint8_t)((n3 + nr[i + 1]) >> 2); }
                    o[2 * i] = (uint8_t)((nr[lores - 2] * 3 + nr[lores - 1] + 2) >> 2);          // as the reference's decoder weights its last pair
                    o[2 * i + 1] = nr[lores - 1];
                }
            } else {
                if (lores == 1) o[0] = o[1] = (uint8_t)((3 * nr[0] + fr[0] + 2) >> 2);
                else {
                    int t1 = 3 * nr[0] + fr[0];
                    o[0] = (uint8_t)((t1 + 2) >> 2);
                    for (int i = 1; i < lores; ++i) { const int t0 = t1; t1 = 3 * nr[i] + fr[i]; o[2 * i - 1] = (uint8_t)((3 * t0 + t1 + 8) >> 4); o[2 * i] = (uint8_t)((3 * t1 + t0 + 8) >> 4); }
                    o[2 * lores - 1] = (uint8_t)((t1 + 2) >> 2);
                }
            }
            row[ci] = o;
        }
        uint8_t* o = rgb + (size_t)3 * width * y;
        if (comps.size() == 1) { for (int x = 0; x < width; ++x) o[3 * x] = o[3 * x + 1] = o[3 * x + 2] = row[0][x]; continue; }
#define BF_FIX(x) (((int)((x) * 4096.0f + 0.5f)) << 8)
        for (int x = 0; x < width; ++x) {
            const int yf = (row[0][x] << 20) + (1 << 19), cb = row[1][x] - 128, cr = row[2][x] - 128;
            int r = yf + cr * BF_FIX(1.40200f);
            int g = yf + (cr * -BF_FIX(0.71414f)) + ((cb * -BF_FIX(0.34414f)) & 0xffff0000);
            int b = yf + cb * BF_FIX(1.77200f);
            r >>= 20; g >>= 20; b >>= 20;
            o[3 * x] = clamp_u8(r); o[3 * x + 1] = clamp_u8(g); o[3 * x + 2] = clamp_u8(b);
        }
#undef BF_FIX
    }
    return BF_SENS_OK;
}

// ---- the container ------------------------------------------------------------------------------------------------------------------------------
struct FrameRec { uint64_t at; float pose[16]; uint64_t tsColor, tsDepth, colorBytes, depthBytes; };
template <class T> bool rd(FILE* f, T* v, size_t n = 1) { return fread(v, sizeof(T), n, f) == n; }
template <class T> bool wr(FILE* f, const T* v, size_t n = 1) { return fwrite(v, sizeof(T), n, f) == n; }

}  // namespace

struct BFSensReader { FILE* f = nullptr; BFSensHeader h; std::vector<FrameRec> frames; std::vector<uint8_t> cbuf, dbuf, tmp; };
struct BFSensWriter { FILE* f = nullptr; BFSensHeader h; long numFramesPos = 0; uint64_t count = 0; std::vector<uint8_t> zbuf; };

BF_API const char* bfSensErrorString(int code) {
    switch (code) {
        case BF_SENS_OK: return "ok"; case BF_SENS_ERR_IO: return "i/o error"; case BF_SENS_ERR_FORMAT: return "malformed data";
        case BF_SENS_ERR_UNSUPPORTED: return "unsupported variant (lossless / arithmetic-coded JPEG, OCCI depth, 16-bit PNG, ...)";
        case BF_SENS_ERR_RANGE: return "frame index out of range"; case BF_SENS_ERR_ARGUMENT: return "invalid argument";
    }
    return "unknown";
}
BF_API int bfSensDecodeJpeg(const uint8_t* data, size_t bytes, uint8_t* rgb, uint32_t* width, uint32_t* height) {
    if (!data || !width || !height) return BF_SENS_ERR_ARGUMENT;
    try { return decode_jpeg(data, bytes, rgb, width, height); } catch (const std::exception&) { return BF_SENS_ERR_FORMAT; }
}
BF_API int bfSensDecodePng(const uint8_t* data, size_t bytes, uint8_t* rgb, uint32_t* width, uint32_t* height) {
    if (!data || !width || !height) return BF_SENS_ERR_ARGUMENT;
    try { return decode_png(data, bytes, rgb, width, height); } catch (const std::exception&) { return BF_SENS_ERR_FORMAT; }
}

BF_API int bfSensOpen(const char* path, BFSensReader** out, BFSensHeader* header) {
    if (!path || !out) return BF_SENS_ERR_ARGUMENT;
    FILE* f = fopen(path, "rb");
    if (!f) return BF_SENS_ERR_IO;
    BFSensReader* r = new BFSensReader(); r->f = f;
    BFSensHeader& h = r->h; memset(&h, 0, sizeof(h));
    int rc = BF_SENS_ERR_FORMAT;
    uint64_t fileSize = 0;
    if (fseeko(f, 0, SEEK_END) == 0) { fileSize = (uint64_t)ftello(f); fseeko(f, 0, SEEK_SET); }
    try {
    do {
        uint64_t strLen = 0;
        if (!rd(f, &h.version) || h.version != 4 || !rd(f, &strLen) || strLen > (1u << 20)) break;                 // sensorData.h:1194-1199
        std::string name(strLen, '\0');
        if (strLen && !rd(f, &name[0], strLen)) break;
        strncpy(h.sensorName, name.c_str(), sizeof(h.sensorName) - 1);
        if (!rd(f, h.colorIntrinsic, 16) || !rd(f, h.colorExtrinsic, 16) || !rd(f, h.depthIntrinsic, 16) || !rd(f, h.depthExtrinsic, 16)) break;
        if (!rd(f, &h.colorCompression) || !rd(f, &h.depthCompression) || !rd(f, &h.colorWidth) || !rd(f, &h.colorHeight) || !rd(f, &h.depthWidth) || !rd(f, &h.depthHeight) ||
            !rd(f, &h.depthShift) || !rd(f, &h.numFrames)) break;
        if (h.colorWidth > (1u << 15) || h.colorHeight > (1u << 15) || h.depthWidth > (1u << 15) || h.depthHeight > (1u << 15)) break;      // image sizes no sensor has
        if (h.numFrames > fileSize / 96) break;                                                                   // a frame record is at least 96 bytes: a corrupt count must not drive an allocation
        r->frames.resize((size_t)h.numFrames);
        bool ok = true;
        for (FrameRec& fr : r->frames) {                                                                          // RGBDFrame::loadFromFile, :686-700
            if (!rd(f, fr.pose, 16) || !rd(f, &fr.tsColor) || !rd(f, &fr.tsDepth) || !rd(f, &fr.colorBytes) || !rd(f, &fr.depthBytes)) { ok = false; break; }
            fr.at = (uint64_t)ftello(f);
            if (fr.colorBytes > fileSize || fr.depthBytes > fileSize || fr.at + fr.colorBytes + fr.depthBytes > fileSize) { ok = false; break; }      // payloads lie inside the file
            if (fseeko(f, (off_t)(fr.colorBytes + fr.depthBytes), SEEK_CUR) != 0) { ok = false; break; }
        }
        if (!ok) break;
        if (!rd(f, &h.numIMUFrames)) h.numIMUFrames = 0;                                                          // older writers stop after the frames
        rc = BF_SENS_OK;
    } while (0);
    } catch (const std::exception&) { rc = BF_SENS_ERR_FORMAT; }                                                     // e.g. bad_alloc on a size a corrupt file claims
    if (rc) { fclose(f); delete r; return rc; }
    if (header) *header = h;
    *out = r;
    return BF_SENS_OK;
}
BF_API void bfSensClose(BFSensReader* r) { if (r) { if (r->f) fclose(r->f); delete r; } }

static int read_payload(BFSensReader* r, uint64_t index, uint16_t* depth, uint8_t* rgb) {
    if (index >= r->frames.size()) return BF_SENS_ERR_RANGE;
    const FrameRec& fr = r->frames[(size_t)index]; const BFSensHeader& h = r->h;
    if (fseeko(r->f, (off_t)fr.at, SEEK_SET) != 0) return BF_SENS_ERR_IO;
    r->cbuf.resize((size_t)fr.colorBytes); r->dbuf.resize((size_t)fr.depthBytes);
    if (fr.colorBytes && !rd(r->f, r->cbuf.data(), r->cbuf.size())) return BF_SENS_ERR_IO;
    if (fr.depthBytes && !rd(r->f, r->dbuf.data(), r->dbuf.size())) return BF_SENS_ERR_IO;
    if (depth) {
        const size_t want = (size_t)h.depthWidth * h.depthHeight * 2;
        if (h.depthCompression == BF_SENS_DEPTH_RAW_USHORT) { if (r->dbuf.size() < want) return BF_SENS_ERR_FORMAT; memcpy(depth, r->dbuf.data(), want); }
        else if (h.depthCompression == BF_SENS_DEPTH_ZLIB_USHORT) {
            const int rc = inflate_all(r->dbuf.data(), r->dbuf.size(), r->tmp, want);
            if (rc) return rc;
            if (r->tmp.size() < want) return BF_SENS_ERR_FORMAT;
            memcpy(depth, r->tmp.data(), want);
        } else return BF_SENS_ERR_UNSUPPORTED;
    }
    if (rgb) {
        const size_t want = (size_t)h.colorWidth * h.colorHeight * 3;
        if (fr.colorBytes == 0) memset(rgb, 0, want);
        else if (h.colorCompression == BF_SENS_COLOR_RAW) { if (r->cbuf.size() < want) return BF_SENS_ERR_FORMAT; memcpy(rgb, r->cbuf.data(), want); }
        else {
            uint32_t w = 0, hh = 0;
            int rc = h.colorCompression == BF_SENS_COLOR_JPEG ? decode_jpeg(r->cbuf.data(), r->cbuf.size(), nullptr, &w, &hh) : (h.colorCompression == BF_SENS_COLOR_PNG ? decode_png(r->cbuf.data(), r->cbuf.size(), nullptr, &w, &hh) : BF_SENS_ERR_UNSUPPORTED);
            if (rc) return rc;
            if (w != h.colorWidth || hh != h.colorHeight) return BF_SENS_ERR_FORMAT;
            rc = h.colorCompression == BF_SENS_COLOR_JPEG ? decode_jpeg(r->cbuf.data(), r->cbuf.size(), rgb, &w, &hh) : decode_png(r->cbuf.data(), r->cbuf.size(), rgb, &w, &hh);
            if (rc) return rc;
        }
    }
    return BF_SENS_OK;
}
BF_API int bfSensReadFrameRaw(BFSensReader* r, uint64_t index, uint16_t* depth, uint8_t* colorRGB) {
    if (!r) return BF_SENS_ERR_ARGUMENT;
    try { return read_payload(r, index, depth, colorRGB); } catch (const std::exception&) { return BF_SENS_ERR_FORMAT; }
}

BF_API int bfSensReadFrame(BFSensReader* r, uint64_t index, float* depthMetres, uint8_t* colorRGBX, float* cameraToWorld, uint64_t* timeStamps) {
    if (!r) return BF_SENS_ERR_ARGUMENT;
    if (index >= r->frames.size()) return BF_SENS_ERR_RANGE;
    const BFSensHeader& h = r->h;
    const size_t nd = (size_t)h.depthWidth * h.depthHeight, nc = (size_t)h.colorWidth * h.colorHeight;
    std::vector<uint16_t> d; std::vector<uint8_t> c;
    int rc;
    try {
        d.resize(depthMetres ? nd : 0); c.resize(colorRGBX ? nc * 3 : 0);
        rc = read_payload(r, index, depthMetres ? d.data() : nullptr, colorRGBX ? c.data() : nullptr);
    } catch (const std::exception&) { rc = BF_SENS_ERR_FORMAT; }
    if (rc) return rc;
    if (depthMetres)                                                       // SensorDataReader::processDepth, FL/SensorDataReader.cpp:104-107
        for (size_t i = 0; i < nd; ++i) depthMetres[i] = d[i] == 0 ? -INFINITY : (float)d[i] / h.depthShift;
    if (colorRGBX)                                                         // vec4uc(vec3uc): (r, g, b, 1), :112-114 with core-math/vec4.h:42-47
        for (size_t i = 0; i < nc; ++i) { colorRGBX[4 * i] = c[3 * i]; colorRGBX[4 * i + 1] = c[3 * i + 1]; colorRGBX[4 * i + 2] = c[3 * i + 2]; colorRGBX[4 * i + 3] = 1; }
    const FrameRec& fr = r->frames[(size_t)index];
    if (cameraToWorld) memcpy(cameraToWorld, fr.pose, 64);
    if (timeStamps) { timeStamps[0] = fr.tsColor; timeStamps[1] = fr.tsDepth; }
    return BF_SENS_OK;
}

BF_API int bfSensCreate(const char* path, const BFSensHeader* header, BFSensWriter** out) {
    if (!path || !header || !out) return BF_SENS_ERR_ARGUMENT;
    if (header->colorCompression != BF_SENS_COLOR_RAW || !(header->depthCompression == BF_SENS_DEPTH_RAW_USHORT || header->depthCompression == BF_SENS_DEPTH_ZLIB_USHORT)) return BF_SENS_ERR_UNSUPPORTED;
    FILE* f = fopen(path, "wb");
    if (!f) return BF_SENS_ERR_IO;
    BFSensWriter* w = new BFSensWriter(); w->f = f; w->h = *header; w->h.version = 4;
    const BFSensHeader& h = w->h;
    const uint64_t strLen = strnlen(h.sensorName, sizeof(h.sensorName)), zero = 0;
    bool ok = wr(f, &h.version) && wr(f, &strLen) && (strLen == 0 || wr(f, h.sensorName, strLen)) && wr(f, h.colorIntrinsic, 16) && wr(f, h.colorExtrinsic, 16) &&
              wr(f, h.depthIntrinsic, 16) && wr(f, h.depthExtrinsic, 16) && wr(f, &h.colorCompression) && wr(f, &h.depthCompression) && wr(f, &h.colorWidth) && wr(f, &h.colorHeight) &&
              wr(f, &h.depthWidth) && wr(f, &h.depthHeight) && wr(f, &h.depthShift);
    w->numFramesPos = ftell(f);
    ok = ok && wr(f, &zero);
    if (!ok) { fclose(f); delete w; return BF_SENS_ERR_IO; }
    *out = w;
    return BF_SENS_OK;
}
BF_API int bfSensAppendFrame(BFSensWriter* w, const uint16_t* depth, const uint8_t* colorRGB, const float* cameraToWorld, uint64_t timeStampColor, uint64_t timeStampDepth) {
    if (!w || !depth) return BF_SENS_ERR_ARGUMENT;
    const BFSensHeader& h = w->h;
    float pose[16];
    if (cameraToWorld) memcpy(pose, cameraToWorld, 64); else for (int k = 0; k < 16; ++k) pose[k] = -INFINITY;       // RGBDFrame(): "no pose" is all -inf
    const size_t nd = (size_t)h.depthWidth * h.depthHeight * 2;
    const uint8_t* dptr = reinterpret_cast<const uint8_t*>(depth); uint64_t dbytes = nd;
    if (h.depthCompression == BF_SENS_DEPTH_ZLIB_USHORT) {
        uLongf cap = compressBound((uLong)nd);
        w->zbuf.resize(cap);
        if (compress2(w->zbuf.data(), &cap, dptr, (uLong)nd, 6) != Z_OK) return BF_SENS_ERR_IO;
        dptr = w->zbuf.data(); dbytes = cap;
    }
    const uint64_t cbytes = colorRGB ? (uint64_t)h.colorWidth * h.colorHeight * 3 : 0;
    FILE* f = w->f;
    const bool ok = wr(f, pose, 16) && wr(f, &timeStampColor) && wr(f, &timeStampDepth) && wr(f, &cbytes) && wr(f, &dbytes) && (cbytes == 0 || wr(f, colorRGB, (size_t)cbytes)) && wr(f, dptr, (size_t)dbytes);
    if (!ok) return BF_SENS_ERR_IO;
    ++w->count;
    return BF_SENS_OK;
}
BF_API int bfSensFinish(BFSensWriter* w) {
    if (!w) return BF_SENS_ERR_ARGUMENT;
    const uint64_t zero = 0;
    bool ok = wr(w->f, &zero);                                             // no IMU frames
    ok = ok && fseek(w->f, w->numFramesPos, SEEK_SET) == 0 && wr(w->f, &w->count);
    ok = (fclose(w->f) == 0) && ok;
    delete w;
    return ok ? BF_SENS_OK : BF_SENS_ERR_IO;
}
