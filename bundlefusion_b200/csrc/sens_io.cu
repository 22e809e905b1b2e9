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
    while (at + 12 <= n) {
        const uint32_t len = be32(d + at); const uint8_t* tag = d + at + 4; const uint8_t* body = d + at + 8;
        if (at + 12 + (size_t)len > n) return BF_SENS_ERR_FORMAT;
        if (!memcmp(tag, "IHDR", 4)) { if (len < 13) return BF_SENS_ERR_FORMAT; w = be32(body); h = be32(body + 4); depth = body[8]; type = body[9]; interlace = body[12]; }
        else if (!memcmp(tag, "IDAT", 4)) idat.insert(idat.end(), body, body + len);
        else if (!memcmp(tag, "IEND", 4)) break;
        at += 12 + (size_t)len;
    }
    if (w == 0 || h == 0) return BF_SENS_ERR_FORMAT;
    if (depth != 8 || interlace != 0 || !(type == 0 || type == 2 || type == 4 || type == 6)) return BF_SENS_ERR_UNSUPPORTED;
    *W = w; *H = h;
    if (!rgb) return BF_SENS_OK;
    const int ch = type == 0 ? 1 : (type == 4 ? 2 : (type == 2 ? 3 : 4));
    const size_t stride = (size_t)w * ch;
    std::vector<uint8_t> raw;
    const int rc = inflate_all(idat.data(), idat.size(), raw, (stride + 1) * h);
    if (rc) return rc;
    if (raw.size() < (stride + 1) * h) return BF_SENS_ERR_FORMAT;
    std::vector<uint8_t> prev(stride, 0), cur(stride);
    for (uint32_t y = 0; y < h; ++y) {
        const uint8_t* s = raw.data() + (stride + 1) * y;
        const int f = s[0]; ++s;
        for (size_t i = 0; i < stride; ++i) {
            const int a = i >= (size_t)ch ? cur[i - ch] : 0, b = prev[i], c = i >= (size_t)ch ? prev[i - ch] : 0;
            int v = s[i];
            switch (f) { case 0: break; case 1: v += a; break; case 2: v += b; break; case 3: v += (a + b) >> 1; break; case 4: v += paeth(a, b, c); break; default: return BF_SENS_ERR_FORMAT; }
            cur[i] = (uint8_t)v;
        }
        uint8_t* o = rgb + (size_t)3 * w * y;
        for (uint32_t x = 0; x < w; ++x) {
            const uint8_t* px = &cur[(size_t)x * ch];
            if (ch <= 2) { o[3 * x] = o[3 * x + 1] = o[3 * x + 2] = px[0]; } else { o[3 * x] = px[0]; o[3 * x + 1] = px[1]; o[3 * x + 2] = px[2]; }
        }
        prev.swap(cur);
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
struct IdctTable { float c[8][8]; IdctTable() { for (int x = 0; x < 8; ++x) for (int u = 0; u < 8; ++u) c[x][u] = (float)((u == 0 ? std::sqrt(0.125) : 0.5) * std::cos((2 * x + 1) * u * M_PI / 16.0)); } };
void idct_block(const float in[64], uint8_t* out, size_t stride) {                              // s(y, x) = sum_v sum_u C(y, v) C(x, u) S(v, u), T.81 A.3.3
    static const IdctTable T;
    float tmp[64];
    for (int v = 0; v < 8; ++v) {
        const float* s = in + 8 * v;
        if (s[1] == 0 && s[2] == 0 && s[3] == 0 && s[4] == 0 && s[5] == 0 && s[6] == 0 && s[7] == 0) { const float d = s[0] * T.c[0][0]; for (int x = 0; x < 8; ++x) tmp[8 * v + x] = d; continue; }
        for (int x = 0; x < 8; ++x) { float a = 0; for (int u = 0; u < 8; ++u) a += T.c[x][u] * s[u]; tmp[8 * v + x] = a; }
    }
    for (int x = 0; x < 8; ++x)
        for (int y = 0; y < 8; ++y) {
            float a = 0;
            for (int v = 0; v < 8; ++v) a += T.c[y][v] * tmp[8 * v + x];
            const int q = (int)std::lrintf(a) + 128;
            out[(size_t)y * stride + x] = (uint8_t)(q < 0 ? 0 : (q > 255 ? 255 : q));
        }
}
struct Comp { int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0, pred = 0; int bw = 0, bh = 0; size_t stride = 0; std::vector<uint8_t> plane; };

int decode_jpeg(const uint8_t* d, size_t n, uint8_t* rgb, uint32_t* W, uint32_t* H) {
    if (n < 4 || d[0] != 0xFF || d[1] != 0xD8) return BF_SENS_ERR_FORMAT;
    uint16_t qt[4][64]; bool qset[4] = { false, false, false, false };
    Huff hdc[4], hac[4];
    std::vector<Comp> comps;
    int width = 0, height = 0, hmax = 1, vmax = 1, restart = 0, adobeTransform = -1;
    bool haveFrame = false, decoded = false;
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
        } else if (m == 0xC0 || m == 0xC1) {                                          // SOF0 / SOF1: sequential DCT, Huffman
            if (sl < 6 || s[0] != 8) return BF_SENS_ERR_UNSUPPORTED;
            height = (s[1] << 8) | s[2]; width = (s[3] << 8) | s[4];
            const int nf = s[5];
            if (width == 0 || height == 0 || !(nf == 1 || nf == 3) || sl < (size_t)(6 + 3 * nf)) return BF_SENS_ERR_UNSUPPORTED;
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
        } else if (m >= 0xC2 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC) return BF_SENS_ERR_UNSUPPORTED;    // progressive, lossless, arithmetic
        else if (m == 0xDD) { if (sl < 2) return BF_SENS_ERR_FORMAT; restart = (s[0] << 8) | s[1]; }
        else if (m == 0xEE && sl >= 12 && !memcmp(s, "Adobe", 5)) adobeTransform = s[11];
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
                if (c->td > 3 || c->ta > 3 || !hdc[c->td].set || !hac[c->ta].set || !qset[c->tq]) return BF_SENS_ERR_FORMAT;
                sc.push_back(c);
            }
            if (s[1 + 2 * ns] != 0 || s[2 + 2 * ns] != 63) return BF_SENS_ERR_UNSUPPORTED;
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
                        untilRestart = restart;
                    }
                    for (Comp* c : sc) {
                        const int nbx = inter ? c->h : 1, nby = inter ? c->v : 1;
                        for (int by = 0; by < nby; ++by)
                            for (int bx = 0; bx < nbx; ++bx) {
                                float blk[64]; for (int k = 0; k < 64; ++k) blk[k] = 0.0f;
                                const int t = huff_decode(br, hdc[c->td]);
                                if (t < 0 || t > 11) return BF_SENS_ERR_FORMAT;
                                c->pred += extend(br.bits(t), t);
                                blk[0] = (float)(c->pred * (int)qt[c->tq][0]);
                                for (int k = 1; k < 64;) {
                                    const int rs = huff_decode(br, hac[c->ta]);
                                    if (rs < 0) return BF_SENS_ERR_FORMAT;
                                    const int r = rs >> 4, sz = rs & 15;
                                    if (sz == 0) { if (r == 15) { k += 16; continue; } break; }
                                    k += r;
                                    if (k > 63) return BF_SENS_ERR_FORMAT;
                                    blk[kZigzag[k]] = (float)(extend(br.bits(sz), sz) * (int)qt[c->tq][k]);
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
    // ---- up-sample to full resolution (triangle filter for 2:1, as libjpeg's "fancy" up-sampling and stb_image's default), colour conversion ----
    const size_t fw = (size_t)comps[0].bw * 8 / comps[0].h * hmax;                    // full width of the MCU-padded image
    std::vector<std::vector<uint8_t>> full(comps.size());
    for (size_t ci = 0; ci < comps.size(); ++ci) {
        Comp& c = comps[ci];
        const int sx = hmax / c.h, sy = vmax / c.v;
        const size_t cw = c.stride, chh = (size_t)c.bh * 8;
        const size_t uw = ((size_t)width * c.h + hmax - 1) / hmax, uh = ((size_t)height * c.v + vmax - 1) / vmax;      // the component's true (un-padded) size
        if (sx == 1 && sy == 1) { full[ci].swap(c.plane); continue; }
        std::vector<uint8_t>& o = full[ci];
        o.assign(fw * chh * sy, 0);
        std::vector<int> colsum(cw);
        for (size_t oy = 0; oy < uh * sy; ++oy) {
            const size_t iy = oy / sy;
            // vertical: 3/4 of the nearer row + 1/4 of the farther one (replicated at the edges); sy == 1: the row itself, scaled by 4 to share the code
            size_t far = iy;
            if (sy == 2) { if (oy & 1) far = iy + 1 < uh ? iy + 1 : iy; else far = iy > 0 ? iy - 1 : iy; }
            const uint8_t* rn = c.plane.data() + iy * cw; const uint8_t* rf = c.plane.data() + far * cw;
            for (size_t x = 0; x < uw; ++x) colsum[x] = sy == 2 ? 3 * rn[x] + rf[x] : 4 * rn[x];
            uint8_t* orow = o.data() + oy * fw;
            if (sx == 1) { for (size_t x = 0; x < uw; ++x) orow[x] = (uint8_t)((colsum[x] + 2) >> 2); continue; }
            for (size_t x = 0; x < uw; ++x) {
                const int t = colsum[x], l = x > 0 ? colsum[x - 1] : t, r = x + 1 < uw ? colsum[x + 1] : t;
                if (sy == 2) { orow[2 * x] = (uint8_t)((3 * t + l + 8) >> 4); orow[2 * x + 1] = (uint8_t)((3 * t + r + 7) >> 4); }                 // h2v2: 16ths of the 2-D triangle
                else { orow[2 * x] = (uint8_t)((3 * (t >> 2) + (l >> 2) + 1) >> 2); orow[2 * x + 1] = (uint8_t)((3 * (t >> 2) + (r >> 2) + 2) >> 2); }   // h2v1: quarters, libjpeg's rounding pattern
            }
        }
    }
    const size_t s0 = comps[0].h == hmax && comps[0].v == vmax ? comps[0].stride : fw;
    for (int y = 0; y < height; ++y)
        for (int x = 0; x < width; ++x) {
            uint8_t* o = rgb + 3 * ((size_t)y * width + x);
            const int Y = full[0][(size_t)y * s0 + x];
            if (comps.size() == 1) { o[0] = o[1] = o[2] = (uint8_t)Y; continue; }
            const size_t s1 = comps[1].h == hmax && comps[1].v == vmax ? comps[1].stride : fw, s2 = comps[2].h == hmax && comps[2].v == vmax ? comps[2].stride : fw;
            const int cb = full[1][(size_t)y * s1 + x], cr = full[2][(size_t)y * s2 + x];
            if (adobeTransform == 0) { o[0] = (uint8_t)Y; o[1] = (uint8_t)cb; o[2] = (uint8_t)cr; continue; }              // Adobe marker: components are RGB
            const float r = Y + 1.402f * (cr - 128), g = Y - 0.344136f * (cb - 128) - 0.714136f * (cr - 128), b = Y + 1.772f * (cb - 128);      // JFIF (ITU-T T.871)
            const int ri = (int)std::lrintf(r), gi = (int)std::lrintf(g), bi = (int)std::lrintf(b);
            o[0] = (uint8_t)(ri < 0 ? 0 : (ri > 255 ? 255 : ri)); o[1] = (uint8_t)(gi < 0 ? 0 : (gi > 255 ? 255 : gi)); o[2] = (uint8_t)(bi < 0 ? 0 : (bi > 255 ? 255 : bi));
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
        case BF_SENS_ERR_UNSUPPORTED: return "unsupported variant (progressive JPEG, OCCI depth, 16-bit / interlaced PNG, ...)";
        case BF_SENS_ERR_RANGE: return "frame index out of range"; case BF_SENS_ERR_ARGUMENT: return "invalid argument";
    }
    return "unknown";
}
BF_API int bfSensDecodeJpeg(const uint8_t* data, size_t bytes, uint8_t* rgb, uint32_t* width, uint32_t* height) {
    if (!data || !width || !height) return BF_SENS_ERR_ARGUMENT;
    return decode_jpeg(data, bytes, rgb, width, height);
}
BF_API int bfSensDecodePng(const uint8_t* data, size_t bytes, uint8_t* rgb, uint32_t* width, uint32_t* height) {
    if (!data || !width || !height) return BF_SENS_ERR_ARGUMENT;
    return decode_png(data, bytes, rgb, width, height);
}

BF_API int bfSensOpen(const char* path, BFSensReader** out, BFSensHeader* header) {
    if (!path || !out) return BF_SENS_ERR_ARGUMENT;
    FILE* f = fopen(path, "rb");
    if (!f) return BF_SENS_ERR_IO;
    BFSensReader* r = new BFSensReader(); r->f = f;
    BFSensHeader& h = r->h; memset(&h, 0, sizeof(h));
    int rc = BF_SENS_ERR_FORMAT;
    do {
        uint64_t strLen = 0;
        if (!rd(f, &h.version) || h.version != 4 || !rd(f, &strLen) || strLen > (1u << 20)) break;                 // sensorData.h:1194-1199
        std::string name(strLen, '\0');
        if (strLen && !rd(f, &name[0], strLen)) break;
        strncpy(h.sensorName, name.c_str(), sizeof(h.sensorName) - 1);
        if (!rd(f, h.colorIntrinsic, 16) || !rd(f, h.colorExtrinsic, 16) || !rd(f, h.depthIntrinsic, 16) || !rd(f, h.depthExtrinsic, 16)) break;
        if (!rd(f, &h.colorCompression) || !rd(f, &h.depthCompression) || !rd(f, &h.colorWidth) || !rd(f, &h.colorHeight) || !rd(f, &h.depthWidth) || !rd(f, &h.depthHeight) ||
            !rd(f, &h.depthShift) || !rd(f, &h.numFrames)) break;
        if (h.numFrames > (1ull << 32)) break;
        r->frames.resize((size_t)h.numFrames);
        bool ok = true;
        for (FrameRec& fr : r->frames) {                                                                          // RGBDFrame::loadFromFile, :686-700
            if (!rd(f, fr.pose, 16) || !rd(f, &fr.tsColor) || !rd(f, &fr.tsDepth) || !rd(f, &fr.colorBytes) || !rd(f, &fr.depthBytes)) { ok = false; break; }
            fr.at = (uint64_t)ftello(f);
            if (fseeko(f, (off_t)(fr.colorBytes + fr.depthBytes), SEEK_CUR) != 0) { ok = false; break; }
        }
        if (!ok) break;
        if (!rd(f, &h.numIMUFrames)) h.numIMUFrames = 0;                                                          // older writers stop after the frames
        rc = BF_SENS_OK;
    } while (0);
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
BF_API int bfSensReadFrameRaw(BFSensReader* r, uint64_t index, uint16_t* depth, uint8_t* colorRGB) { return r ? read_payload(r, index, depth, colorRGB) : BF_SENS_ERR_ARGUMENT; }

BF_API int bfSensReadFrame(BFSensReader* r, uint64_t index, float* depthMetres, uint8_t* colorRGBX, float* cameraToWorld, uint64_t* timeStamps) {
    if (!r) return BF_SENS_ERR_ARGUMENT;
    if (index >= r->frames.size()) return BF_SENS_ERR_RANGE;
    const BFSensHeader& h = r->h;
    const size_t nd = (size_t)h.depthWidth * h.depthHeight, nc = (size_t)h.colorWidth * h.colorHeight;
    std::vector<uint16_t> d(depthMetres ? nd : 0); std::vector<uint8_t> c(colorRGBX ? nc * 3 : 0);
    const int rc = read_payload(r, index, depthMetres ? d.data() : nullptr, colorRGBX ? c.data() : nullptr);
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
