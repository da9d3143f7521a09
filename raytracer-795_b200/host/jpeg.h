// JPEG decoder (8-bit, Huffman; sequential SOF0 / SOF1 and progressive SOF2) for image textures - the reference reads them
// through libjpeg (Texture::ReadJPG, src/Texture.cpp:191-240: jpeg_read_header + jpeg_start_decompress with the library
// defaults + jpeg_read_scanlines), which is not available here, so the decode is re-implemented with libjpeg's default
// arithmetic: the "islow" integer inverse DCT, "fancy" (triangle-filter) h2v1 / h2v2 chroma upsampling, and the fixed-point
// YCbCr -> RGB tables - so the texels equal what the reference would see (checked bit for bit against another libjpeg-based
// decoder in tests/test_host_loader.py).  Sequential (SOF0 / SOF1) and progressive (SOF2) Huffman files are decoded;
// arithmetic-coded, lossless, 12-bit and CMYK files are rejected with a message.  Header-only, no dependencies.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace rt795 {
namespace jpeg_detail {

struct Huff {
    uint8_t bits[17] = {0};
    uint8_t vals[256] = {0};
    int mincode[17], maxcode[18], valptr[17];
    bool present = false;
    void Build() {
        int code = 0, k = 0;
        for (int l = 1; l <= 16; l++) {
            valptr[l] = k; mincode[l] = code;
            code += bits[l]; k += bits[l];
            maxcode[l] = bits[l] ? code - 1 : -1;
            code <<= 1;
        }
        maxcode[17] = 0x7fffffff;
        present = true;
    }
};

struct Component {
    int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0;
    int wblocks = 0, hblocks = 0;          // allocated blocks (padded to whole MCUs)
    int dw = 0, dh = 0;                    // downsampled_width / downsampled_height (real samples)
    int pred = 0;
    std::vector<uint8_t> plane;            // wblocks*8 x hblocks*8
    std::vector<int16_t> coefs;            // progressive files: wblocks*hblocks blocks of 64 (natural order), refined scan by scan
};

struct BitReader {
    const uint8_t* p; const uint8_t* end;
    uint32_t acc = 0; int nbits = 0;
    bool hit_marker = false;
    void Fill() {
        while (nbits <= 24) {
            int b = 0;
            if (!hit_marker && p < end) {
                b = *p;
                if (b == 0xff) {
                    if (p + 1 < end && p[1] == 0x00) p += 2;              // stuffed zero
                    else { hit_marker = true; b = 0; }                     // a marker: feed zeros, leave p on it
                } else p++;
            }
            acc |= (uint32_t)b << (24 - nbits);
            nbits += 8;
        }
    }
    int Get(int n) {
        if (n == 0) return 0;
        if (nbits < n) Fill();
        int v = (int)(acc >> (32 - n));
        acc <<= n; nbits -= n;
        return v;
    }
    void Reset() { acc = 0; nbits = 0; hit_marker = false; }
};

inline int Extend(int v, int s) { return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v; }

inline bool DecodeSym(BitReader& br, const Huff& h, int* out) {
    int code = 0;
    for (int l = 1; l <= 16; l++) {
        code = (code << 1) | br.Get(1);
        if (h.maxcode[l] >= 0 && code <= h.maxcode[l] && code >= h.mincode[l]) { *out = h.vals[h.valptr[l] + code - h.mincode[l]]; return true; }
    }
    return false;
}

// jidctint.c (libjpeg "islow"): CONST_BITS 13, PASS1_BITS 2
inline int Descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }
inline uint8_t Clamp8(int x) { return (uint8_t)(x < 0 ? 0 : (x > 255 ? 255 : x)); }
inline void IdctIslow(const int* in /*dequantised, natural order*/, uint8_t* out, int stride) {
    const int C0_298 = 2446, C0_390 = 3196, C0_541 = 4433, C0_765 = 6270, C0_899 = 7373, C1_175 = 9633, C1_501 = 12299, C1_847 = 15137, C1_961 = 16069,
              C2_053 = 16819, C2_562 = 20995, C3_072 = 25172;
    int ws[64];
    for (int pass = 0; pass < 2; pass++) {
        for (int i = 0; i < 8; i++) {
            int s0, s1, s2, s3, s4, s5, s6, s7;
            if (pass == 0) { s0 = in[i]; s1 = in[8 + i]; s2 = in[16 + i]; s3 = in[24 + i]; s4 = in[32 + i]; s5 = in[40 + i]; s6 = in[48 + i]; s7 = in[56 + i]; }
            else { const int* w = ws + 8 * i; s0 = w[0]; s1 = w[1]; s2 = w[2]; s3 = w[3]; s4 = w[4]; s5 = w[5]; s6 = w[6]; s7 = w[7]; }
            int z1 = (s2 + s6) * C0_541;
            int tmp2 = z1 + s6 * (-C1_847), tmp3 = z1 + s2 * C0_765;
            int tmp0 = (s0 + s4) << 13, tmp1 = (s0 - s4) << 13;
            int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
            int t0 = s7, t1 = s5, t2 = s3, t3 = s1;
            int y1 = t0 + t3, y2 = t1 + t2, y3 = t0 + t2, y4 = t1 + t3;
            int y5 = (y3 + y4) * C1_175;
            t0 *= C0_298; t1 *= C2_053; t2 *= C3_072; t3 *= C1_501;
            y1 *= -C0_899; y2 *= -C2_562; y3 *= -C1_961; y4 *= -C0_390;
            y3 += y5; y4 += y5;
            t0 += y1 + y3; t1 += y2 + y4; t2 += y2 + y3; t3 += y1 + y4;
            if (pass == 0) {
                const int n = 13 - 2;
                ws[i] = Descale(tmp10 + t3, n); ws[56 + i] = Descale(tmp10 - t3, n);
                ws[8 + i] = Descale(tmp11 + t2, n); ws[48 + i] = Descale(tmp11 - t2, n);
                ws[16 + i] = Descale(tmp12 + t1, n); ws[40 + i] = Descale(tmp12 - t1, n);
                ws[24 + i] = Descale(tmp13 + t0, n); ws[32 + i] = Descale(tmp13 - t0, n);
            } else {
                const int n = 13 + 2 + 3;
                uint8_t* o = out + (size_t)i * stride;
                o[0] = Clamp8(Descale(tmp10 + t3, n) + 128); o[7] = Clamp8(Descale(tmp10 - t3, n) + 128);
                o[1] = Clamp8(Descale(tmp11 + t2, n) + 128); o[6] = Clamp8(Descale(tmp11 - t2, n) + 128);
                o[2] = Clamp8(Descale(tmp12 + t1, n) + 128); o[5] = Clamp8(Descale(tmp12 - t1, n) + 128);
                o[3] = Clamp8(Descale(tmp13 + t0, n) + 128); o[4] = Clamp8(Descale(tmp13 - t0, n) + 128);
            }
        }
    }
}

const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6,  7,  14, 21, 28,
                             35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

}  // namespace jpeg_detail

// Decodes to interleaved 8-bit samples, `*ncomp` (1 or 3) per pixel, rows top to bottom - what jpeg_read_scanlines hands the
// reference.  Returns false with a message for unsupported or corrupt files.
inline bool DecodeJpeg(const std::vector<uint8_t>& file, int* width, int* height, int* ncomp, std::vector<uint8_t>* pixels, std::string* err) {
    using namespace jpeg_detail;
    auto fail = [&](const char* m) { *err = m; return false; };
    const uint8_t* p = file.data();
    const uint8_t* end = p + file.size();
    if (file.size() < 4 || p[0] != 0xff || p[1] != 0xd8) return fail("not a JPEG file (no SOI)");
    p += 2;
    uint16_t qt[4][64]; bool qt_ok[4] = {false, false, false, false};
    Huff hdc[4], hac[4];
    std::vector<Component> comps;
    int W = 0, H = 0, hmax = 1, vmax = 1, restart = 0;
    bool have_frame = false, adobe = false, progressive = false, geometry = false; int adobe_transform = -1;
    int scans = 0, mcux = 0, mcuy = 0;
    bool done = false;
    while (!done) {
        // next marker
        while (p < end && *p != 0xff) p++;
        while (p < end && *p == 0xff) p++;
        if (p >= end) { if (progressive && scans > 0) break; return fail("JPEG: no image data (missing SOS/EOI)"); }
        int m = *p++;
        if (m == 0xd9) { if (progressive && scans > 0) break; return fail("JPEG: EOI before any scan"); }
        if (m == 0x01 || (m >= 0xd0 && m <= 0xd7)) continue;
        if (p + 2 > end) return fail("JPEG: truncated segment");
        int len = (p[0] << 8) | p[1];
        if (len < 2 || p + len > end) return fail("JPEG: bad segment length");
        const uint8_t* s = p + 2; const uint8_t* se = p + len;
        p += len;
        switch (m) {
            case 0xdb:   // DQT
                while (s < se) {
                    int pq = *s >> 4, tq = *s & 15; s++;
                    if (tq > 3 || pq > 1) return fail("JPEG: bad quantisation table");
                    if (s + 64 * (pq + 1) > se) return fail("JPEG: truncated DQT");
                    for (int i = 0; i < 64; i++) { qt[tq][kZigzag[i]] = pq ? (uint16_t)((s[0] << 8) | s[1]) : s[0]; s += pq + 1; }
                    qt_ok[tq] = true;
                }
                break;
            case 0xc4:   // DHT
                while (s < se) {
                    int tc = *s >> 4, th = *s & 15; s++;
                    if (tc > 1 || th > 3 || s + 16 > se) return fail("JPEG: bad Huffman table");
                    Huff& h = tc ? hac[th] : hdc[th];
                    int total = 0;
                    h.bits[0] = 0;
                    for (int l = 1; l <= 16; l++) { h.bits[l] = s[l - 1]; total += s[l - 1]; }
                    s += 16;
                    if (total > 256 || s + total > se) return fail("JPEG: bad Huffman table");
                    memcpy(h.vals, s, total); s += total;
                    h.Build();
                }
                break;
            case 0xc2: progressive = true;   // SOF2: same frame header; coefficients arrive spread over several scans
            // fall through
            case 0xc0: case 0xc1: {   // SOF0 / SOF1
                if (have_frame) return fail("JPEG: second frame header");     // a later SOF would re-size the planes under scans already decoded
                if (se - s < 6) return fail("JPEG: truncated SOF");
                if (s[0] != 8) return fail("JPEG: only 8-bit samples are supported");
                H = (s[1] << 8) | s[2]; W = (s[3] << 8) | s[4];
                int nc = s[5];
                if (W < 1 || H < 1) return fail("JPEG: empty image");
                if (nc == 4) return fail("JPEG: CMYK / 4-component files are not supported");
                if (nc != 1 && nc != 3) return fail("JPEG: unsupported component count");
                if (se - s < 6 + 3 * nc) return fail("JPEG: truncated SOF");
                comps.assign(nc, Component());
                for (int i = 0; i < nc; i++) {
                    Component& c = comps[i];
                    c.id = s[6 + 3 * i]; c.h = s[7 + 3 * i] >> 4; c.v = s[7 + 3 * i] & 15; c.tq = s[8 + 3 * i];
                    if (c.h < 1 || c.h > 4 || c.v < 1 || c.v > 4 || c.tq > 3) return fail("JPEG: bad component spec");
                    hmax = c.h > hmax ? c.h : hmax; vmax = c.v > vmax ? c.v : vmax;
                }
                have_frame = true;
            } break;
            case 0xc3: case 0xc5: case 0xc6: case 0xc7: case 0xc9: case 0xca: case 0xcb: case 0xcd: case 0xce: case 0xcf:
                return fail("JPEG: lossless / hierarchical / arithmetic-coded files are not supported");
            case 0xdd: if (se - s < 2) return fail("JPEG: truncated DRI"); restart = (s[0] << 8) | s[1]; break;
            case 0xee: if (se - s >= 12 && !memcmp(s, "Adobe", 5)) { adobe = true; adobe_transform = s[11]; } break;
            case 0xda: {   // SOS
                if (!have_frame) return fail("JPEG: SOS before SOF");
                if (se - s < 1) return fail("JPEG: truncated SOS");
                int ns = s[0];
                if (ns < 1 || ns > (int)comps.size() || se - s < 1 + 2 * ns + 3) return fail("JPEG: bad SOS");
                if (!progressive && ns != (int)comps.size()) return fail("JPEG: non-interleaved sequential files are not supported");
                Component* sc[4] = {nullptr, nullptr, nullptr, nullptr};
                for (int i = 0; i < ns; i++) {
                    int cid = s[1 + 2 * i], tt = s[2 + 2 * i];
                    Component* c = nullptr;
                    for (auto& cc : comps) if (cc.id == cid) c = &cc;
                    if (!c) return fail("JPEG: SOS names an unknown component");
                    c->td = tt >> 4; c->ta = tt & 15;
                    if (c->td > 3 || c->ta > 3 || !qt_ok[c->tq]) return fail("JPEG: scan refers to a missing table");
                    sc[i] = c;
                }
                if (!geometry) {
                    if (comps.size() == 1) { comps[0].h = comps[0].v = 1; hmax = vmax = 1; }   // a single-component image is never subsampled
                    mcux = (W + 8 * hmax - 1) / (8 * hmax); mcuy = (H + 8 * vmax - 1) / (8 * vmax);
                    if ((long long)mcux * mcuy * 64 * hmax * vmax > (1ll << 31)) return fail("JPEG: image too large");
                    for (auto& c : comps) {
                        c.wblocks = mcux * c.h; c.hblocks = mcuy * c.v;
                        c.dw = (W * c.h + hmax - 1) / hmax; c.dh = (H * c.v + vmax - 1) / vmax;
                        c.plane.assign((size_t)c.wblocks * 8 * c.hblocks * 8, 0);
                        if (progressive) c.coefs.assign((size_t)c.wblocks * c.hblocks * 64, 0);
                    }
                    geometry = true;
                }
                if (!progressive) {
                    for (int i = 0; i < ns; i++) if (!hdc[sc[i]->td].present || !hac[sc[i]->ta].present) return fail("JPEG: scan refers to a missing table");
                    done = true;
                    break;
                }
                // ---- one progressive scan (ITU T.81 Annex G; libjpeg jdphuff.c): spectral band [Ss, Se], successive approximation Ah -> Al
                const int Ss = s[1 + 2 * ns], Se = s[2 + 2 * ns], Ah = s[3 + 2 * ns] >> 4, Al = s[3 + 2 * ns] & 15;
                if (Ss > Se || Se > 63 || Al > 13 || (Ss == 0 && Se != 0) || (Ss > 0 && ns != 1)) return fail("JPEG: bad progressive scan parameters");
                for (int i = 0; i < ns; i++) {
                    if (Ss == 0 && Ah == 0 && !hdc[sc[i]->td].present) return fail("JPEG: scan refers to a missing table");
                    if (Ss > 0 && !hac[sc[i]->ta].present) return fail("JPEG: scan refers to a missing table");
                }
                BitReader br{p, end};
                int eobrun = 0, until = restart;
                for (auto& c : comps) c.pred = 0;
                auto do_restart = [&]() -> bool {
                    const uint8_t* q = br.p;
                    while (q + 1 < end && !(q[0] == 0xff && q[1] >= 0xd0 && q[1] <= 0xd7)) q++;
                    if (q + 1 >= end) return false;
                    br.p = q + 2; br.Reset();
                    for (auto& c : comps) c.pred = 0;
                    eobrun = 0; until = restart;
                    return true;
                };
                auto decode_block = [&](Component& c, int16_t* blk) -> bool {
                    if (Ss == 0) {
                        if (Ah == 0) {                                    // DC first
                            int sz;
                            if (!DecodeSym(br, hdc[c.td], &sz) || sz > 11) return false;
                            c.pred += sz ? Extend(br.Get(sz), sz) : 0;
                            blk[0] = (int16_t)(c.pred * (1 << Al));
                        } else if (br.Get(1)) blk[0] |= (int16_t)(1 << Al);   // DC refinement: one more bit
                        return true;
                    }
                    if (Ah == 0) {                                        // AC first
                        if (eobrun > 0) { eobrun--; return true; }
                        for (int k = Ss; k <= Se; k++) {
                            int rs;
                            if (!DecodeSym(br, hac[c.ta], &rs)) return false;
                            int r = rs >> 4, sz = rs & 15;
                            if (sz) {
                                k += r;
                                if (k > 63) return false;
                                blk[kZigzag[k]] = (int16_t)(Extend(br.Get(sz), sz) * (1 << Al));
                            } else if (r == 15) k += 15;
                            else { eobrun = (1 << r) - 1; if (r) eobrun += br.Get(r); break; }
                        }
                        return true;
                    }
                    // AC refinement
                    const int p1 = 1 << Al, m1 = -(1 << Al);
                    int k = Ss;
                    if (eobrun == 0) {
                        for (; k <= Se; k++) {
                            int rs;
                            if (!DecodeSym(br, hac[c.ta], &rs)) return false;
                            int r = rs >> 4, sz = rs & 15, newval = 0;
                            if (sz) { if (sz != 1) return false; newval = br.Get(1) ? p1 : m1; }
                            else if (r != 15) { eobrun = 1 << r; if (r) eobrun += br.Get(r); break; }
                            do {                                          // skip r zero-history coefficients, refining the non-zero ones passed
                                int16_t* co = blk + kZigzag[k];
                                if (*co != 0) {
                                    if (br.Get(1) && (*co & p1) == 0) *co = (int16_t)(*co + (*co >= 0 ? p1 : m1));
                                } else if (--r < 0) break;
                                k++;
                            } while (k <= Se);
                            if (sz) { if (k > 63) return false; blk[kZigzag[k]] = (int16_t)newval; }
                        }
                    }
                    if (eobrun > 0) {
                        for (; k <= Se; k++) {
                            int16_t* co = blk + kZigzag[k];
                            if (*co != 0 && br.Get(1) && (*co & p1) == 0) *co = (int16_t)(*co + (*co >= 0 ? p1 : m1));
                        }
                        eobrun--;
                    }
                    return true;
                };
                if (ns == 1) {                                            // non-interleaved: the component's own block raster (not MCU-padded)
                    Component& c = *sc[0];
                    int bw = (c.dw + 7) / 8, bh = (c.dh + 7) / 8;
                    for (int by = 0; by < bh; by++)
                        for (int bx = 0; bx < bw; bx++) {
                            if (restart && until == 0 && !do_restart()) return fail("JPEG: missing restart marker");
                            if (!decode_block(c, c.coefs.data() + ((size_t)by * c.wblocks + bx) * 64)) return fail("JPEG: corrupt progressive scan");
                            if (restart) until--;
                        }
                } else {
                    for (int my = 0; my < mcuy; my++)
                        for (int mx = 0; mx < mcux; mx++) {
                            if (restart && until == 0 && !do_restart()) return fail("JPEG: missing restart marker");
                            for (int i = 0; i < ns; i++) {
                                Component& c = *sc[i];
                                for (int by = 0; by < c.v; by++)
                                    for (int bx = 0; bx < c.h; bx++)
                                        if (!decode_block(c, c.coefs.data() + ((size_t)(my * c.v + by) * c.wblocks + (mx * c.h + bx)) * 64)) return fail("JPEG: corrupt progressive scan");
                            }
                            if (restart) until--;
                        }
                }
                p = br.p;
                scans++;
            } break;
            default: break;   // APPn, COM, ...
        }
    }
    if (!geometry) return fail("JPEG: no scan");
    const int mcuw = 8 * hmax, mcuh = 8 * vmax;
    (void)mcuw; (void)mcuh;
    int coef[64];
    if (progressive) {
        // every scan has been folded into the coefficient arrays: dequantise + inverse DCT
        for (auto& c : comps)
            for (int by = 0; by < c.hblocks; by++)
                for (int bx = 0; bx < c.wblocks; bx++) {
                    const int16_t* blk = c.coefs.data() + ((size_t)by * c.wblocks + bx) * 64;
                    for (int k = 0; k < 64; k++) coef[k] = blk[k] * qt[c.tq][k];
                    IdctIslow(coef, c.plane.data() + (size_t)by * 8 * c.wblocks * 8 + bx * 8, c.wblocks * 8);
                }
    }
    // ---- entropy-coded data of a sequential file: one interleaved scan
    BitReader br{p, end};
    int until_restart = restart;
    for (int my = 0; !progressive && my < mcuy; my++) {
        for (int mx = 0; mx < mcux; mx++) {
            if (restart && until_restart == 0) {
                // RSTn: byte-align, skip the marker, reset predictors
                const uint8_t* q = br.p;
                while (q + 1 < end && !(q[0] == 0xff && q[1] >= 0xd0 && q[1] <= 0xd7)) q++;
                if (q + 1 >= end) return fail("JPEG: missing restart marker");
                br.p = q + 2; br.Reset();
                for (auto& c : comps) c.pred = 0;
                until_restart = restart;
            }
            for (auto& c : comps) {
                for (int by = 0; by < c.v; by++) {
                    for (int bx = 0; bx < c.h; bx++) {
                        memset(coef, 0, sizeof coef);
                        int s;
                        if (!DecodeSym(br, hdc[c.td], &s)) return fail("JPEG: corrupt Huffman data");
                        if (s > 11) return fail("JPEG: corrupt DC coefficient");
                        int diff = s ? Extend(br.Get(s), s) : 0;
                        c.pred += diff;
                        coef[0] = c.pred * qt[c.tq][0];
                        for (int k = 1; k < 64;) {
                            int rs;
                            if (!DecodeSym(br, hac[c.ta], &rs)) return fail("JPEG: corrupt Huffman data");
                            int r = rs >> 4, sz = rs & 15;
                            if (sz == 0) { if (r == 15) { k += 16; continue; } break; }
                            k += r;
                            if (k > 63) return fail("JPEG: corrupt AC coefficients");
                            int nat = kZigzag[k];
                            coef[nat] = Extend(br.Get(sz), sz) * qt[c.tq][nat];
                            k++;
                        }
                        int px = (mx * c.h + bx) * 8, py = (my * c.v + by) * 8;
                        IdctIslow(coef, c.plane.data() + (size_t)py * c.wblocks * 8 + px, c.wblocks * 8);
                    }
                }
            }
            if (restart) until_restart--;
        }
    }
    // ---- upsample + colour convert
    const int nc = (int)comps.size();
    *width = W; *height = H; *ncomp = nc;
    pixels->assign((size_t)W * H * nc, 0);
    std::vector<std::vector<uint8_t>> full(nc);
    for (int ci = 0; ci < nc; ci++) {
        Component& c = comps[ci];
        const int stride = c.wblocks * 8;
        auto row = [&](int y) { y = y < 0 ? 0 : (y >= c.dh ? c.dh - 1 : y); return c.plane.data() + (size_t)y * stride; };
        std::vector<uint8_t>& o = full[ci];
        const int ow = c.dw * (hmax / c.h > 0 ? hmax / c.h : 1) + 2;
        o.assign((size_t)(ow) * H, 0);
        bool h2 = c.h * 2 == hmax, h1 = c.h == hmax, v2 = c.v * 2 == vmax, v1 = c.v == vmax;
        if (h1 && v1) {
            for (int y = 0; y < H; y++) memcpy(&o[(size_t)y * ow], row(y), (size_t)(W < stride ? W : stride));
        } else if (h2 && v1 && c.dw > 2) {                              // jdsample.c h2v1_fancy_upsample
            for (int y = 0; y < H; y++) {
                const uint8_t* in = row(y); uint8_t* out = &o[(size_t)y * ow];
                int n = c.dw, inv = in[0];
                *out++ = (uint8_t)inv; *out++ = (uint8_t)((inv * 3 + in[1] + 2) >> 2);
                for (int i = 1; i < n - 1; i++) { inv = in[i] * 3; *out++ = (uint8_t)((inv + in[i - 1] + 1) >> 2); *out++ = (uint8_t)((inv + in[i + 1] + 2) >> 2); }
                inv = in[n - 1];
                *out++ = (uint8_t)((inv * 3 + in[n - 2] + 1) >> 2); *out++ = (uint8_t)inv;
            }
        } else if (h2 && v2 && c.dw > 2) {                              // jdsample.c h2v2_fancy_upsample
            for (int y = 0; y < H; y++) {
                int sy = y >> 1;
                const uint8_t* in0 = row(sy);
                const uint8_t* in1 = row((y & 1) ? sy + 1 : sy - 1);
                uint8_t* out = &o[(size_t)y * ow];
                int n = c.dw;
                int thiscol = in0[0] * 3 + in1[0], nextcol = in0[1] * 3 + in1[1], lastcol;
                *out++ = (uint8_t)((thiscol * 4 + 8) >> 4); *out++ = (uint8_t)((thiscol * 3 + nextcol + 7) >> 4);
                lastcol = thiscol; thiscol = nextcol;
                for (int i = 2; i < n; i++) {
                    nextcol = in0[i] * 3 + in1[i];
                    *out++ = (uint8_t)((thiscol * 3 + lastcol + 8) >> 4); *out++ = (uint8_t)((thiscol * 3 + nextcol + 7) >> 4);
                    lastcol = thiscol; thiscol = nextcol;
                }
                *out++ = (uint8_t)((thiscol * 3 + lastcol + 8) >> 4); *out++ = (uint8_t)((thiscol * 4 + 7) >> 4);
            }
        } else {                                                         // any other ratio: pixel replication (int_upsample)
            if (hmax % c.h || vmax % c.v) return fail("JPEG: fractional sampling ratios are not supported");
            int hx = hmax / c.h, vx = vmax / c.v;
            for (int y = 0; y < H; y++) {
                const uint8_t* in = row(y / vx); uint8_t* out = &o[(size_t)y * ow];
                for (int x = 0; x < W; x++) out[x] = in[x / hx];
            }
        }
    }
    const int ow0 = comps[0].dw * (hmax / comps[0].h) + 2;
    if (nc == 1) {
        for (int y = 0; y < H; y++) memcpy(&(*pixels)[(size_t)y * W], &full[0][(size_t)y * ow0], W);
        return true;
    }
    const bool ycc = !adobe || adobe_transform != 0;                     // jdapimin.c default_decompress_parms: 3 components = YCbCr unless Adobe says RGB
    // jdcolor.c build_ycc_rgb_table
    int cr_r[256], cb_b[256], cr_g[256], cb_g[256];
    for (int i = 0; i < 256; i++) {
        int x = i - 128;
        cr_r[i] = (int)((91881LL * x + 32768) >> 16);                    // FIX(1.40200)
        cb_b[i] = (int)((116130LL * x + 32768) >> 16);                   // FIX(1.77200)
        cr_g[i] = -46802 * x;                                            // FIX(0.71414)
        cb_g[i] = -22554 * x + 32768;                                    // FIX(0.34414) + ONE_HALF
    }
    const int ow1 = comps[1].dw * (hmax / comps[1].h) + 2, ow2 = comps[2].dw * (hmax / comps[2].h) + 2;
    for (int y = 0; y < H; y++) {
        const uint8_t *a = &full[0][(size_t)y * ow0], *b = &full[1][(size_t)y * ow1], *c = &full[2][(size_t)y * ow2];
        uint8_t* o = &(*pixels)[(size_t)y * W * 3];
        for (int x = 0; x < W; x++) {
            if (ycc) {
                int Y = a[x], cb = b[x], cr = c[x];
                o[3 * x + 0] = Clamp8(Y + cr_r[cr]);
                o[3 * x + 1] = Clamp8(Y + ((cb_g[cb] + cr_g[cr]) >> 16));
                o[3 * x + 2] = Clamp8(Y + cb_b[cb]);
            } else { o[3 * x + 0] = a[x]; o[3 * x + 1] = b[x]; o[3 * x + 2] = c[x]; }
        }
    }
    return true;
}

}  // namespace rt795
