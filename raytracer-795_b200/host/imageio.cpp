#include "imageio.h"
#include "jpeg.h"
#include <zlib.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>

namespace rt795 {

// src/Texture.cpp:133-183 / src/Image.cpp:36-60: state machine '.' 'p' 'n' 'g' over the whole name.
bool NameHasExt(const char* name, const char* e) {
    int c = 0;
    for (int i = 0; name[i] != '\0'; i++) {
        char ch = name[i];
        if (ch == '.') c = 1;
        else if (ch == e[0] && c == 1) c = 2;
        else if (ch == e[1] && c == 2) c = 3;
        else if (ch == e[2] && c == 3) return true;
        else c = 0;
    }
    return false;
}

static bool ReadAll(const std::string& path, std::vector<unsigned char>* out, std::string* err) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) { *err = "cannot open " + path; return false; }
    unsigned char buf[1 << 16];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) out->insert(out->end(), buf, buf + n);
    fclose(f);
    return true;
}

uint16_t FloatToHalf(float f) {
    uint32_t x; memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u;
    int32_t exp = (int32_t)((x >> 23) & 0xff) - 127 + 15;
    uint32_t man = x & 0x7fffffu;
    if (((x >> 23) & 0xff) == 0xff) return (uint16_t)(sign | 0x7c00u | (man ? 0x200u : 0));   // inf / nan
    if (exp >= 31) return (uint16_t)(sign | 0x7c00u);                                          // overflow -> inf (F8)
    if (exp <= 0) {
        if (exp < -10) return (uint16_t)sign;
        man |= 0x800000u;
        uint32_t shift = (uint32_t)(14 - exp);
        uint32_t h = man >> shift;
        uint32_t rem = man & ((1u << shift) - 1), half = 1u << (shift - 1);
        if (rem > half || (rem == half && (h & 1))) h++;
        return (uint16_t)(sign | h);
    }
    uint32_t h = ((uint32_t)exp << 10) | (man >> 13);
    uint32_t rem = man & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (h & 1))) h++;
    return (uint16_t)(sign | h);
}

float HalfToFloat(uint16_t h) {
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1f, man = h & 0x3ffu, x;
    if (exp == 0) {
        if (man == 0) x = sign;
        else { int e = -1; do { e++; man <<= 1; } while (!(man & 0x400u)); x = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3ffu) << 13); }
    } else if (exp == 31) x = sign | 0x7f800000u | (man << 13);
    else x = sign | ((exp - 15 + 127) << 23) | (man << 13);
    float f; memcpy(&f, &x, 4); return f;
}

// ---------------------------------------------------------------------------------------------- PNG
static uint32_t BE32(const unsigned char* b) { return ((uint32_t)b[0] << 24) | (b[1] << 16) | (b[2] << 8) | b[3]; }

bool LoadPng(const std::string& path, LoadedImage* out, std::string* err) {
    std::vector<unsigned char> file;
    if (!ReadAll(path, &file, err)) return false;
    static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (file.size() < 8 || memcmp(file.data(), sig, 8) != 0) { *err = "not a PNG: " + path; return false; }
    int w = 0, h = 0, depth = 0, ctype = 0, interlace = 0;
    std::vector<unsigned char> idat, plte, trns;
    size_t pos = 8;
    while (pos + 12 <= file.size()) {
        uint32_t len = BE32(&file[pos]);
        if (pos + 12 + len > file.size()) { *err = "truncated PNG: " + path; return false; }
        const unsigned char* type = &file[pos + 4];
        const unsigned char* d = &file[pos + 8];
        if (!memcmp(type, "IHDR", 4)) {
            if (len != 13) { *err = "PNG: bad IHDR length: " + path; return false; }
            w = BE32(d); h = BE32(d + 4); depth = d[8]; ctype = d[9]; interlace = d[12];
        }
        else if (!memcmp(type, "PLTE", 4)) plte.assign(d, d + len);
        else if (!memcmp(type, "tRNS", 4)) trns.assign(d, d + len);
        else if (!memcmp(type, "IDAT", 4)) idat.insert(idat.end(), d, d + len);
        else if (!memcmp(type, "IEND", 4)) break;
        pos += 12 + len;
    }
    if (interlace) { *err = "interlaced PNG not supported: " + path; return false; }
    if (depth != 8 && depth != 16) { *err = "PNG bit depth must be 8 or 16: " + path; return false; }
    int ch = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : 4;
    int bps = depth / 8, bpp = ch * bps;
    size_t stride = (size_t)w * bpp;
    std::vector<unsigned char> raw((stride + 1) * h);
    uLongf rawLen = raw.size();
    if (uncompress(raw.data(), &rawLen, idat.data(), idat.size()) != Z_OK || rawLen != raw.size()) { *err = "PNG inflate failed: " + path; return false; }
    std::vector<unsigned char> img(stride * h);
    for (int y = 0; y < h; y++) {
        int ft = raw[y * (stride + 1)];
        const unsigned char* in = &raw[y * (stride + 1) + 1];
        unsigned char* o = &img[y * stride];
        const unsigned char* up = y ? &img[(size_t)(y - 1) * stride] : nullptr;
        for (size_t i = 0; i < stride; i++) {
            int a = i >= (size_t)bpp ? o[i - bpp] : 0, b = up ? up[i] : 0, c = (up && i >= (size_t)bpp) ? up[i - bpp] : 0, pred = 0;
            switch (ft) {
                case 1: pred = a; break;
                case 2: pred = b; break;
                case 3: pred = (a + b) >> 1; break;
                case 4: { int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c); pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c); } break;
                default: break;
            }
            o[i] = (unsigned char)(in[i] + pred);
        }
    }
    out->width = w; out->height = h; out->format = 0;
    out->bytes.resize((size_t)w * h * 4);
    for (size_t i = 0; i < (size_t)w * h; i++) {
        const unsigned char* s = &img[i * bpp];
        unsigned char* d = &out->bytes[i * 4];
        auto S = [&](int k) { return s[k * bps]; };   // 16-bit: strip to the high byte (png_set_strip_16)
        if (ctype == 0) { d[0] = d[1] = d[2] = S(0); d[3] = 255; }
        else if (ctype == 2) { d[0] = S(0); d[1] = S(1); d[2] = S(2); d[3] = 255; }
        else if (ctype == 3) { size_t k = s[0]; if (k * 3 + 2 >= plte.size() + 0 && plte.size() < k * 3 + 3) { *err = "PNG palette index out of range"; return false; }
                               d[0] = plte[k * 3]; d[1] = plte[k * 3 + 1]; d[2] = plte[k * 3 + 2]; d[3] = k < trns.size() ? trns[k] : 255; }
        else if (ctype == 4) { d[0] = d[1] = d[2] = S(0); d[3] = S(1); }
        else { d[0] = S(0); d[1] = S(1); d[2] = S(2); d[3] = S(3); }
    }
    return true;
}

// ---------------------------------------------------------------------------------------------- EXR
bool LoadExr(const std::string& path, LoadedImage* out, std::string* err) {
    std::vector<unsigned char> f;
    if (!ReadAll(path, &f, err)) return false;
    if (f.size() < 8 || f[0] != 0x76 || f[1] != 0x2f || f[2] != 0x31 || f[3] != 0x01) { *err = "not an EXR: " + path; return false; }
    uint32_t version; memcpy(&version, &f[4], 4);
    if (version & 0x200u) { *err = "tiled EXR not supported: " + path; return false; }
    if (version & 0x1800u) { *err = "deep/multipart EXR not supported: " + path; return false; }
    size_t pos = 8;
    struct Chan { std::string name; int type; };
    std::vector<Chan> chans;
    int compression = 0, xmin = 0, ymin = 0, xmax = -1, ymax = -1;
    auto cstr = [&](std::string* s) { s->clear(); while (pos < f.size() && f[pos]) s->push_back((char)f[pos++]); pos++; return pos <= f.size(); };
    for (;;) {
        std::string name, type;
        if (pos >= f.size()) { *err = "truncated EXR header"; return false; }
        if (f[pos] == 0) { pos++; break; }
        if (!cstr(&name) || !cstr(&type) || pos + 4 > f.size()) { *err = "truncated EXR header"; return false; }
        int32_t sz; memcpy(&sz, &f[pos], 4); pos += 4;
        if (sz < 0 || pos + (size_t)sz > f.size()) { *err = "truncated EXR header"; return false; }
        const size_t aend = pos + (size_t)sz;
        if (name == "channels") {
            size_t p = pos;
            while (p < aend && f[p]) {
                Chan c;
                while (p < aend && f[p]) c.name.push_back((char)f[p++]);
                p++;
                if (p + 16 > aend) { *err = "bad EXR channel list"; return false; }
                int32_t t; memcpy(&t, &f[p], 4); c.type = t; p += 16;
                if (t < 0 || t > 2 || chans.size() >= 64) { *err = "bad EXR channel list"; return false; }
                chans.push_back(c);
            }
        } else if (name == "compression") { if (sz < 1) { *err = "bad EXR header"; return false; } compression = f[pos]; }
        else if (name == "dataWindow") {
            if (sz < 16) { *err = "bad EXR header"; return false; }
            int32_t b[4]; memcpy(b, &f[pos], 16); xmin = b[0]; ymin = b[1]; xmax = b[2]; ymax = b[3];
        }
        pos += sz;
    }
    if (compression != 0 && compression != 2 && compression != 3) { *err = "EXR compression not supported (NONE/ZIPS/ZIP only): " + path; return false; }
    long long wl = (long long)xmax - xmin + 1, hl = (long long)ymax - ymin + 1;
    if (wl <= 0 || hl <= 0 || wl > 65536 || hl > 65536 || chans.empty()) { *err = "bad EXR header: " + path; return false; }
    int w = (int)wl, h = (int)hl;
    int lines_per_block = compression == 3 ? 16 : 1;
    int nblocks = (h + lines_per_block - 1) / lines_per_block;
    if (pos + 8ull * nblocks > f.size()) { *err = "truncated EXR offsets"; return false; }
    std::vector<uint64_t> offs(nblocks);
    memcpy(offs.data(), &f[pos], 8ull * nblocks);
    size_t line_bytes = 0;
    std::vector<size_t> chan_off(chans.size());
    for (size_t c = 0; c < chans.size(); c++) { chan_off[c] = line_bytes; line_bytes += (size_t)w * (chans[c].type == 1 ? 2 : 4); }
    int cidx[4] = {-1, -1, -1, -1};
    for (size_t c = 0; c < chans.size(); c++) {
        if (chans[c].name == "R") cidx[0] = c; else if (chans[c].name == "G") cidx[1] = c;
        else if (chans[c].name == "B") cidx[2] = c; else if (chans[c].name == "A") cidx[3] = c;
    }
    if (cidx[0] < 0 && chans.size() == 1) cidx[0] = cidx[1] = cidx[2] = 0;   // single channel -> grey
    out->width = w; out->height = h; out->format = 1;
    out->bytes.assign((size_t)w * h * 16, 0);
    float* px = (float*)out->bytes.data();
    for (size_t i = 0; i < (size_t)w * h; i++) px[i * 4 + 3] = 1.0f;
    std::vector<unsigned char> block, tmp;
    for (int b = 0; b < nblocks; b++) {
        size_t p = offs[b];
        if (p + 8 > f.size()) { *err = "truncated EXR block"; return false; }
        int32_t y0, dsz; memcpy(&y0, &f[p], 4); memcpy(&dsz, &f[p + 4], 4); p += 8;
        if (dsz < 0 || p + (size_t)dsz > f.size()) { *err = "truncated EXR block"; return false; }
        if (y0 < ymin || y0 > ymax) { *err = "bad EXR block header"; return false; }
        int nl = std::min(lines_per_block, ymax - y0 + 1);
        size_t want = line_bytes * nl;
        block.resize(want);
        if (compression == 0 || (size_t)dsz == want) {
            if ((size_t)dsz < want) { *err = "truncated EXR block"; return false; }
            memcpy(block.data(), &f[p], want);
        }
        else {
            tmp.resize(want);
            uLongf got = want;
            if (uncompress(tmp.data(), &got, &f[p], dsz) != Z_OK || got != want) { *err = "EXR inflate failed"; return false; }
            for (size_t i = 1; i < want; i++) tmp[i] = (unsigned char)(tmp[i - 1] + tmp[i] - 128);   // predictor
            size_t half = (want + 1) / 2;                                                             // de-interleave
            for (size_t i = 0; i < want; i++) block[i] = (i & 1) ? tmp[half + i / 2] : tmp[i / 2];
        }
        for (int l = 0; l < nl; l++) {
            int y = y0 - ymin + l;
            const unsigned char* line = &block[line_bytes * l];
            for (int k = 0; k < 4; k++) {
                if (cidx[k] < 0) continue;
                const unsigned char* src = line + chan_off[cidx[k]];
                int t = chans[cidx[k]].type;
                for (int x = 0; x < w; x++) {
                    float v;
                    if (t == 1) { uint16_t hv; memcpy(&hv, src + 2 * x, 2); v = HalfToFloat(hv); }
                    else if (t == 2) memcpy(&v, src + 4 * x, 4);
                    else { uint32_t u; memcpy(&u, src + 4 * x, 4); v = (float)u; }
                    px[((size_t)y * w + x) * 4 + k] = v;
                }
            }
        }
    }
    return true;
}

bool LoadImageFile(const std::string& path, LoadedImage* out, std::string* err) {
    // dispatch exactly as the Texture ctor does (src/Texture.cpp:8-22): ".png" anywhere, else ".exr", else JPEG
    if (NameHasExt(path.c_str(), "png")) return LoadPng(path, out, err);
    if (NameHasExt(path.c_str(), "exr")) return LoadExr(path, out, err);
    return LoadJpeg(path, out, err);
}

// Texture::ReadJPG (src/Texture.cpp:191-240) keeps libjpeg's interleaved scanlines and numComponents; GetColorAtPixel
// (:62-66) then reads raw[(j*w+i)*nc + 0..2].  For 3 components that is the pixel's RGB.  For a greyscale file (nc = 1) it
// is the pixel and its two right-hand neighbours in scanline order (wrapping into the next row, running past the buffer at
// the very end) - reproduced here, with the last two reads clamped to the final sample.
bool LoadJpeg(const std::string& path, LoadedImage* out, std::string* err) {
    std::vector<unsigned char> file;
    if (!ReadAll(path, &file, err)) return false;
    int w = 0, h = 0, nc = 0;
    std::vector<uint8_t> px;
    std::string why;
    if (!DecodeJpeg(file, &w, &h, &nc, &px, &why)) { *err = why + ": " + path; return false; }
    out->width = w; out->height = h; out->format = 0 /* RT_IMG_RGBA8 */;
    out->bytes.resize((size_t)w * h * 4);
    const size_t n = (size_t)w * h * nc;
    for (size_t i = 0; i < (size_t)w * h; i++) {
        for (int c = 0; c < 3; c++) { size_t k = i * nc + c; out->bytes[i * 4 + c] = px[k < n ? k : n - 1]; }
        out->bytes[i * 4 + 3] = 255;
    }
    return true;
}

// ---------------------------------------------------------------------------------------------- writers
bool SavePpmP3(const std::string& path, const float* rgb, int nx, int ny, std::string* err) {
    FILE* o = fopen(path.c_str(), "w");
    if (!o) { *err = "cannot write " + path; return false; }
    fprintf(o, "P3\n%d %d\n255\n", nx, ny);
    // "<value> " per component through a 256-entry table (6.2 M values for a 1080p frame: printf per value was 40 ms of the CLI)
    char tab[256][4]; unsigned char len[256];
    for (int v = 0; v < 256; v++) len[v] = (unsigned char)snprintf(tab[v], sizeof tab[v], "%d", v);
    std::vector<char> line((size_t)nx * 3 * 4 + 1);
    bool ok = true;
    for (int y = 0; y < ny && ok; y++) {
        char* w = line.data();
        for (int i = 0; i < nx * 3; i++) {
            float v = rgb[(size_t)y * nx * 3 + i];
            if (v > 255) v = 255;                      // src/Image.cpp:64-68
            if (!(v >= 0)) v = 0;                      // the reference's (unsigned char) cast of a negative/NaN is UB; clamp
            int b = (int)(unsigned char)v;
            memcpy(w, tab[b], len[b]); w += len[b];
            *w++ = ' ';
        }
        *w++ = '\n';
        ok = fwrite(line.data(), 1, (size_t)(w - line.data()), o) == (size_t)(w - line.data());
    }
    if (fclose(o) != 0) ok = false;
    if (!ok) { *err = "write failed: " + path; return false; }
    return true;
}

static void PutAttr(std::vector<unsigned char>& h, const char* name, const char* type, const void* data, int32_t n) {
    h.insert(h.end(), name, name + strlen(name) + 1);
    h.insert(h.end(), type, type + strlen(type) + 1);
    const unsigned char* p = (const unsigned char*)&n; h.insert(h.end(), p, p + 4);
    p = (const unsigned char*)data; h.insert(h.end(), p, p + n);
}

bool SaveExrHalfBGR(const std::string& path, const float* rgb, int nx, int ny, std::string* err) {
    std::vector<unsigned char> h;
    uint32_t magic = 20000630, ver = 2;
    h.insert(h.end(), (unsigned char*)&magic, (unsigned char*)&magic + 4);
    h.insert(h.end(), (unsigned char*)&ver, (unsigned char*)&ver + 4);
    std::vector<unsigned char> ch;
    for (const char* n : {"B", "G", "R"}) {
        ch.push_back((unsigned char)n[0]); ch.push_back(0);
        int32_t v[4] = {1 /*HALF*/, 0, 1, 1};
        ch.insert(ch.end(), (unsigned char*)v, (unsigned char*)v + 16);
    }
    ch.push_back(0);
    PutAttr(h, "channels", "chlist", ch.data(), (int32_t)ch.size());
    unsigned char zero = 0;
    PutAttr(h, "compression", "compression", &zero, 1);
    int32_t box[4] = {0, 0, nx - 1, ny - 1};
    PutAttr(h, "dataWindow", "box2i", box, 16);
    PutAttr(h, "displayWindow", "box2i", box, 16);
    PutAttr(h, "lineOrder", "lineOrder", &zero, 1);
    float one = 1.0f, c2[2] = {0, 0};
    PutAttr(h, "pixelAspectRatio", "float", &one, 4);
    PutAttr(h, "screenWindowCenter", "v2f", c2, 8);
    PutAttr(h, "screenWindowWidth", "float", &one, 4);
    h.push_back(0);
    size_t line_bytes = (size_t)nx * 2 * 3;
    uint64_t off0 = h.size() + 8ull * ny;
    FILE* o = fopen(path.c_str(), "wb");
    if (!o) { *err = "cannot write " + path; return false; }
    fwrite(h.data(), 1, h.size(), o);
    for (int y = 0; y < ny; y++) { uint64_t off = off0 + (uint64_t)y * (8 + line_bytes); fwrite(&off, 8, 1, o); }
    std::vector<uint16_t> line((size_t)nx * 3);
    for (int y = 0; y < ny; y++) {
        int32_t hd[2] = {y, (int32_t)line_bytes};
        fwrite(hd, 4, 2, o);
        for (int x = 0; x < nx; x++) {
            const float* p = &rgb[((size_t)y * nx + x) * 3];
            line[x] = FloatToHalf(p[2]); line[nx + x] = FloatToHalf(p[1]); line[2 * (size_t)nx + x] = FloatToHalf(p[0]);
        }
        fwrite(line.data(), 2, line.size(), o);
    }
    fclose(o);
    return true;
}

bool SaveF32(const std::string& path, const float* rgb, int nx, int ny, std::string* err) {
    FILE* o = fopen(path.c_str(), "wb");
    if (!o) { *err = "cannot write " + path; return false; }
    int32_t hd[2] = {nx, ny};
    fwrite(hd, 4, 2, o);
    fwrite(rgb, sizeof(float), (size_t)nx * ny * 3, o);
    fclose(o);
    return true;
}

}  // namespace rt795
