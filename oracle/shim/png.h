// Stand-in for the libpng calls made by /root/reference/src/Texture.cpp:239-299 (TEST INFRASTRUCTURE ONLY).
// libpng is not installed in this image.  This shim decodes non-interlaced 8-bit PNGs (colour types
// 0, 2, 4, 6) with zlib and hands the reference RGBA8 rows, which is what its ReadPNG asks libpng for.
#ifndef RT795_ORACLE_SHIM_PNG_H
#define RT795_ORACLE_SHIM_PNG_H
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <csetjmp>
#include <vector>
#include <zlib.h>

typedef unsigned char png_byte;
typedef png_byte* png_bytep;
typedef unsigned int png_uint_32;

#define PNG_LIBPNG_VER_STRING "shim"
#define PNG_COLOR_TYPE_GRAY 0
#define PNG_COLOR_TYPE_RGB 2
#define PNG_COLOR_TYPE_PALETTE 3
#define PNG_COLOR_TYPE_GRAY_ALPHA 4
#define PNG_COLOR_TYPE_RGB_ALPHA 6
#define PNG_INFO_tRNS 0x0010
#define PNG_FILLER_AFTER 1

struct rt795_png_struct {
    FILE* fp = nullptr;
    jmp_buf jb;
    int width = 0, height = 0, color_type = 0, bit_depth = 0;
    std::vector<unsigned char> rgba;  // decoded RGBA8
};
typedef rt795_png_struct* png_structp;
typedef rt795_png_struct* png_infop;

#define png_jmpbuf(png) ((png)->jb)

static inline png_structp png_create_read_struct(const char*, void*, void*, void*) { return new rt795_png_struct(); }
static inline png_infop png_create_info_struct(png_structp p) { return p; }
static inline void png_init_io(png_structp p, FILE* fp) { p->fp = fp; }

static inline unsigned rt795_be32(const unsigned char* b) { return (b[0] << 24) | (b[1] << 16) | (b[2] << 8) | b[3]; }

static inline void png_read_info(png_structp p, png_infop) {
    if (!p->fp) { fprintf(stderr, "png shim: file not open\n"); abort(); }
    std::vector<unsigned char> file;
    unsigned char buf[65536];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, p->fp)) > 0) file.insert(file.end(), buf, buf + n);
    static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (file.size() < 8 || memcmp(file.data(), sig, 8) != 0) { fprintf(stderr, "png shim: bad signature\n"); abort(); }
    std::vector<unsigned char> idat;
    int interlace = 0;
    size_t pos = 8;
    while (pos + 12 <= file.size()) {
        unsigned len = rt795_be32(&file[pos]);
        const unsigned char* type = &file[pos + 4];
        const unsigned char* data = &file[pos + 8];
        if (!memcmp(type, "IHDR", 4)) {
            p->width = rt795_be32(data); p->height = rt795_be32(data + 4);
            p->bit_depth = data[8]; p->color_type = data[9]; interlace = data[12];
        } else if (!memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), data, data + len);
        } else if (!memcmp(type, "IEND", 4)) {
            break;
        }
        pos += 12 + len;
    }
    if (p->bit_depth != 8 || interlace != 0 || p->color_type == PNG_COLOR_TYPE_PALETTE) {
        fprintf(stderr, "png shim: only non-interlaced 8-bit gray/rgb(/alpha) supported\n"); abort();
    }
    int ch = p->color_type == 0 ? 1 : p->color_type == 2 ? 3 : p->color_type == 4 ? 2 : 4;
    size_t stride = (size_t)p->width * ch;
    std::vector<unsigned char> raw((stride + 1) * p->height);
    uLongf rawLen = raw.size();
    if (uncompress(raw.data(), &rawLen, idat.data(), idat.size()) != Z_OK) { fprintf(stderr, "png shim: inflate failed\n"); abort(); }
    std::vector<unsigned char> img(stride * p->height);
    for (int y = 0; y < p->height; y++) {
        int ft = raw[y * (stride + 1)];
        const unsigned char* in = &raw[y * (stride + 1) + 1];
        unsigned char* out = &img[y * stride];
        const unsigned char* up = y ? &img[(y - 1) * stride] : nullptr;
        for (size_t i = 0; i < stride; i++) {
            int a = i >= (size_t)ch ? out[i - ch] : 0, b = up ? up[i] : 0, c = (up && i >= (size_t)ch) ? up[i - ch] : 0;
            int pred = 0;
            switch (ft) {
                case 0: pred = 0; break;
                case 1: pred = a; break;
                case 2: pred = b; break;
                case 3: pred = (a + b) >> 1; break;
                case 4: { int pp = a + b - c, pa = abs(pp - a), pb = abs(pp - b), pc = abs(pp - c);
                          pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c); } break;
            }
            out[i] = (unsigned char)(in[i] + pred);
        }
    }
    p->rgba.resize((size_t)p->width * p->height * 4);
    for (size_t i = 0; i < (size_t)p->width * p->height; i++) {
        const unsigned char* s = &img[i * ch];
        unsigned char* d = &p->rgba[i * 4];
        if (ch == 1) { d[0] = d[1] = d[2] = s[0]; d[3] = 255; }
        else if (ch == 2) { d[0] = d[1] = d[2] = s[0]; d[3] = s[1]; }
        else if (ch == 3) { d[0] = s[0]; d[1] = s[1]; d[2] = s[2]; d[3] = 255; }
        else { d[0] = s[0]; d[1] = s[1]; d[2] = s[2]; d[3] = s[3]; }
    }
}
static inline png_uint_32 png_get_image_width(png_structp p, png_infop) { return p->width; }
static inline png_uint_32 png_get_image_height(png_structp p, png_infop) { return p->height; }
static inline png_byte png_get_color_type(png_structp p, png_infop) { return (png_byte)p->color_type; }
static inline png_byte png_get_bit_depth(png_structp p, png_infop) { return (png_byte)p->bit_depth; }
static inline void png_set_strip_16(png_structp) {}
static inline void png_set_palette_to_rgb(png_structp) {}
static inline void png_set_expand_gray_1_2_4_to_8(png_structp) {}
static inline png_uint_32 png_get_valid(png_structp, png_infop, png_uint_32) { return 0; }
static inline void png_set_tRNS_to_alpha(png_structp) {}
static inline void png_set_filler(png_structp, int, int) {}
static inline void png_set_gray_to_rgb(png_structp) {}
static inline void png_read_update_info(png_structp, png_infop) {}
static inline size_t png_get_rowbytes(png_structp p, png_infop) { return (size_t)p->width * 4; }
static inline void png_read_image(png_structp p, png_bytep* rows) {
    for (int y = 0; y < p->height; y++) memcpy(rows[y], &p->rgba[(size_t)y * p->width * 4], (size_t)p->width * 4);
}
static inline void png_destroy_read_struct(png_structp* p, png_infop*, void*) { delete *p; *p = nullptr; }
#endif
