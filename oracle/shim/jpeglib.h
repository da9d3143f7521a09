// Stand-in for libjpeg (not installed; TEST INFRASTRUCTURE ONLY).  The reference reads JPEG textures with
// jpeg_read_header / jpeg_start_decompress / jpeg_read_scanlines (src/Texture.cpp:191-237).  To pin the product's own JPEG
// decoder (raytracer-795_b200/host/jpeg.h) against an INDEPENDENT libjpeg, this shim hands the file to the Python
// interpreter on PATH, whose Pillow bundles libjpeg-turbo, and serves the decoded scanlines back through the libjpeg entry
// points: the reference binary then sees exactly what a real libjpeg would give it.  Greyscale files stay one component
// ("L"), everything else is three ("RGB"), as jpeg_start_decompress's defaults produce.
#ifndef RT795_ORACLE_SHIM_JPEGLIB_H
#define RT795_ORACLE_SHIM_JPEGLIB_H
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
typedef unsigned char JSAMPLE;
typedef JSAMPLE* JSAMPROW;
typedef JSAMPROW* JSAMPARRAY;
#ifndef TRUE
#define TRUE 1
#endif
struct jpeg_error_mgr { int unused; };
struct jpeg_decompress_struct {
    jpeg_error_mgr* err;
    unsigned image_width, image_height, output_width, output_height, output_scanline;
    int num_components;
    FILE* rt795_src;
    std::vector<unsigned char>* rt795_pixels;
};
static inline void rt795_jpeg_die(const char* why) { fprintf(stderr, "jpeg shim: %s\n", why); abort(); }
static inline jpeg_error_mgr* jpeg_std_error(jpeg_error_mgr* e) { return e; }
static inline void jpeg_create_decompress(jpeg_decompress_struct* c) { c->rt795_src = nullptr; c->rt795_pixels = new std::vector<unsigned char>(); c->output_scanline = 0; }
static inline void jpeg_stdio_src(jpeg_decompress_struct* c, FILE* f) { c->rt795_src = f; }
static inline int jpeg_read_header(jpeg_decompress_struct* c, int) {
    char link[64], path[4096];
    snprintf(link, sizeof link, "/proc/self/fd/%d", fileno(c->rt795_src));
    ssize_t n = readlink(link, path, sizeof path - 1);
    if (n <= 0) rt795_jpeg_die("cannot resolve the texture's path");
    path[n] = 0;
    std::string raw = std::string("/tmp/rt795_jpeg_shim_") + std::to_string((long)getpid()) + ".raw";
    std::string cmd = "python3 -c \"import sys, struct\nfrom PIL import Image\nim = Image.open(sys.argv[1])\nim = im if im.mode == 'L' else im.convert('RGB')\n"
                      "open(sys.argv[2], 'wb').write(struct.pack('<III', im.width, im.height, len(im.getbands())) + im.tobytes())\" '" + std::string(path) + "' '" + raw + "'";
    if (system(cmd.c_str()) != 0) rt795_jpeg_die("python3 + Pillow could not decode the file");
    FILE* f = fopen(raw.c_str(), "rb");
    if (!f) rt795_jpeg_die("decoded file missing");
    unsigned hdr[3];
    if (fread(hdr, 4, 3, f) != 3) rt795_jpeg_die("decoded file truncated");
    c->image_width = hdr[0]; c->image_height = hdr[1]; c->num_components = (int)hdr[2];
    c->rt795_pixels->resize((size_t)hdr[0] * hdr[1] * hdr[2]);
    if (fread(c->rt795_pixels->data(), 1, c->rt795_pixels->size(), f) != c->rt795_pixels->size()) rt795_jpeg_die("decoded file truncated");
    fclose(f);
    remove(raw.c_str());
    return 1;
}
static inline int jpeg_start_decompress(jpeg_decompress_struct* c) { c->output_width = c->image_width; c->output_height = c->image_height; c->output_scanline = 0; return 1; }
static inline unsigned jpeg_read_scanlines(jpeg_decompress_struct* c, JSAMPARRAY rows, unsigned) {
    size_t stride = (size_t)c->output_width * c->num_components;
    memcpy(rows[0], c->rt795_pixels->data() + stride * c->output_scanline, stride);
    c->output_scanline++;
    return 1;
}
static inline int jpeg_finish_decompress(jpeg_decompress_struct*) { return 1; }
static inline void jpeg_destroy_decompress(jpeg_decompress_struct* c) { delete c->rt795_pixels; c->rt795_pixels = nullptr; }
#endif
