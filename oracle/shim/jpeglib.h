// Stand-in for libjpeg (not installed; TEST INFRASTRUCTURE ONLY).  JPEG textures are out of scope for the
// oracle scenes; every entry point the reference calls (src/Texture.cpp:191-237)
// exists so the file compiles, and aborts with a message if a JPEG is actually requested.
#ifndef RT795_ORACLE_SHIM_JPEGLIB_H
#define RT795_ORACLE_SHIM_JPEGLIB_H
#include <cstdio>
#include <cstdlib>
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
};
static inline void rt795_jpeg_unsupported() { fprintf(stderr, "jpeg shim: JPEG textures are not supported by the oracle build\n"); abort(); }
static inline jpeg_error_mgr* jpeg_std_error(jpeg_error_mgr* e) { return e; }
static inline void jpeg_create_decompress(jpeg_decompress_struct*) { rt795_jpeg_unsupported(); }
static inline void jpeg_stdio_src(jpeg_decompress_struct*, FILE*) {}
static inline int jpeg_read_header(jpeg_decompress_struct*, int) { return 0; }
static inline int jpeg_start_decompress(jpeg_decompress_struct*) { return 0; }
static inline unsigned jpeg_read_scanlines(jpeg_decompress_struct*, JSAMPARRAY, unsigned) { return 0; }
static inline int jpeg_finish_decompress(jpeg_decompress_struct*) { return 0; }
static inline void jpeg_destroy_decompress(jpeg_decompress_struct*) {}
#endif
