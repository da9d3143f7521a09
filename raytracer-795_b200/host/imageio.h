// Image file I/O of the host shell (cold path): what the reference does through libpng / tinyexr
// (src/Texture.cpp:185-300, src/Helper.cpp:345-413) and its own PPM writer (src/Image.cpp:62-103).
// Own code on top of zlib only: PNG (non-interlaced, 8/16-bit, grey/RGB/palette +alpha) -> RGBA8 rows as the
// reference asks libpng for; OpenEXR scanline files (NONE / ZIPS / ZIP compression, half/float/uint channels)
// -> RGBA float as tinyexr's LoadEXR returns; writers: P3 PPM (for "*.png*" names, as the reference), half-float
// B,G,R EXR, raw .f32 dump.  JPEG is not supported.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace rt795 {

struct LoadedImage {
    int width = 0, height = 0;
    int format = 0;                      // RT_IMG_RGBA8 / RT_IMG_RGBA32F
    std::vector<unsigned char> bytes;    // RGBA8 or RGBA float32
};

bool NameHasExt(const char* name, const char* ext3);   // the reference's ".png"/".exr" substring scan (src/Texture.cpp:133-183)
bool LoadPng(const std::string& path, LoadedImage* out, std::string* err);
bool LoadExr(const std::string& path, LoadedImage* out, std::string* err);
bool LoadJpeg(const std::string& path, LoadedImage* out, std::string* err);   // baseline / extended sequential (host/jpeg.h)
bool LoadImageFile(const std::string& path, LoadedImage* out, std::string* err);

bool SavePpmP3(const std::string& path, const float* rgb, int nx, int ny, std::string* err);
bool SaveExrHalfBGR(const std::string& path, const float* rgb, int nx, int ny, std::string* err);
bool SaveF32(const std::string& path, const float* rgb, int nx, int ny, std::string* err);

uint16_t FloatToHalf(float f);
float HalfToFloat(uint16_t h);

}  // namespace rt795
