// Link-time replacement for /root/reference/src/Image.cpp (TEST INFRASTRUCTURE ONLY; our own code).
//
// Same class, same float framebuffer layout (`_data[(row*width+col)*3 + c]`, reference src/Image.cpp:17-24),
// but saveImage() dumps the pre-tone-map float32 radiance verbatim instead of the reference's clamped P3
// PPM / half-float EXR (src/Image.cpp:62-107), because parity is judged on un-clamped float radiance.
// File format "<ImageName>.f32": int32 width, int32 height, then width*height*3 float32.
//
// It also brackets the threaded render: the reference constructs Image immediately before spawning its 8
// worker threads (src/Scene.cpp:336-347) and calls saveImage right after joining them (:349-361), so
// ctor->saveImage wall time is the hot path alone (no parse, no BVH build).  Printed to stderr as
// "RT795_RENDER_SECONDS <s>" for bench.py's reference arm.
#include "Image.h"
#include <chrono>
#include <cstdio>
#include <string>

static std::chrono::steady_clock::time_point g_t0;

Image::Image(int width, int height) : width(width), height(height)
{
    _data = new float[(size_t)width * height * 3]();
    g_t0 = std::chrono::steady_clock::now();
}

void Image::setPixelValue(int col, int row, const Eigen::Vector3f& pixelColor)
{
    size_t startIndex = ((size_t)row * width + col) * 3;
    _data[startIndex] = pixelColor[0];
    _data[startIndex + 1] = pixelColor[1];
    _data[startIndex + 2] = pixelColor[2];
}

void Image::saveImage(const char* imageName)
{
    double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - g_t0).count();
    fprintf(stderr, "RT795_RENDER_SECONDS %.6f\n", secs);
    std::string path = std::string(imageName) + ".f32";
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) { fprintf(stderr, "oracle Image: cannot open %s\n", path.c_str()); return; }
    int hdr[2] = {width, height};
    fwrite(hdr, sizeof(int), 2, f);
    fwrite(_data, sizeof(float), (size_t)width * height * 3, f);
    fclose(f);
}

bool Image::IsPNG(const char*) { return false; }
void Image::SavePng(const char*) {}
void Image::SaveExr(const char*) {}
