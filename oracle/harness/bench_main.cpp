// Reference-arm bench harness over the UNMODIFIED reference sources (TEST INFRASTRUCTURE ONLY; our own code).
//
// Replaces /root/reference/src/main.cpp at link time.  Parses the scene and runs the public prologue of
// Scene::renderScene ONCE (src/Scene.cpp:296-323: Perlin, ComputeObjectTransformations, smooth normals, one BVH per
// object), then renders camera 0 `frames` times exactly as renderScene does (src/Scene.cpp:336-356: a fresh Image, the
// reference's 8 hard-coded std::threads on Scene::ThreadedRendering, join) and prints, per frame,
//   RT795_FRAME <index> seconds <render-only wall seconds> queries <BVHMethods::FindIntersection calls>
// The query count comes from the link-time wrapper in count_wrap.cpp.  `#define private public` only widens access in
// this translation unit (Scene::ThreadedRendering is private); the reference objects are compiled unmodified.
#define private public
#include "Scene.h"
#include "Camera.h"
#include "BVH.h"
#include "Shape.h"
#include "Helper.h"
#include "Perlin.h"
#include "Instance.h"
#include "Image.h"
#undef private
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>

Scene* pScene;
extern "C" unsigned long long rt795_ray_queries(int reset);

int main(int argc, char* argv[])
{
    if (argc < 3) { fprintf(stderr, "usage: %s scene.xml frames [out.f32]\n", argv[0]); return 2; }
    int frames = atoi(argv[2]);
    pScene = new Scene(argv[1]);
    Scene* s = pScene;
    s->perlin = new Perlin();
    Transforming::ComputeObjectTransformations(s->objects, s->instances, s->translations, s->scalings, s->rotations, s->composites);
    for (size_t i = 0; i < s->vertices.size(); i++) s->vertexNormals.push_back(Eigen::Vector3f{0, 0, 0});
    for (size_t i = 0; i < s->objects.size(); i++) s->objects[i]->ComputeSmoothNormals();
    for (size_t i = 0; i < s->vertexNormals.size(); i++) s->vertexNormals[i] = s->vertexNormals[i].normalized();
    for (size_t i = 0; i < s->objects.size(); i++) s->objects[i]->bvh = new BVH(s->objects[i]);
    Camera* cam = s->cameras[0];
    for (int f = 0; f < frames; f++) {
        Image image(cam->imgPlane.nx, cam->imgPlane.ny);
        rt795_ray_queries(1);
        auto t0 = std::chrono::steady_clock::now();
        std::thread th[8];
        for (int k = 0; k < 8; k++) th[k] = std::thread(&Scene::ThreadedRendering, s, k, std::ref(image), cam);
        for (int k = 0; k < 8; k++) th[k].join();
        double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        printf("RT795_FRAME %d seconds %.6f queries %llu\n", f, secs, rt795_ray_queries(0));
        fflush(stdout);
        if (f == frames - 1 && argc > 3) image.saveImage(argv[3]);
    }
    return 0;
}
