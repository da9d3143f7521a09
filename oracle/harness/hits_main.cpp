// Primary-hit dump harness over the UNMODIFIED reference sources (TEST INFRASTRUCTURE ONLY; our own code).
//
// Replaces /root/reference/src/main.cpp at link time.  Repeats the public prologue of Scene::renderScene
// (src/Scene.cpp:296-323: Perlin, ComputeObjectTransformations, smooth normals, one BVH per object) and then,
// for every pixel of camera 0, calls Camera::getPrimaryRay (src/Camera.cpp:63-72) and
// BVHMethods::FindIntersection (src/Helper.cpp:18-80) and writes one record per pixel.
//
// The reference hit record has no primitive id (src/defs.h:13-22).  Leaf hits return primitives[i]->matIndex
// (src/BVH.cpp:170) and instance hits are overwritten with the instance's matIndex (src/Helper.cpp:68), so the
// harness relabels: primitive k of object i gets matIndex = base_i + k (in BVH-independent, parse order: spheres
// and standalone triangles have one primitive, meshes one per face in face order), instance j gets
// matIndex = -(j+1).  `#define private public` only widens access in this translation unit.
//
// Output "<out>": int32 nx, ny, then per pixel {int32 full, int32 label, float t, float px,py,pz, float nx,ny,nz}.
#define private public
#include "Scene.h"
#include "Camera.h"
#include "BVH.h"
#include "Shape.h"
#include "Helper.h"
#include "Perlin.h"
#include "Instance.h"
#undef private
#include <cstdio>
#include <map>

Scene* pScene;

int main(int argc, char* argv[])
{
    if (argc < 3) { fprintf(stderr, "usage: %s scene.xml out.hits\n", argv[0]); return 2; }
    pScene = new Scene(argv[1]);
    Scene* s = pScene;
    s->perlin = new Perlin();
    Transforming::ComputeObjectTransformations(s->objects, s->instances, s->translations, s->scalings, s->rotations, s->composites);
    for (size_t i = 0; i < s->vertices.size(); i++) s->vertexNormals.push_back(Eigen::Vector3f{0, 0, 0});
    for (size_t i = 0; i < s->objects.size(); i++) s->objects[i]->ComputeSmoothNormals();
    for (size_t i = 0; i < s->vertexNormals.size(); i++) s->vertexNormals[i] = s->vertexNormals[i].normalized();
    for (size_t i = 0; i < s->objects.size(); i++) s->objects[i]->bvh = new BVH(s->objects[i]);

    // Relabel.  Mesh BVHs hold the mesh's own Triangle* (src/Shape.cpp:236-242); sphere / triangle BVHs hold a
    // fresh copy (src/Shape.cpp:50-53, 157-160), reachable through the BVH's primitive list.
    int base = 1;
    for (size_t i = 0; i < s->objects.size(); i++) {
        Mesh* mesh = dynamic_cast<Mesh*>(s->objects[i]);
        if (mesh) {
            for (size_t k = 0; k < mesh->faces.size(); k++) mesh->faces[k]->matIndex = base + (int)k;
            base += (int)mesh->faces.size();
        } else {
            s->objects[i]->bvh->primitives[0]->matIndex = base;
            base += 1;
        }
    }
    for (size_t j = 0; j < s->instances.size(); j++) s->instances[j]->matIndex = -((int)j + 1);

    Camera* cam = s->cameras[0];
    int nx = cam->imgPlane.nx, ny = cam->imgPlane.ny;
    FILE* f = fopen(argv[2], "wb");
    if (!f) { fprintf(stderr, "cannot open %s\n", argv[2]); return 1; }
    int hdr[2] = {nx, ny};
    fwrite(hdr, sizeof(int), 2, f);
    for (int row = 0; row < ny; row++) {
        for (int col = 0; col < nx; col++) {
            Ray ray = cam->getPrimaryRay(col, row);
            ReturnVal r = BVHMethods::FindIntersection(ray, s->objects, s->instances);
            int full = r.full ? 1 : 0;
            int label = r.full ? r.matIndex : 0;
            float rec[7] = {0, 0, 0, 0, 0, 0, 0};
            if (r.full) {
                rec[0] = (r.point - ray.origin).dot(ray.direction);
                rec[1] = r.point[0]; rec[2] = r.point[1]; rec[3] = r.point[2];
                rec[4] = r.normal[0]; rec[5] = r.normal[1]; rec[6] = r.normal[2];
            }
            fwrite(&full, sizeof(int), 1, f);
            fwrite(&label, sizeof(int), 1, f);
            fwrite(rec, sizeof(float), 7, f);
        }
    }
    fclose(f);
    return 0;
}
