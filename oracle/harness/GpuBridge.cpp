// Route A of INTEGRATION.md, compiled for real: the UNMODIFIED reference sources (parser, scene classes, transform and
// smooth-normal prologue) with their render loop - the block src/Scene.cpp:336-356, 8 threads on ThreadedRendering - replaced by
// ONE call into librt795.so.  TEST INFRASTRUCTURE (it proves the drop-in boundary with the reference's OWN parser and data
// structures instead of this repo's loader); our own code, nothing copied.
//
// Links: the reference objects oracle/Makefile already builds (all of src/*.cpp except main.cpp and Image.cpp) + this file +
// oracle/harness/Image.cpp (raw float32 frame dump) + raytracer-795_b200/lib/librt795.so  ->  oracle/_ref/ref_raytracer_gpu.
// `ref_raytracer_gpu scene.xml` writes "<ImageName>.f32" exactly like oracle/_ref/ref_raytracer does, so the two can be diffed.
//
// A maintainer adding this file to the reference would make renderSceneGpu() a member of Scene and add getters for the few
// private fields read below (Sphere::cIndex/R, Triangle::p?Index, Mesh::faces, the cameras' and lights' constructor arguments,
// Texture::width/height); to leave every reference file untouched, this translation unit widens access with
// `#define private public` instead (as oracle/harness/hits_main.cpp does).
#define private public
#define protected public
#include "Scene.h"
#include "Camera.h"
#include "Shape.h"
#include "Helper.h"
#include "Perlin.h"
#include "Instance.h"
#include "Light.h"
#include "Material.h"
#include "Texture.h"
#include "Image.h"
#undef private
#undef protected
#include "rt795.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>

Scene* pScene;

namespace {

void Fail(const char* what) { fprintf(stderr, "GpuBridge: %s: %s\n", what, rt_last_error()); exit(1); }

// glm matrices are column-major (m[col][row]); the ABI wants rows.
void FillObject(RtObject& o, int kind, int geom, int matIndex, const std::vector<int>& tex, int texOffset, const glm::mat4& inv, const glm::mat4& invT,
                const glm::vec3& blur, bool isBlur) {
    memset(&o, 0, sizeof o);
    o.kind = kind; o.geom = geom; o.mat_index = matIndex;                                 // Shape::matIndex / Instance::matIndex (src/Helper.cpp:68)
    o.tex[0] = tex.size() > 0 ? tex[0] : 0; o.tex[1] = tex.size() > 1 ? tex[1] : 0;       // Shape::textures (src/BVH.cpp:58)
    o.tex_offset = texOffset;                                                             // Shape::textureOffset (src/Parser.h:1102, 1147)
    for (int r = 0; r < 3; r++) for (int c = 0; c < 4; c++) o.inv[r * 4 + c] = inv[c][r];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) o.inv_t[r * 3 + c] = invT[c][r];
    if (isBlur) { o.blur[0] = blur.x; o.blur[1] = blur.y; o.blur[2] = blur.z; }           // src/Helper.cpp:112-117
}

struct ImageStore {
    std::vector<std::vector<unsigned char>> bytes;
    std::vector<RtImage> images;
    std::map<Texture*, int> index;
    // Every pixel through the reference's own accessor (Texture::GetColorAtPixel, src/Texture.cpp:41-74), so PNG / JPEG / EXR
    // storage quirks (RGBA8 rows, the greyscale-JPEG neighbour read) arrive exactly as the reference's lookups see them.
    int Add(Texture* t) {
        auto it = index.find(t);
        if (it != index.end()) return it->second;
        const int w = t->width, h = t->height;
        RtImage im; im.width = w; im.height = h;
        if (t->_type == _exr) {
            im.format = RT_IMG_RGBA32F;
            bytes.emplace_back((size_t)w * h * 16);
            float* px = reinterpret_cast<float*>(bytes.back().data());
            for (int j = 0; j < h; j++) for (int i = 0; i < w; i++) {
                Eigen::Vector3f c = t->GetColorAtPixel(i, j);
                float* p = px + ((size_t)j * w + i) * 4; p[0] = c[0]; p[1] = c[1]; p[2] = c[2]; p[3] = 1.f;
            }
        } else {
            im.format = RT_IMG_RGBA8;
            bytes.emplace_back((size_t)w * h * 4);
            unsigned char* px = bytes.back().data();
            for (int j = 0; j < h; j++) for (int i = 0; i < w; i++) {
                Eigen::Vector3f c = t->GetColorAtPixel(i, j);
                unsigned char* p = px + ((size_t)j * w + i) * 4; p[0] = (unsigned char)c[0]; p[1] = (unsigned char)c[1]; p[2] = (unsigned char)c[2]; p[3] = 255;
            }
        }
        im.pixels = bytes.back().data();
        images.push_back(im);
        index[t] = (int)images.size() - 1;
        return index[t];
    }
};

void Copy3(float* d, const Eigen::Vector3f& v) { d[0] = v[0]; d[1] = v[1]; d[2] = v[2]; }

}  // namespace

// What `Scene::renderSceneGpu()` would be (route A).
void renderSceneGpu(Scene* s, int device)
{
    // ---- the reference's own prologue, unchanged (src/Scene.cpp:296-318); the per-object `new BVH(objects[i])` (:321-323) is not needed
    s->perlin = new Perlin();
    Transforming::ComputeObjectTransformations(s->objects, s->instances, s->translations, s->scalings, s->rotations, s->composites);
    for (size_t i = 0; i < s->vertices.size(); i++) s->vertexNormals.push_back(Eigen::Vector3f{0, 0, 0});
    for (size_t i = 0; i < s->objects.size(); i++) s->objects[i]->ComputeSmoothNormals();
    for (size_t i = 0; i < s->vertexNormals.size(); i++) s->vertexNormals[i] = s->vertexNormals[i].normalized();

    // ---- flatten
    RtSceneDesc d; memset(&d, 0, sizeof d);
    d.abi_version = RT795_ABI_VERSION;
    d.max_recursion_depth = s->maxRecursionDepth; d.shadow_ray_eps = s->shadowRayEps; d.int_test_eps = s->intTestEps;
    Copy3(d.background_color, s->backgroundColor); Copy3(d.ambient_light, s->ambientLight);
    d.env_light_index = s->environmentLightIndex; d.background_texture = s->backgroundTexture;
    std::vector<float> verts(s->vertices.size() * 3), norms(s->vertexNormals.size() * 3), uvs(s->textureCoordinates.size() * 2);
    for (size_t i = 0; i < s->vertices.size(); i++) Copy3(&verts[3 * i], s->vertices[i]);
    for (size_t i = 0; i < s->vertexNormals.size(); i++) Copy3(&norms[3 * i], s->vertexNormals[i]);
    for (size_t i = 0; i < s->textureCoordinates.size(); i++) { uvs[2 * i] = s->textureCoordinates[i][0]; uvs[2 * i + 1] = s->textureCoordinates[i][1]; }
    d.num_vertices = (int)s->vertices.size(); d.vertices = verts.data(); d.vertex_normals = norms.data();
    d.num_texcoords = (int)s->textureCoordinates.size(); d.texcoords = uvs.empty() ? nullptr : uvs.data();

    std::vector<int32_t> tri;
    std::vector<RtMesh> meshes;
    std::vector<RtSphere> spheres;
    std::vector<RtObject> objects;
    std::map<Shape*, int> mesh_of;                                                        // base mesh -> RtMesh index (instances share it, src/Helper.cpp:58)
    for (Shape* sh : s->objects) {                                                        // loop order of src/Helper.cpp:32-51
        RtObject o;
        if (Sphere* sp = dynamic_cast<Sphere*>(sh)) {
            RtSphere r; Copy3(r.center, s->vertices[sp->cIndex - 1]); r.radius = sp->R;
            spheres.push_back(r);
            FillObject(o, RT_OBJ_SPHERE, (int)spheres.size() - 1, sh->matIndex, sh->textures, 0, *sh->inverse_tMatrix, *sh->inverseTranspose_tMatrix, sh->blurTransformation, sh->isBlur);
        } else {
            RtMesh m; m.first_tri = (int)tri.size() / 3;
            if (Triangle* t = dynamic_cast<Triangle*>(sh)) {
                tri.push_back(t->p1Index - 1); tri.push_back(t->p2Index - 1); tri.push_back(t->p3Index - 1);
                m.tri_count = 1; m.smooth = t->isSmooth;                                  // always true by now (src/Shape.cpp:262-276)
            } else {
                Mesh* me = dynamic_cast<Mesh*>(sh);
                for (Triangle* f : me->faces) { tri.push_back(f->p1Index - 1); tri.push_back(f->p2Index - 1); tri.push_back(f->p3Index - 1); }
                m.tri_count = (int)me->faces.size(); m.smooth = me->faces.empty() ? 0 : me->faces[0]->isSmooth;
            }
            meshes.push_back(m);
            mesh_of[sh] = (int)meshes.size() - 1;
            bool textured = !sh->textures.empty();
            FillObject(o, RT_OBJ_MESH, (int)meshes.size() - 1, sh->matIndex, sh->textures, textured ? sh->textureOffset : 0, *sh->inverse_tMatrix, *sh->inverseTranspose_tMatrix,
                       sh->blurTransformation, sh->isBlur);
        }
        objects.push_back(o);
    }
    for (Instance* in : s->instances) {                                                   // src/Helper.cpp:53-73
        RtObject o;
        Shape* base = in->baseMesh;
        bool textured = !base->textures.empty();
        FillObject(o, RT_OBJ_MESH, mesh_of.at(base), in->matIndex, base->textures, textured ? base->textureOffset : 0, *in->inverse_tMatrix, *in->inverseTranspose_tMatrix,
                   in->blurTransformation, in->isBlur);
        objects.push_back(o);
    }
    d.num_triangles = (int)tri.size() / 3; d.tri_indices = tri.data();
    d.num_meshes = (int)meshes.size(); d.meshes = meshes.data();
    d.num_spheres = (int)spheres.size(); d.spheres = spheres.data();
    d.num_objects = (int)objects.size(); d.objects = objects.data();

    std::vector<RtMaterial> mats(s->materials.size());
    for (size_t i = 0; i < mats.size(); i++) {                                            // src/Material.h:10-33, enums value for value
        const Material* m = s->materials[i];
        RtMaterial& r = mats[i]; memset(&r, 0, sizeof r);
        r.type = m->type; r.brdf = m->_brdfType; r.phong_exp = m->phongExp; r.is_rough = m->isRough;
        Copy3(r.ambient, m->ambientRef); Copy3(r.diffuse, m->diffuseRef); Copy3(r.specular, m->specularRef); Copy3(r.mirror, m->mirrorRef);
        r.refraction_index = m->refractionIndex; r.absorption_index = m->absorptionIndex; r.roughness = m->roughness;
        Copy3(r.absorption_coeff, m->absorptionCoefficient);
    }
    d.num_materials = (int)mats.size(); d.materials = mats.data();

    ImageStore store;
    std::vector<RtTexture> texs(s->textures.size());
    for (size_t i = 0; i < texs.size(); i++) {                                            // src/Texture.h:15-31
        Texture* t = s->textures[i];
        RtTexture& r = texs[i]; memset(&r, 0, sizeof r);
        r.type = t->type; r.decal_mode = t->decalMode; r.interpolation = t->interpolation; r.noise_conversion = t->nc;
        r.normalizer = t->normalizer; r.noise_scale = t->noiseScale; r.bump_factor = t->bumpFactor;
        r.image = t->type == ImageTexture ? store.Add(t) : -1;
    }
    d.num_textures = (int)texs.size(); d.textures = texs.empty() ? nullptr : texs.data();

    std::vector<RtLight> lights(s->lights.size());
    for (size_t i = 0; i < lights.size(); i++) {                                          // src/Light.h, constructor arguments
        RtLight& r = lights[i]; memset(&r, 0, sizeof r);
        Light* l = s->lights[i];
        r.type = l->GetType();
        if (PointLight* p = dynamic_cast<PointLight*>(l)) { Copy3(r.position, p->position); Copy3(r.intensity, p->intensity); }
        else if (DirectionalLight* q = dynamic_cast<DirectionalLight*>(l)) { Copy3(r.direction, q->_direction); Copy3(r.intensity, q->_radiance); }
        else if (SpotLight* sp = dynamic_cast<SpotLight*>(l)) {
            Copy3(r.position, sp->_position); Copy3(r.direction, sp->_direction); Copy3(r.intensity, sp->_intensity);
            // the ctor stores HALF angles in radians (src/Light.cpp:332-333); the ABI carries the XML's full angles in degrees
            r.coverage_deg = sp->_coverage * 2.0f * 180.0f / (float)M_PI; r.falloff_deg = sp->_fall * 2.0f * 180.0f / (float)M_PI;
        } else if (AreaLight* a = dynamic_cast<AreaLight*>(l)) { Copy3(r.position, a->_position); Copy3(r.direction, a->_normal); Copy3(r.intensity, a->_radiance); r.size = a->_size; }
        else if (EnvironmentLight* e = dynamic_cast<EnvironmentLight*>(l)) r.image = store.Add(e->_image);
    }
    d.num_lights = (int)lights.size(); d.lights = lights.empty() ? nullptr : lights.data();
    d.num_images = (int)store.images.size(); d.images = store.images.empty() ? nullptr : store.images.data();
    d.bvh_builder = RT_BVH_AUTO;

    RtScene* gpu = nullptr;
    if (rt_scene_create(&d, device, &gpu) != RT_OK) Fail("rt_scene_create");

    // ---- per camera: the block src/Scene.cpp:330-362 with the thread spawn / join replaced by rt_render
    for (size_t i = 0; i < s->cameras.size(); i++) {
        Camera* cam = s->cameras[i];
        Image image(cam->imgPlane.nx, cam->imgPlane.ny);                                  // src/Scene.cpp:336
        RtCameraDesc c; memset(&c, 0, sizeof c);
        Copy3(c.position, cam->pos); Copy3(c.gaze, cam->gaze); Copy3(c.up, cam->up);      // the stored basis is orthonormal: re-deriving it (src/Camera.cpp:33-42) is the identity
        Eigen::Vector3f w = cam->right.cross(cam->up);                                    // w = -gaze, or +gaze for handedness="left"
        c.left_handed = w.dot(cam->gaze) > 0 ? 1 : 0;
        c.left = cam->imgPlane.left; c.right = cam->imgPlane.right; c.bottom = cam->imgPlane.bottom; c.top = cam->imgPlane.top; c.distance = cam->imgPlane.distance;
        c.nx = cam->imgPlane.nx; c.ny = cam->imgPlane.ny;
        c.num_samples = cam->totalSampleCount; c.focus_distance = cam->focusDistance; c.aperture_size = cam->apertureSize; c.is_dof = cam->isDof;
        RtRenderOpts o; memset(&o, 0, sizeof o);
        o.integrator = RT_INTEGRATOR_REFERENCE; o.seed = 795;
        if (rt_render(gpu, &c, &o, image._data) != RT_OK) Fail("rt_render");
        image.saveImage(cam->imageName);                                                  // src/Scene.cpp:361, unchanged
    }
    rt_scene_destroy(gpu);
}

int main(int argc, char* argv[])                                                          // src/main.cpp with renderScene() -> renderSceneGpu()
{
    if (argc < 2) { fprintf(stderr, "usage: %s scene.xml\n", argv[0]); return 2; }
    pScene = new Scene(argv[1]);
    renderSceneGpu(pScene, 0);
    return 0;
}
