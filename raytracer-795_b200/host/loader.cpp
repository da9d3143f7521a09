// Host shell: scene XML (+PLY, +PNG/EXR) -> prepared, flattened RtSceneDesc.
//
// Re-implements, from the grammar in SURVEY.md Appendix B, what the reference does in
//   Parser::Parse*                         src/Parser.h:17-1315      (XML -> scene objects)
//   Transforming::ComputeObjectTransformations  src/Helper.cpp:135-226    (model / inverse / inverse-transpose)
//   Scene::renderScene prologue            src/Scene.cpp:303-318     (smooth vertex normals)
// and flattens the result into the plain arrays of include/rt795.h.  Cold path, pure host C++, no CUDA.
// Quirks of the reference parser that change what is rendered are reproduced and cited inline.
#include "../../include/rt795.h"
#include "imageio.h"
#include "ply.h"
#include "xml.h"

#include <sys/stat.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <new>
#include <stdexcept>
#include <cstring>
#include <string>
#include <vector>

namespace rt795 {
void SetLastError(const std::string& msg);   // api_error.cpp
}
using namespace rt795;

namespace {

struct Mat4 {   // row-major 4x4 float
    float m[4][4];
    static Mat4 Identity() { Mat4 r; memset(r.m, 0, sizeof r.m); for (int i = 0; i < 4; i++) r.m[i][i] = 1.f; return r; }
};
Mat4 Mul(const Mat4& a, const Mat4& b) {
    Mat4 r;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            float s = a.m[i][0] * b.m[0][j];
            for (int k = 1; k < 4; k++) s += a.m[i][k] * b.m[k][j];
            r.m[i][j] = s;
        }
    return r;
}
Mat4 Translate(float x, float y, float z) { Mat4 r = Mat4::Identity(); r.m[0][3] = x; r.m[1][3] = y; r.m[2][3] = z; return r; }
Mat4 Scale(float x, float y, float z) { Mat4 r = Mat4::Identity(); r.m[0][0] = x; r.m[1][1] = y; r.m[2][2] = z; return r; }
// glm::rotate(radians(angle), axis): Rodrigues with the axis normalised (src/Helper.cpp:166-170)
Mat4 Rotate(float deg, float ax, float ay, float az) {
    float a = deg * 0.01745329251994329576923690768489f;
    float c = std::cos(a), s = std::sin(a);
    float len = std::sqrt(ax * ax + ay * ay + az * az);
    float x = ax / len, y = ay / len, z = az / len;
    float tx = (1 - c) * x, ty = (1 - c) * y, tz = (1 - c) * z;
    Mat4 r = Mat4::Identity();
    r.m[0][0] = c + tx * x;     r.m[0][1] = ty * x - s * z; r.m[0][2] = tz * x + s * y;
    r.m[1][0] = tx * y + s * z; r.m[1][1] = c + ty * y;     r.m[1][2] = tz * y - s * x;
    r.m[2][0] = tx * z - s * y; r.m[2][1] = ty * z + s * x; r.m[2][2] = c + tz * z;
    return r;
}
bool Inverse(const Mat4& M, Mat4* out) {
    // cofactor expansion in float, as glm::inverse does
    const float* a = &M.m[0][0];
    float inv[16];
    inv[0] = a[5]*a[10]*a[15] - a[5]*a[11]*a[14] - a[9]*a[6]*a[15] + a[9]*a[7]*a[14] + a[13]*a[6]*a[11] - a[13]*a[7]*a[10];
    inv[4] = -a[4]*a[10]*a[15] + a[4]*a[11]*a[14] + a[8]*a[6]*a[15] - a[8]*a[7]*a[14] - a[12]*a[6]*a[11] + a[12]*a[7]*a[10];
    inv[8] = a[4]*a[9]*a[15] - a[4]*a[11]*a[13] - a[8]*a[5]*a[15] + a[8]*a[7]*a[13] + a[12]*a[5]*a[11] - a[12]*a[7]*a[9];
    inv[12] = -a[4]*a[9]*a[14] + a[4]*a[10]*a[13] + a[8]*a[5]*a[14] - a[8]*a[6]*a[13] - a[12]*a[5]*a[10] + a[12]*a[6]*a[9];
    inv[1] = -a[1]*a[10]*a[15] + a[1]*a[11]*a[14] + a[9]*a[2]*a[15] - a[9]*a[3]*a[14] - a[13]*a[2]*a[11] + a[13]*a[3]*a[10];
    inv[5] = a[0]*a[10]*a[15] - a[0]*a[11]*a[14] - a[8]*a[2]*a[15] + a[8]*a[3]*a[14] + a[12]*a[2]*a[11] - a[12]*a[3]*a[10];
    inv[9] = -a[0]*a[9]*a[15] + a[0]*a[11]*a[13] + a[8]*a[1]*a[15] - a[8]*a[3]*a[13] - a[12]*a[1]*a[11] + a[12]*a[3]*a[9];
    inv[13] = a[0]*a[9]*a[14] - a[0]*a[10]*a[13] - a[8]*a[1]*a[14] + a[8]*a[2]*a[13] + a[12]*a[1]*a[10] - a[12]*a[2]*a[9];
    inv[2] = a[1]*a[6]*a[15] - a[1]*a[7]*a[14] - a[5]*a[2]*a[15] + a[5]*a[3]*a[14] + a[13]*a[2]*a[7] - a[13]*a[3]*a[6];
    inv[6] = -a[0]*a[6]*a[15] + a[0]*a[7]*a[14] + a[4]*a[2]*a[15] - a[4]*a[3]*a[14] - a[12]*a[2]*a[7] + a[12]*a[3]*a[6];
    inv[10] = a[0]*a[5]*a[15] - a[0]*a[7]*a[13] - a[4]*a[1]*a[15] + a[4]*a[3]*a[13] + a[12]*a[1]*a[7] - a[12]*a[3]*a[5];
    inv[14] = -a[0]*a[5]*a[14] + a[0]*a[6]*a[13] + a[4]*a[1]*a[14] - a[4]*a[2]*a[13] - a[12]*a[1]*a[6] + a[12]*a[2]*a[5];
    inv[3] = -a[1]*a[6]*a[11] + a[1]*a[7]*a[10] + a[5]*a[2]*a[11] - a[5]*a[3]*a[10] - a[9]*a[2]*a[7] + a[9]*a[3]*a[6];
    inv[7] = a[0]*a[6]*a[11] - a[0]*a[7]*a[10] - a[4]*a[2]*a[11] + a[4]*a[3]*a[10] + a[8]*a[2]*a[7] - a[8]*a[3]*a[6];
    inv[11] = -a[0]*a[5]*a[11] + a[0]*a[7]*a[9] + a[4]*a[1]*a[11] - a[4]*a[3]*a[9] - a[8]*a[1]*a[7] + a[8]*a[3]*a[5];
    inv[15] = a[0]*a[5]*a[10] - a[0]*a[6]*a[9] - a[4]*a[1]*a[10] + a[4]*a[2]*a[9] + a[8]*a[1]*a[6] - a[8]*a[2]*a[5];
    float det = a[0] * inv[0] + a[1] * inv[4] + a[2] * inv[8] + a[3] * inv[12];
    if (det == 0.f || det != det) return false;
    float id = 1.0f / det;
    for (int i = 0; i < 16; i++) (&out->m[0][0])[i] = inv[i] * id;
    return true;
}

enum TKind { T_TRANSLATION, T_SCALING, T_ROTATION, T_COMPOSITE };
struct TRef { TKind kind; int id; };

struct ParsedObject {
    int kind;                 // RT_OBJ_*
    int geom;
    int id = 0;
    int mat_index = 0;
    std::vector<int> textures;
    int tex_offset = 0;
    std::vector<TRef> transforms;
    float blur[3] = {0, 0, 0};
    bool is_instance = false;
    bool reset_transform = false;
    int base_object = -1;     // instances: index (into objects_) of the base mesh object
    bool is_mesh_elem = false;
    bool is_light = false;    // <LightSphere> / <LightMesh>
    float radiance[3] = {0, 0, 0};
    Mat4 model;
};

bool StartsWith(const char* s, const char* lit) { return s && strncmp(s, lit, strlen(lit)) == 0; }

}  // namespace

struct RtHostScene {
    RtSceneDesc desc;
    std::vector<float> vertices, normals, texcoords;
    std::vector<int32_t> tris;
    std::vector<RtMesh> meshes;
    std::vector<RtSphere> spheres;
    std::vector<RtObject> objects;
    std::vector<RtMaterial> materials;
    std::vector<RtTexture> textures;
    std::vector<LoadedImage> image_store;
    std::vector<RtImage> images;
    std::vector<RtLight> lights;
    std::vector<RtCameraDesc> cameras;
    std::vector<std::string> image_names;
    std::vector<std::string> deps;        // every file the scene was read from (XML first): the blob cache's staleness check
    std::vector<RtRenderOpts> cam_opts;   // per camera: <Renderer> / <RendererParams> (extensions for the path tracer)
    std::vector<RtTonemap> cam_tonemap;   // per camera: <Tonemap> (extension)
};

namespace {

// Points the descriptor's array fields at the vectors the host scene owns (after a parse or a blob load).
void BindDesc(RtHostScene* hs) {
    hs->images.clear();
    for (auto& li : hs->image_store) {
        RtImage im; im.width = li.width; im.height = li.height; im.format = li.format; im.pixels = li.bytes.data();
        hs->images.push_back(im);
    }
    RtSceneDesc& d = hs->desc;
    d.num_vertices = (int)hs->vertices.size() / 3; d.vertices = hs->vertices.data(); d.vertex_normals = hs->normals.data();
    d.num_texcoords = (int)hs->texcoords.size() / 2; d.texcoords = hs->texcoords.data();
    d.num_triangles = (int)hs->tris.size() / 3; d.tri_indices = hs->tris.data();
    d.num_meshes = (int)hs->meshes.size(); d.meshes = hs->meshes.data();
    d.num_spheres = (int)hs->spheres.size(); d.spheres = hs->spheres.data();
    d.num_objects = (int)hs->objects.size(); d.objects = hs->objects.data();
    d.num_materials = (int)hs->materials.size(); d.materials = hs->materials.data();
    d.num_textures = (int)hs->textures.size(); d.textures = hs->textures.data();
    d.num_images = (int)hs->images.size(); d.images = hs->images.data();
    d.num_lights = (int)hs->lights.size(); d.lights = hs->lights.data();
}

class Loader {
public:
    Loader(const char* path, RtHostScene* hs) : path_(path), hs_(hs) {}
    bool Run(std::string* err) {
        XmlDoc doc;
        {
            FILE* probe = fopen(path_.c_str(), "rb");
            if (!probe) { io_error_ = true; *err = "cannot open " + path_; return false; }
            fclose(probe);
        }
        if (!doc.LoadFile(path_.c_str(), err)) return false;
        hs_->deps.push_back(path_);
        root_ = doc.root();
        // directory prefix exactly as src/Parser.h:488-498: everything up to and including the last '/',
        // and, when there is no '/', the FIRST CHARACTER of the path (lastIndex defaults to 0) - SURVEY A.11.
        size_t last = 0;
        for (size_t i = 0; i < path_.size(); i++) if (path_[i] == '/') last = i;
        base_ = path_.substr(0, last + 1);
        return SceneAttributes(err) && Cameras(err) && Brdfs(err) && Materials(err) && Textures(err) && Transformations(err) &&
               Vertices(err) && TexCoords(err) && Objects(err) && Lights(err) && Prepare(err) && Flatten(err);
    }
    bool io_error() const { return io_error_; }

private:
    std::string path_, base_;
    RtHostScene* hs_;
    const XmlElem* root_ = nullptr;
    bool io_error_ = false;

    int max_depth_ = 1;
    float shadow_eps_ = 0.002f, int_eps_ = 0.001f;
    float bg_[3] = {0, 0, 0}, ambient_[3] = {0, 0, 0};
    struct Brdf { int type, id, exponent; };
    std::vector<Brdf> brdfs_;
    std::vector<std::string> image_paths_;
    struct Xf { int id; float v[3]; float angle; Mat4 composite; };
    std::vector<Xf> translations_, scalings_, rotations_, composites_;
    std::vector<ParsedObject> objects_, instances_;
    int mesh_start_ = 0;
    int env_light_ = -1;

    bool Fail(std::string* err, const std::string& m) { *err = m + " (" + path_ + ")"; return false; }
    static bool Vec3(const XmlElem* e, float* v) { return e && e->GetText() && sscanf(e->GetText(), "%f %f %f", v, v + 1, v + 2) == 3; }

    // ---- src/Parser.h:17-50
    bool SceneAttributes(std::string*) {
        if (auto e = root_->FirstChildElement("MaxRecursionDepth")) e->QueryIntText(&max_depth_);
        if (auto e = root_->FirstChildElement("BackgroundColor")) Vec3(e, bg_);
        if (auto e = root_->FirstChildElement("ShadowRayEpsilon")) e->QueryFloatText(&shadow_eps_);
        if (auto e = root_->FirstChildElement("IntersectionTestEpsilon")) e->QueryFloatText(&int_eps_);
        return true;
    }

    // ---- src/Parser.h:52-164
    bool Cameras(std::string* err) {
        auto cams = root_->FirstChildElement("Cameras");
        if (!cams) return Fail(err, "<Cameras> missing");
        for (auto c = cams->FirstChildElement("Camera"); c; c = c->NextSiblingElement("Camera")) {
            RtCameraDesc cd;
            memset(&cd, 0, sizeof cd);
            if (auto a = c->FindAttributePrefix("handedness")) cd.left_handed = StartsWith(a->value.c_str(), "left");
            if (auto e = c->FirstChildElement("FocusDistance")) { e->QueryFloatText(&cd.focus_distance); cd.is_dof = 1; }
            if (auto e = c->FirstChildElement("ApertureSize")) e->QueryFloatText(&cd.aperture_size);
            cd.num_samples = 1;
            if (auto e = c->FirstChildElement("NumSamples")) e->QueryIntText(&cd.num_samples);
            if (!Vec3(c->FirstChildElement("Position"), cd.position)) return Fail(err, "Camera/Position missing");
            bool have_gaze = false;
            if (auto e = c->FirstChildElement("Gaze")) have_gaze = Vec3(e, cd.gaze);
            if (auto e = c->FirstChildElement("GazePoint")) {
                float gp[3];
                if (Vec3(e, gp)) { for (int k = 0; k < 3; k++) cd.gaze[k] = gp[k] - cd.position[k]; have_gaze = true; }
            }
            if (!have_gaze) return Fail(err, "Camera needs Gaze or GazePoint");
            if (!Vec3(c->FirstChildElement("Up"), cd.up)) return Fail(err, "Camera/Up missing");
            auto nd = c->FirstChildElement("NearDistance");
            if (!nd || !nd->QueryFloatText(&cd.distance)) return Fail(err, "Camera/NearDistance missing");
            auto res = c->FirstChildElement("ImageResolution");
            if (!res || !res->GetText() || sscanf(res->GetText(), "%d %d", &cd.nx, &cd.ny) != 2) return Fail(err, "Camera/ImageResolution missing");
            auto nm = c->FirstChildElement("ImageName");
            if (!nm || !nm->GetText()) return Fail(err, "Camera/ImageName missing");
            bool have_plane = false;
            if (auto e = c->FirstChildElement("NearPlane"))
                have_plane = e->GetText() && sscanf(e->GetText(), "%f %f %f %f", &cd.left, &cd.right, &cd.bottom, &cd.top) == 4;
            if (auto e = c->FirstChildElement("FovY")) {
                float fov;
                if (e->QueryFloatText(&fov)) {
                    fov = (fov * 0.5f) * (M_PI / 180.0f);                     // float = float * double, as the reference
                    float aspect = (float)cd.nx / (float)cd.ny;
                    float y = tan(fov) * cd.distance;
                    float x = aspect * y;
                    cd.top = y; cd.bottom = -y; cd.left = -x; cd.right = x;
                    have_plane = true;
                }
            }
            if (!have_plane) return Fail(err, "Camera needs NearPlane or FovY");
            hs_->cameras.push_back(cd);
            hs_->image_names.push_back(nm->GetText());
            // ---- extensions beyond src/Parser.h (SURVEY App. B "not parsed by this snapshot"), in the course's XML vocabulary:
            //   <Renderer>PathTracing</Renderer>  <RendererParams>NextEventEstimation RussianRoulette ImportanceSampling</RendererParams>
            //   <Tonemap><TMO>Photographic</TMO><TMOOptions>key burn%</TMOOptions><Saturation>s</Saturation><Gamma>g</Gamma></Tonemap>
            RtRenderOpts ro; memset(&ro, 0, sizeof ro);
            ro.integrator = RT_INTEGRATOR_REFERENCE; ro.seed = 795; ro.pt_nee = 1; ro.pt_russian_roulette = 1; ro.pt_importance = 1;
            if (auto e = c->FirstChildElement("Renderer")) {
                const char* t = e->GetText();
                if (t && StartsWith(t, "PathTracing")) {
                    ro.integrator = RT_INTEGRATOR_PATH; ro.pt_nee = 0; ro.pt_russian_roulette = 0; ro.pt_importance = 0;   // opted into one by one
                    if (auto pe = c->FirstChildElement("RendererParams")) {
                        std::string ps = pe->GetText() ? pe->GetText() : "";
                        if (ps.find("NextEventEstimation") != std::string::npos) ro.pt_nee = 1;
                        if (ps.find("RussianRoulette") != std::string::npos) ro.pt_russian_roulette = 1;
                        if (ps.find("ImportanceSampling") != std::string::npos) ro.pt_importance = 1;
                    }
                }
            }
            RtTonemap tm; memset(&tm, 0, sizeof tm);
            if (auto e = c->FirstChildElement("Tonemap")) {
                tm.enabled = 1; tm.key = 0.18f; tm.burn_percent = 1.0f; tm.saturation = 1.0f; tm.gamma = 2.2f;
                if (auto o = e->FirstChildElement("TMOOptions")) { const char* t = o->GetText(); if (t) sscanf(t, "%f %f", &tm.key, &tm.burn_percent); }
                if (auto o = e->FirstChildElement("Saturation")) o->QueryFloatText(&tm.saturation);
                if (auto o = e->FirstChildElement("Gamma")) o->QueryFloatText(&tm.gamma);
            }
            hs_->cam_opts.push_back(ro);
            hs_->cam_tonemap.push_back(tm);
        }
        if (hs_->cameras.empty()) return Fail(err, "no <Camera>");
        return true;
    }

    // ---- src/Parser.h:166-302.  Order of the list matters only through "last match wins" in Materials().
    bool Brdfs(std::string* err) {
        auto b = root_->FirstChildElement("BRDFs");
        if (!b) return true;
        struct Kind { const char* tag; const char* flag; int plain, flagged; };
        const Kind kinds[] = {{"ModifiedBlinnPhong", "normalized", RT_BRDF_MBP, RT_BRDF_MBPN},
                              {"OriginalBlinnPhong", nullptr, RT_BRDF_OBP, RT_BRDF_OBP},
                              {"ModifiedPhong", "normalized", RT_BRDF_MP, RT_BRDF_MPN},
                              {"OriginalPhong", nullptr, RT_BRDF_OP, RT_BRDF_OP},
                              {"TorranceSparrow", "kdfresnel", RT_BRDF_TS, RT_BRDF_TSF}};
        for (const Kind& k : kinds)
            for (auto e = b->FirstChildElement(k.tag); e; e = e->NextSiblingElement(k.tag)) {
                Brdf br{k.plain, 0, 0};
                e->QueryIntAttribute("id", &br.id);
                auto ex = e->FirstChildElement("Exponent");
                if (!ex || !ex->QueryIntText(&br.exponent)) return Fail(err, std::string(k.tag) + "/Exponent missing");
                if (k.flag) if (auto a = e->FindAttributePrefix(k.flag)) if (StartsWith(a->value.c_str(), "true")) br.type = k.flagged;
                brdfs_.push_back(br);
            }
        return true;
    }

    // ---- src/Parser.h:304-474
    bool Materials(std::string* err) {
        auto ms = root_->FirstChildElement("Materials");
        if (!ms) return Fail(err, "<Materials> missing");
        for (auto m = ms->FirstChildElement("Material"); m; m = m->NextSiblingElement("Material")) {
            RtMaterial mat;
            memset(&mat, 0, sizeof mat);
            bool degamma = false;
            if (auto a = m->FindAttributePrefix("degamma")) degamma = StartsWith(a->value.c_str(), "true");
            int brdf = -1;
            m->QueryIntAttribute("BRDF", &brdf);
            mat.brdf = RT_BRDF_NONE;
            if (brdf != -1)
                for (auto& b : brdfs_) if (b.id == brdf) { mat.phong_exp = b.exponent; mat.brdf = b.type; }
            if (!Vec3(m->FirstChildElement("AmbientReflectance"), mat.ambient)) return Fail(err, "Material/AmbientReflectance missing");
            if (!Vec3(m->FirstChildElement("DiffuseReflectance"), mat.diffuse)) return Fail(err, "Material/DiffuseReflectance missing");
            if (!Vec3(m->FirstChildElement("SpecularReflectance"), mat.specular)) return Fail(err, "Material/SpecularReflectance missing");
            if (degamma) {
                float g = 2.2f;
                for (int k = 0; k < 3; k++) { mat.ambient[k] = pow(mat.ambient[k], g); mat.diffuse[k] = pow(mat.diffuse[k], g); mat.specular[k] = pow(mat.specular[k], g); }
            }
            if (auto e = m->FirstChildElement("Roughness")) { e->QueryFloatText(&mat.roughness); mat.is_rough = 1; }
            if (auto e = m->FirstChildElement("MirrorReflectance")) Vec3(e, mat.mirror);
            if (auto e = m->FirstChildElement("PhongExponent")) e->QueryIntText(&mat.phong_exp);
            mat.type = RT_MAT_NORMAL;
            if (auto a = m->FindAttributePrefix("type")) {
                const char* v = a->value.c_str();
                if (StartsWith(v, "dielectric")) mat.type = RT_MAT_DIELECTRIC;
                else if (StartsWith(v, "conductor")) mat.type = RT_MAT_CONDUCTOR;
                else if (StartsWith(v, "mirror")) mat.type = RT_MAT_MIRROR;
            }
            if (auto e = m->FirstChildElement("RefractionIndex")) e->QueryFloatText(&mat.refraction_index);
            if (auto e = m->FirstChildElement("AbsorptionIndex")) e->QueryFloatText(&mat.absorption_index);
            if (auto e = m->FirstChildElement("AbsorptionCoefficient")) Vec3(e, mat.absorption_coeff);
            if (auto e = m->FirstChildElement("Emission")) Vec3(e, mat.emission);          // extension: object lights (path tracer)
            hs_->materials.push_back(mat);
        }
        return true;
    }

    // ---- src/Parser.h:476-605.  dm / interpolation / normalizer / noiseScale / bumpFactor / nc / imageId / isImage are
    // declared OUTSIDE the <TextureMap> loop: unspecified fields inherit the previous map's values.
    bool Textures(std::string* err) {
        auto ts = root_->FirstChildElement("Textures");
        if (!ts) return true;
        if (auto imgs = ts->FirstChildElement("Images"))
            for (auto im = imgs->FirstChildElement("Image"); im; im = im->NextSiblingElement("Image"))
                image_paths_.push_back(base_ + (im->GetText() ? im->GetText() : ""));
        bool is_image = false;
        int image_id = 0, normalizer = 255;
        float noise_scale = 1, bump_factor = 1;
        int dm = RT_DM_NONE, nc = RT_NC_LINEAR, interp = RT_INTERP_NN;
        for (auto t = ts->FirstChildElement("TextureMap"); t; t = t->NextSiblingElement("TextureMap")) {
            if (auto a = t->FindAttributePrefix("type")) is_image = StartsWith(a->value.c_str(), "image");
            if (auto e = t->FirstChildElement("ImageId")) e->QueryIntText(&image_id);
            if (auto e = t->FirstChildElement("DecalMode")) {
                const char* s = e->GetText();
                if (StartsWith(s, "blend_kd")) dm = RT_DM_BLEND_KD;
                else if (StartsWith(s, "replace_kd")) dm = RT_DM_REPLACE_KD;
                else if (StartsWith(s, "replace_all")) dm = RT_DM_REPLACE_ALL;
                else if (StartsWith(s, "bump_normal")) dm = RT_DM_BUMP_NORMAL;
                else if (StartsWith(s, "replace_normal")) dm = RT_DM_REPLACE_NORMAL;
                else if (StartsWith(s, "replace_background")) dm = RT_DM_REPLACE_BACKGROUND;
            }
            if (auto e = t->FirstChildElement("NoiseConversion")) nc = StartsWith(e->GetText(), "absval") ? RT_NC_ABSVAL : RT_NC_LINEAR;
            if (auto e = t->FirstChildElement("Interpolation")) {
                if (StartsWith(e->GetText(), "nearest")) interp = RT_INTERP_NN;
                else if (StartsWith(e->GetText(), "bilinear")) interp = RT_INTERP_BILINEAR;
            }
            if (auto e = t->FirstChildElement("Normalizer")) e->QueryIntText(&normalizer);
            if (auto e = t->FirstChildElement("NoiseScale")) e->QueryFloatText(&noise_scale);
            if (auto e = t->FirstChildElement("BumpFactor")) e->QueryFloatText(&bump_factor);
            RtTexture tx;
            memset(&tx, 0, sizeof tx);
            tx.decal_mode = dm; tx.interpolation = interp; tx.normalizer = normalizer; tx.bump_factor = bump_factor;
            if (is_image) {
                if (image_id < 1 || image_id > (int)image_paths_.size()) return Fail(err, "TextureMap/ImageId out of range");
                tx.type = RT_TEX_IMAGE;
                tx.noise_conversion = RT_NC_NONE;
                tx.noise_scale = 1;
                tx.image = ImageIndex(image_paths_[image_id - 1], err);
                if (tx.image < 0) return false;
            } else {
                tx.type = RT_TEX_PERLIN; tx.noise_conversion = nc; tx.noise_scale = noise_scale; tx.image = -1;
            }
            hs_->textures.push_back(tx);
        }
        return true;
    }
    std::vector<std::string> loaded_paths_;
    int ImageIndex(const std::string& p, std::string* err) {
        for (size_t i = 0; i < loaded_paths_.size(); i++) if (loaded_paths_[i] == p) return (int)i;
        LoadedImage li;
        if (!LoadImageFile(p, &li, err)) { io_error_ = true; return -1; }
        hs_->deps.push_back(p);
        hs_->image_store.push_back(std::move(li));
        loaded_paths_.push_back(p);
        return (int)loaded_paths_.size() - 1;
    }

    // ---- src/Parser.h:607-682
    bool Transformations(std::string* err) {
        auto ts = root_->FirstChildElement("Transformations");
        if (!ts) return true;
        for (auto e = ts->FirstChildElement("Translation"); e; e = e->NextSiblingElement("Translation")) {
            Xf x{}; e->QueryIntAttribute("id", &x.id);
            if (!Vec3(e, x.v)) return Fail(err, "bad <Translation>");
            translations_.push_back(x);
        }
        for (auto e = ts->FirstChildElement("Scaling"); e; e = e->NextSiblingElement("Scaling")) {
            Xf x{}; e->QueryIntAttribute("id", &x.id);
            if (!Vec3(e, x.v)) return Fail(err, "bad <Scaling>");
            scalings_.push_back(x);
        }
        for (auto e = ts->FirstChildElement("Rotation"); e; e = e->NextSiblingElement("Rotation")) {
            Xf x{}; e->QueryIntAttribute("id", &x.id);
            if (!e->GetText() || sscanf(e->GetText(), "%f %f %f %f", &x.angle, x.v, x.v + 1, x.v + 2) != 4) return Fail(err, "bad <Rotation>");
            rotations_.push_back(x);
        }
        for (auto e = ts->FirstChildElement("Composite"); e; e = e->NextSiblingElement("Composite")) {
            Xf x{}; e->QueryIntAttribute("id", &x.id);
            float* m = &x.composite.m[0][0];   // 16 numbers, row-major text (read into glm[col][row] = row-major here)
            if (!e->GetText() || sscanf(e->GetText(), "%f %f %f %f %f %f %f %f %f %f %f %f %f %f %f %f", m, m + 1, m + 2, m + 3, m + 4, m + 5,
                                        m + 6, m + 7, m + 8, m + 9, m + 10, m + 11, m + 12, m + 13, m + 14, m + 15) != 16)
                return Fail(err, "bad <Composite>");
            composites_.push_back(x);
        }
        return true;
    }

    // ---- src/Parser.h:684-767: atof on whitespace-separated tokens
    static void ParseFloats(const char* s, std::vector<float>* out) {
        const char* p = s;
        for (;;) {
            while (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r') p++;
            if (!*p) break;
            char* end;
            double v = strtod(p, &end);
            if (end == p) { while (*p && *p != ' ' && *p != '\t' && *p != '\n' && *p != '\r') p++; v = 0; } else p = end;   // atof of junk = 0
            out->push_back((float)v);
        }
    }
    bool Vertices(std::string*) {
        if (auto e = root_->FirstChildElement("VertexData")) if (e->GetText()) {
            ParseFloats(e->GetText(), &hs_->vertices);
            hs_->vertices.resize(hs_->vertices.size() / 3 * 3);
        }
        return true;
    }
    bool TexCoords(std::string*) {
        if (auto e = root_->FirstChildElement("TexCoordData")) if (e->GetText()) {
            ParseFloats(e->GetText(), &hs_->texcoords);
            hs_->texcoords.resize(hs_->texcoords.size() / 2 * 2);
        }
        return true;
    }

    // ---- src/Parser.h:769-796.  'c' is honoured only where the cursor lands on it: the skip loop stops at s/t/r
    // and end of string only, so a composite is seen only as the first character.
    static void ObjectTransformations(const char* s, std::vector<TRef>* out) {
        size_t cur = 0;
        while (s[cur] != '\0') {
            char ch = s[cur];
            if (ch == 't') out->push_back({T_TRANSLATION, atoi(s + cur + 1)});
            else if (ch == 's') out->push_back({T_SCALING, atoi(s + cur + 1)});
            else if (ch == 'r') out->push_back({T_ROTATION, atoi(s + cur + 1)});
            else if (ch == 'c') out->push_back({T_COMPOSITE, atoi(s + cur + 1)});
            cur++;
            while (s[cur] != 's' && s[cur] != 't' && s[cur] != 'r' && s[cur] != '\0') cur++;
        }
    }
    // common children of Sphere/Triangle/Mesh (src/Parser.h:818-861 and twins)
    bool CommonObject(const XmlElem* o, ParsedObject* po, bool with_textures, std::string* err) {
        o->QueryIntAttribute("id", &po->id);
        auto m = o->FirstChildElement("Material");
        if (!m || !m->QueryIntText(&po->mat_index)) return Fail(err, o->name + "/Material missing");
        if (auto e = o->FirstChildElement("Transformations")) if (e->GetText()) ObjectTransformations(e->GetText(), &po->transforms);
        if (with_textures)
            if (auto e = o->FirstChildElement("Textures")) if (const char* s = e->GetText()) {
                bool two = strchr(s, ' ') != nullptr;
                int a = 0, b = 0;
                int n = two ? sscanf(s, "%d %d", &a, &b) : sscanf(s, "%d", &a);
                if (n >= 1) po->textures.push_back(a);
                if (two && n >= 2) po->textures.push_back(b);
            }
        if (auto e = o->FirstChildElement("MotionBlur")) Vec3(e, po->blur);
        return true;
    }

    // ---- src/Parser.h:798-1195
    bool Objects(std::string* err) {
        auto objs = root_->FirstChildElement("Objects");
        if (!objs) return Fail(err, "<Objects> missing");
        // <LightSphere> / <LightMesh> (course HW7 object lights, pages/Page7.md:9-13; not parsed by the reference snapshot): a <Sphere> /
        // <Mesh> with a <Radiance> child.  They follow the plain elements of their kind in the top-level order.
        for (const char* tag : {"Sphere", "LightSphere"})
        for (auto o = objs->FirstChildElement(tag); o; o = o->NextSiblingElement(tag)) {
            ParsedObject po; po.kind = RT_OBJ_SPHERE;
            if (!CommonObject(o, &po, true, err)) return false;
            if (tag[0] == 'L' && (!Vec3(o->FirstChildElement("Radiance"), po.radiance) || !(po.is_light = true))) return Fail(err, "LightSphere needs <Radiance>");
            int c = 0; float R = 0;
            auto ce = o->FirstChildElement("Center"); auto re = o->FirstChildElement("Radius");
            if (!ce || !ce->QueryIntText(&c) || !re || !re->QueryFloatText(&R)) return Fail(err, "Sphere needs Center and Radius");
            if (c < 1 || c * 3 > (int)hs_->vertices.size()) return Fail(err, "Sphere/Center vertex index out of range");
            RtSphere s; memcpy(s.center, &hs_->vertices[(c - 1) * 3], 12); s.radius = R;
            po.geom = (int)hs_->spheres.size();
            hs_->spheres.push_back(s);
            objects_.push_back(po);
        }
        for (auto o = objs->FirstChildElement("Triangle"); o; o = o->NextSiblingElement("Triangle")) {
            ParsedObject po; po.kind = RT_OBJ_MESH;
            if (!CommonObject(o, &po, true, err)) return false;
            int a, b, c;
            auto ie = o->FirstChildElement("Indices");
            if (!ie || !ie->GetText() || sscanf(ie->GetText(), "%d %d %d", &a, &b, &c) != 3) return Fail(err, "Triangle/Indices missing");
            RtMesh mesh{(int)hs_->tris.size() / 3, 1, 1};   // standalone triangles are ALWAYS smooth-shaded
            hs_->tris.insert(hs_->tris.end(), {a - 1, b - 1, c - 1});
            po.geom = (int)hs_->meshes.size();
            hs_->meshes.push_back(mesh);
            objects_.push_back(po);
        }
        mesh_start_ = (int)objects_.size();
        for (const char* tag : {"Mesh", "LightMesh"})
        for (auto o = objs->FirstChildElement(tag); o; o = o->NextSiblingElement(tag)) {
            ParsedObject po; po.kind = RT_OBJ_MESH; po.is_mesh_elem = tag[0] == 'M';    // MeshInstance looks its base up among <Mesh> elements only
            if (!CommonObject(o, &po, true, err)) return false;
            if (tag[0] == 'L' && (!Vec3(o->FirstChildElement("Radiance"), po.radiance) || !(po.is_light = true))) return Fail(err, "LightMesh needs <Radiance>");
            bool smooth = false;
            if (auto a = o->FindAttributePrefix("shadingMode")) smooth = StartsWith(a->value.c_str(), "smooth");
            auto fe = o->FirstChildElement("Faces");
            if (!fe) return Fail(err, "Mesh/Faces missing");
            RtMesh mesh{(int)hs_->tris.size() / 3, 0, smooth ? 1 : 0};
            if (auto ply = fe->FindAttributePrefix("plyFile")) {
                PlyMesh pm;
                if (!LoadPly(base_ + ply->value, &pm, err)) { io_error_ = true; return false; }
                hs_->deps.push_back(base_ + ply->value);
                int texture_offset = (int)hs_->texcoords.size() / 2 + 1;          // Parser.h:1051
                hs_->texcoords.insert(hs_->texcoords.end(), pm.uv.begin(), pm.uv.end());
                int vertex_count = (int)hs_->vertices.size() / 3 + 1;             // Parser.h:1065
                for (auto& f : pm.faces) {
                    if (f.size() == 4) {                                           // quads split (0,1,2),(2,3,0): Parser.h:1068-1078
                        hs_->tris.insert(hs_->tris.end(), {(int)f[0] + vertex_count - 1, (int)f[1] + vertex_count - 1, (int)f[2] + vertex_count - 1});
                        hs_->tris.insert(hs_->tris.end(), {(int)f[2] + vertex_count - 1, (int)f[3] + vertex_count - 1, (int)f[0] + vertex_count - 1});
                    } else if (f.size() >= 3) {
                        hs_->tris.insert(hs_->tris.end(), {(int)f[0] + vertex_count - 1, (int)f[1] + vertex_count - 1, (int)f[2] + vertex_count - 1});
                    } else return Fail(err, "PLY face with fewer than 3 vertices");
                }
                hs_->vertices.insert(hs_->vertices.end(), pm.xyz.begin(), pm.xyz.end());
                po.tex_offset = texture_offset - vertex_count;                      // Parser.h:1102
            } else {
                int vertex_offset = 0, texture_offset = 0;
                fe->QueryIntAttribute("vertexOffset", &vertex_offset);
                fe->QueryIntAttribute("textureOffset", &texture_offset);
                std::vector<float> idx;
                if (fe->GetText()) {
                    const char* p = fe->GetText();
                    std::vector<int> iv;
                    for (;;) {
                        while (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r') p++;
                        if (!*p) break;
                        iv.push_back(atoi(p) + vertex_offset - 1);
                        while (*p && *p != ' ' && *p != '\t' && *p != '\n' && *p != '\r') p++;
                    }
                    iv.resize(iv.size() / 3 * 3);
                    hs_->tris.insert(hs_->tris.end(), iv.begin(), iv.end());
                }
                po.tex_offset = texture_offset - vertex_offset;                     // Parser.h:1147
            }
            mesh.tri_count = (int)hs_->tris.size() / 3 - mesh.first_tri;
            if (mesh.tri_count == 0) return Fail(err, "Mesh has no faces");
            po.geom = (int)hs_->meshes.size();
            hs_->meshes.push_back(mesh);
            objects_.push_back(po);
        }
        for (auto o = objs->FirstChildElement("MeshInstance"); o; o = o->NextSiblingElement("MeshInstance")) {
            ParsedObject po; po.kind = RT_OBJ_MESH; po.is_instance = true;
            if (!CommonObject(o, &po, false, err)) return false;
            int base_id = 0;
            o->QueryIntAttribute("baseMeshId", &base_id);
            o->QueryBoolAttribute("resetTransform", &po.reset_transform);
            for (int i = mesh_start_; i < (int)objects_.size(); i++)               // last match wins (no break), Parser.h:1184-1188
                if (objects_[i].id == base_id && objects_[i].is_mesh_elem) po.base_object = i;
            if (po.base_object < 0) return Fail(err, "MeshInstance baseMeshId does not name a <Mesh>");
            const ParsedObject& base = objects_[po.base_object];
            po.geom = base.geom;
            po.textures = base.textures;      // textures live in the base mesh's BVH (src/BVH.cpp:58-59)
            po.tex_offset = base.tex_offset;
            instances_.push_back(po);
        }
        int nv = (int)hs_->vertices.size() / 3;
        for (int v : hs_->tris) if (v < 0 || v >= nv) return Fail(err, "face vertex index out of range");
        return true;
    }

    // ---- src/Parser.h:1197-1315: order point, directional, spot, area, environment
    bool Lights(std::string* err) {
        auto ls = root_->FirstChildElement("Lights");
        if (!ls) return Fail(err, "<Lights> missing");
        if (auto a = ls->FirstChildElement("AmbientLight")) Vec3(a, ambient_);
        auto mk = [](int type) { RtLight l; memset(&l, 0, sizeof l); l.type = type; l.image = -1; return l; };
        for (auto e = ls->FirstChildElement("PointLight"); e; e = e->NextSiblingElement("PointLight")) {
            RtLight l = mk(RT_LIGHT_POINT);
            if (!Vec3(e->FirstChildElement("Position"), l.position) || !Vec3(e->FirstChildElement("Intensity"), l.intensity)) return Fail(err, "bad <PointLight>");
            hs_->lights.push_back(l);
        }
        for (auto e = ls->FirstChildElement("DirectionalLight"); e; e = e->NextSiblingElement("DirectionalLight")) {
            RtLight l = mk(RT_LIGHT_DIRECTIONAL);
            if (!Vec3(e->FirstChildElement("Direction"), l.direction) || !Vec3(e->FirstChildElement("Radiance"), l.intensity)) return Fail(err, "bad <DirectionalLight>");
            hs_->lights.push_back(l);
        }
        for (auto e = ls->FirstChildElement("SpotLight"); e; e = e->NextSiblingElement("SpotLight")) {
            RtLight l = mk(RT_LIGHT_SPOT);
            auto ca = e->FirstChildElement("CoverageAngle"); auto fa = e->FirstChildElement("FalloffAngle");
            if (!Vec3(e->FirstChildElement("Position"), l.position) || !Vec3(e->FirstChildElement("Direction"), l.direction) ||
                !Vec3(e->FirstChildElement("Intensity"), l.intensity) || !ca || !ca->QueryFloatText(&l.coverage_deg) || !fa || !fa->QueryFloatText(&l.falloff_deg))
                return Fail(err, "bad <SpotLight>");
            hs_->lights.push_back(l);
        }
        for (auto e = ls->FirstChildElement("AreaLight"); e; e = e->NextSiblingElement("AreaLight")) {
            RtLight l = mk(RT_LIGHT_AREA);
            auto rad = e->FirstChildElement("Radiance");
            if (!rad) rad = e->FirstChildElement("Intensity");
            auto sz = e->FirstChildElement("Size");
            if (!Vec3(e->FirstChildElement("Position"), l.position) || !Vec3(e->FirstChildElement("Normal"), l.direction) || !Vec3(rad, l.intensity) ||
                !sz || !sz->QueryFloatText(&l.size))
                return Fail(err, "bad <AreaLight>");
            hs_->lights.push_back(l);
        }
        for (auto e = ls->FirstChildElement("SphericalDirectionalLight"); e; e = e->NextSiblingElement("SphericalDirectionalLight")) {
            RtLight l = mk(RT_LIGHT_ENVIRONMENT);
            int image_id = 0;
            auto ie = e->FirstChildElement("ImageId");
            if (!ie || !ie->QueryIntText(&image_id) || image_id < 1 || image_id > (int)image_paths_.size()) return Fail(err, "bad <SphericalDirectionalLight>/ImageId");
            l.image = ImageIndex(image_paths_[image_id - 1], err);
            if (l.image < 0) return false;
            env_light_ = (int)hs_->lights.size();
            hs_->lights.push_back(l);
        }
        return true;
    }

    // ---- src/Helper.cpp:135-226 + src/Scene.cpp:303-318
    bool Compose(const std::vector<TRef>& list, Mat4* out, std::string* err) {
        Mat4 M = Mat4::Identity();
        for (int j = (int)list.size() - 1; j >= 0; j--) {     // glm post-multiplication from the LAST listed to the first: M = T(n-1)...T1.T0
            int id = list[j].id;
            auto get = [&](std::vector<Xf>& v, const char* what) -> Xf* {
                if (id < 1 || id > (int)v.size()) { Fail(err, std::string("object references missing ") + what + " #" + std::to_string(id)); return nullptr; }
                return &v[id - 1];                             // positional, like translations[transformIndex-1]
            };
            switch (list[j].kind) {
                case T_TRANSLATION: { Xf* x = get(translations_, "Translation"); if (!x) return false; M = Mul(M, Translate(x->v[0], x->v[1], x->v[2])); } break;
                case T_SCALING: { Xf* x = get(scalings_, "Scaling"); if (!x) return false; M = Mul(M, Scale(x->v[0], x->v[1], x->v[2])); } break;
                case T_ROTATION: { Xf* x = get(rotations_, "Rotation"); if (!x) return false; M = Mul(M, Rotate(x->angle, x->v[0], x->v[1], x->v[2])); } break;
                case T_COMPOSITE: { Xf* x = get(composites_, "Composite"); if (!x) return false; M = x->composite; } break;   // REPLACES (Helper.cpp:173-175)
            }
        }
        *out = M;
        return true;
    }
    bool Prepare(std::string* err) {
        for (auto& o : objects_) if (!Compose(o.transforms, &o.model, err)) return false;
        for (auto& in : instances_) {
            if (!Compose(in.transforms, &in.model, err)) return false;
            if (!in.reset_transform) in.model = Mul(in.model, objects_[in.base_object].model);   // Helper.cpp:216-218
        }
        // smooth normals: unweighted sum of unit face normals per vertex over every standalone triangle and every
        // smooth mesh, in object order, then normalised; the array is GLOBAL (shared by all shapes), SURVEY A.4
        size_t nv = hs_->vertices.size() / 3;
        hs_->normals.assign(nv * 3, 0.f);
        const float* V = hs_->vertices.data();
        for (auto& o : objects_) {
            if (o.kind != RT_OBJ_MESH) continue;
            const RtMesh& m = hs_->meshes[o.geom];
            if (!m.smooth) continue;
            for (int t = m.first_tri; t < m.first_tri + m.tri_count; t++) {
                const int* ix = &hs_->tris[t * 3];
                const float *a = V + ix[0] * 3, *b = V + ix[1] * 3, *c = V + ix[2] * 3;
                float e1[3] = {c[0] - b[0], c[1] - b[1], c[2] - b[2]}, e2[3] = {a[0] - b[0], a[1] - b[1], a[2] - b[2]};
                float n[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
                float len = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
                if (len > 0) { n[0] /= len; n[1] /= len; n[2] /= len; }
                for (int k = 0; k < 3; k++) { float* d = &hs_->normals[ix[k] * 3]; d[0] += n[0]; d[1] += n[1]; d[2] += n[2]; }
            }
        }
        for (size_t i = 0; i < nv; i++) {
            float* n = &hs_->normals[i * 3];
            float len = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
            if (len > 0) { n[0] /= len; n[1] /= len; n[2] /= len; }
        }
        return true;
    }

    bool Flatten(std::string* err) {
        auto emit = [&](const ParsedObject& po) -> bool {
            RtObject o;
            memset(&o, 0, sizeof o);
            o.kind = po.kind; o.geom = po.geom; o.mat_index = po.mat_index; o.tex_offset = po.tex_offset;
            for (size_t k = 0; k < po.textures.size() && k < 2; k++) o.tex[k] = po.textures[k];
            Mat4 inv;
            if (!Inverse(po.model, &inv)) return Fail(err, "singular object transformation");
            for (int r = 0; r < 3; r++) for (int c = 0; c < 4; c++) o.inv[r * 4 + c] = inv.m[r][c];
            for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) o.inv_t[r * 3 + c] = inv.m[c][r];
            memcpy(o.blur, po.blur, 12);
            if (po.is_light) {
                memcpy(o.radiance, po.radiance, 12);
                RtLight l; memset(&l, 0, sizeof l);
                l.type = po.kind == RT_OBJ_SPHERE ? RT_LIGHT_SPHERE : RT_LIGHT_MESH; l.image = -1;
                l.object = (int)hs_->objects.size(); memcpy(l.intensity, po.radiance, 12);
                hs_->lights.push_back(l);                                          // after the <Lights> elements, in top-level order
            }
            if (o.mat_index < 1 || o.mat_index > (int)hs_->materials.size()) return Fail(err, "object Material index out of range");
            for (int k = 0; k < 2; k++) if (o.tex[k] < 0 || o.tex[k] > (int)hs_->textures.size()) return Fail(err, "object Textures id out of range");
            hs_->objects.push_back(o);
            return true;
        };
        for (auto& o : objects_) if (!emit(o)) return false;
        for (auto& in : instances_) if (!emit(in)) return false;

        // textured triangle meshes ALWAYS read uv (src/Shape.cpp:509-518); the reference would read out of bounds
        int ntc = (int)hs_->texcoords.size() / 2;
        for (auto& o : hs_->objects) {
            if (o.kind != RT_OBJ_MESH || (o.tex[0] == 0 && o.tex[1] == 0)) continue;
            const RtMesh& m = hs_->meshes[o.geom];
            for (int t = m.first_tri * 3; t < (m.first_tri + m.tri_count) * 3; t++) {
                int uvi = hs_->tris[t] + o.tex_offset;
                if (uvi < 0 || uvi >= ntc) return Fail(err, "textured mesh lacks texture coordinates for some vertex (SURVEY A.9)");
            }
        }
        RtSceneDesc& d = hs_->desc;
        memset(&d, 0, sizeof d);
        d.abi_version = RT795_ABI_VERSION;
        d.max_recursion_depth = max_depth_; d.shadow_ray_eps = shadow_eps_; d.int_test_eps = int_eps_;
        memcpy(d.background_color, bg_, 12); memcpy(d.ambient_light, ambient_, 12);
        d.env_light_index = env_light_;
        d.background_texture = -1;
        for (size_t i = 0; i < hs_->textures.size(); i++) if (hs_->textures[i].decal_mode == RT_DM_REPLACE_BACKGROUND) d.background_texture = (int)i;   // last wins, Scene.cpp:494-500
        BindDesc(hs_);
        return true;
    }
};

}  // namespace

// ---- scene blob (scene ingest).  The prepared, flattened scene written as one little-endian file so the
// next run skips XML / PLY / PNG / EXR parsing, transform composition and smooth-normal computation (reference:
// src/Parser.h:684-727, 1109-1143 text parsing with atof, src/Scene.cpp:303-318).  The blob records size + mtime of every
// source file it came from; a changed source makes it stale.
namespace {

const char kBlobMagic[8] = {'R', 'T', '7', '9', '5', 'S', 'B', '3'};

struct BlobWriter {
    FILE* f; bool ok = true;
    void Raw(const void* p, size_t n) { if (ok && n && fwrite(p, 1, n, f) != n) ok = false; }
    void U64(uint64_t v) { Raw(&v, 8); }
    void Str(const std::string& s) { U64(s.size()); Raw(s.data(), s.size()); }
    template <class T> void Vec(const std::vector<T>& v) { U64(v.size()); Raw(v.data(), v.size() * sizeof(T)); }
};
struct BlobReader {
    FILE* f; bool ok = true; uint64_t limit;
    void Raw(void* p, size_t n) { if (ok && n && fread(p, 1, n, f) != n) ok = false; }
    uint64_t U64() { uint64_t v = 0; Raw(&v, 8); return v; }
    bool Str(std::string* s) { uint64_t n = U64(); if (!ok || n > limit) return ok = false; s->resize(n); Raw(&(*s)[0], n); return ok; }
    template <class T> bool Vec(std::vector<T>* v) { uint64_t n = U64(); if (!ok || n > limit / sizeof(T)) return ok = false; v->resize(n); Raw(v->data(), n * sizeof(T)); return ok; }
};

// layout fingerprint: a blob written by a library with different struct layouts must not be read
uint64_t LayoutStamp() {
    const uint64_t sz[] = {sizeof(RtSceneDesc), sizeof(RtObject), sizeof(RtMaterial), sizeof(RtTexture), sizeof(RtLight), sizeof(RtCameraDesc),
                           sizeof(RtMesh), sizeof(RtSphere), sizeof(RtRenderOpts), sizeof(RtTonemap), (uint64_t)RT795_ABI_VERSION};
    uint64_t h = 1469598103934665603ull;
    for (uint64_t v : sz) { h ^= v; h *= 1099511628211ull; }
    return h;
}

bool StatFile(const std::string& path, uint64_t* size, int64_t* mtime_ns) {
    struct stat sb;
    if (stat(path.c_str(), &sb) != 0) return false;
    *size = (uint64_t)sb.st_size;
    *mtime_ns = (int64_t)sb.st_mtim.tv_sec * 1000000000ll + (int64_t)sb.st_mtim.tv_nsec;
    return true;
}

}  // namespace


extern "C" {

int rt_host_load_scene(const char* xml_path, RtHostScene** out) {
    if (!xml_path || !out) { SetLastError("rt_host_load_scene: null argument"); return RT_ERR_INVALID; }
    *out = nullptr;
    RtHostScene* hs = nullptr;
    try {                                                                // nothing may throw across the C boundary (a corrupt file can ask for absurd sizes)
        hs = new RtHostScene();
        Loader ld(xml_path, hs);
        std::string err;
        if (!ld.Run(&err)) {
            SetLastError("rt_host_load_scene: " + err);
            delete hs;
            return ld.io_error() ? RT_ERR_IO : RT_ERR_PARSE;
        }
    } catch (const std::bad_alloc&) {
        delete hs;
        SetLastError("rt_host_load_scene: out of host memory");
        return RT_ERR_OOM;
    } catch (const std::exception& e) {
        delete hs;
        SetLastError(std::string("rt_host_load_scene: malformed input (") + e.what() + ")");
        return RT_ERR_PARSE;
    }
    *out = hs;
    return RT_OK;
}
void rt_host_free_scene(RtHostScene* hs) { delete hs; }
int rt_host_save_scene_blob(const RtHostScene* hs, const char* blob_path) {
    if (!hs || !blob_path) { SetLastError("rt_host_save_scene_blob: null argument"); return RT_ERR_INVALID; }
    FILE* f = nullptr;
    try {
    std::string tmp = std::string(blob_path) + ".tmp";
    f = fopen(tmp.c_str(), "wb");
    if (!f) { SetLastError(std::string("rt_host_save_scene_blob: cannot write ") + tmp); return RT_ERR_IO; }
    BlobWriter w{f};
    w.Raw(kBlobMagic, 8); w.U64(LayoutStamp());
    w.U64(hs->deps.size());
    for (const std::string& d : hs->deps) {
        uint64_t size = 0; int64_t mt = 0;
        StatFile(d, &size, &mt);
        w.Str(d); w.U64(size); w.U64((uint64_t)mt);
    }
    {   // scalars only: the pointer fields are rebound on load, so they are written as NULL (blobs are byte-reproducible and carry no addresses)
        RtSceneDesc d = hs->desc;
        d.vertices = nullptr; d.vertex_normals = nullptr; d.texcoords = nullptr; d.tri_indices = nullptr; d.meshes = nullptr; d.spheres = nullptr;
        d.objects = nullptr; d.materials = nullptr; d.textures = nullptr; d.images = nullptr; d.lights = nullptr;
        w.Raw(&d, sizeof(RtSceneDesc));
    }
    w.Vec(hs->vertices); w.Vec(hs->normals); w.Vec(hs->texcoords); w.Vec(hs->tris);
    w.Vec(hs->meshes); w.Vec(hs->spheres); w.Vec(hs->objects); w.Vec(hs->materials); w.Vec(hs->textures); w.Vec(hs->lights); w.Vec(hs->cameras); w.Vec(hs->cam_opts); w.Vec(hs->cam_tonemap);
    w.U64(hs->image_names.size());
    for (const std::string& n : hs->image_names) w.Str(n);
    w.U64(hs->image_store.size());
    for (const LoadedImage& li : hs->image_store) { w.U64((uint64_t)li.width); w.U64((uint64_t)li.height); w.U64((uint64_t)li.format); w.Vec(li.bytes); }
    w.Raw(kBlobMagic, 8);                                                // trailer: a truncated file is rejected
    bool ok = w.ok;
    if (fclose(f) != 0) ok = false;
    f = nullptr;
    if (!ok || rename(tmp.c_str(), blob_path) != 0) { remove(tmp.c_str()); SetLastError(std::string("rt_host_save_scene_blob: write failed for ") + blob_path); return RT_ERR_IO; }
    return RT_OK;
    } catch (const std::exception& e) {                                    // e.g. bad_alloc building a path string: never across the C boundary
        if (f) fclose(f);
        SetLastError(std::string("rt_host_save_scene_blob: ") + e.what());
        return RT_ERR_OOM;
    }
}

int rt_host_load_scene_blob(const char* blob_path, int check_sources, RtHostScene** out) {
    if (!blob_path || !out) { SetLastError("rt_host_load_scene_blob: null argument"); return RT_ERR_INVALID; }
    *out = nullptr;
    FILE* f = fopen(blob_path, "rb");
    if (!f) { SetLastError(std::string("rt_host_load_scene_blob: cannot open ") + blob_path); return RT_ERR_IO; }
    uint64_t fsize = 0; int64_t fmt = 0;
    StatFile(blob_path, &fsize, &fmt);
    BlobReader r{f, true, fsize};
    RtHostScene* hs = nullptr;
    bool f_open = true;
    try {
    hs = new RtHostScene();
    auto fail = [&](const std::string& why) { fclose(f); f_open = false; delete hs; hs = nullptr; SetLastError(std::string("rt_host_load_scene_blob: ") + blob_path + ": " + why); return RT_ERR_PARSE; };
    char magic[8];
    r.Raw(magic, 8);
    if (!r.ok || memcmp(magic, kBlobMagic, 8) != 0) return fail("not a scene blob");
    if (r.U64() != LayoutStamp()) return fail("written by an incompatible library version");
    uint64_t ndeps = r.U64();
    if (!r.ok || ndeps > 1u << 20) return fail("corrupt header");
    for (uint64_t i = 0; i < ndeps; i++) {
        std::string path;
        if (!r.Str(&path)) return fail("corrupt header");
        uint64_t size = r.U64(); int64_t mt = (int64_t)r.U64();
        hs->deps.push_back(path);
        if (check_sources) {
            uint64_t now_size = 0; int64_t now_mt = 0;
            if (!StatFile(path, &now_size, &now_mt) || now_size != size || now_mt != mt) return fail("stale: " + path + " changed since the blob was written");
        }
    }
    r.Raw(&hs->desc, sizeof(RtSceneDesc));
    r.Vec(&hs->vertices); r.Vec(&hs->normals); r.Vec(&hs->texcoords); r.Vec(&hs->tris);
    r.Vec(&hs->meshes); r.Vec(&hs->spheres); r.Vec(&hs->objects); r.Vec(&hs->materials); r.Vec(&hs->textures); r.Vec(&hs->lights); r.Vec(&hs->cameras); r.Vec(&hs->cam_opts); r.Vec(&hs->cam_tonemap);
    uint64_t nn = r.U64();
    if (!r.ok || nn > 1u << 20) return fail("corrupt body");
    hs->image_names.resize(nn);
    for (auto& n : hs->image_names) if (!r.Str(&n)) return fail("corrupt body");
    uint64_t ni = r.U64();
    if (!r.ok || ni > 1u << 20) return fail("corrupt body");
    hs->image_store.resize(ni);
    for (auto& li : hs->image_store) {
        li.width = (int)r.U64(); li.height = (int)r.U64(); li.format = (int)r.U64();
        if (!r.Vec(&li.bytes)) return fail("corrupt body");
        size_t px = li.format == RT_IMG_RGBA8 ? 4 : 16;
        if (li.width < 1 || li.height < 1 || li.bytes.size() != (size_t)li.width * li.height * px) return fail("corrupt image");
    }
    r.Raw(magic, 8);
    if (!r.ok || memcmp(magic, kBlobMagic, 8) != 0) return fail("truncated");
    fclose(f); f_open = false;
    if (hs->cameras.size() != hs->image_names.size() || hs->cam_opts.size() != hs->cameras.size() || hs->cam_tonemap.size() != hs->cameras.size() || hs->vertices.size() % 3 || hs->tris.size() % 3 || hs->texcoords.size() % 2 ||
        (!hs->normals.empty() && hs->normals.size() != hs->vertices.size())) { delete hs; SetLastError(std::string("rt_host_load_scene_blob: ") + blob_path + ": inconsistent arrays"); return RT_ERR_PARSE; }
    hs->desc.abi_version = RT795_ABI_VERSION;
    BindDesc(hs);                                                         // rt_scene_create validates every index range again
    } catch (const std::exception& e) {
        if (f_open) fclose(f);
        delete hs;
        SetLastError(std::string("rt_host_load_scene_blob: ") + blob_path + ": " + e.what());
        return RT_ERR_PARSE;
    }
    *out = hs;
    return RT_OK;
}

int rt_host_load_scene_cached(const char* xml_path, const char* blob_path, RtHostScene** out, int* cache_hit) {
    if (!xml_path || !out) { SetLastError("rt_host_load_scene_cached: null argument"); return RT_ERR_INVALID; }
    std::string blob = blob_path ? std::string(blob_path) : std::string(xml_path) + ".rt795blob";
    if (cache_hit) *cache_hit = 0;
    if (rt_host_load_scene_blob(blob.c_str(), 1, out) == RT_OK) {
        if (!(*out)->deps.empty() && (*out)->deps[0] == xml_path) { if (cache_hit) *cache_hit = 1; return RT_OK; }
        rt_host_free_scene(*out); *out = nullptr;                         // a blob of some other scene
    }
    int rc = rt_host_load_scene(xml_path, out);
    if (rc != RT_OK) return rc;
    rt_host_save_scene_blob(*out, blob.c_str());                          // best effort: a read-only directory just means no cache
    return RT_OK;
}

const RtSceneDesc* rt_host_scene_desc(const RtHostScene* hs) { return hs ? &hs->desc : nullptr; }
int rt_host_num_cameras(const RtHostScene* hs) { return hs ? (int)hs->cameras.size() : 0; }
const RtCameraDesc* rt_host_camera(const RtHostScene* hs, int i) { return (hs && i >= 0 && i < (int)hs->cameras.size()) ? &hs->cameras[i] : nullptr; }
int rt_host_camera_render_opts(const RtHostScene* hs, int i, RtRenderOpts* out) {
    if (!hs || !out || i < 0 || i >= (int)hs->cam_opts.size()) { SetLastError("rt_host_camera_render_opts: bad argument"); return RT_ERR_INVALID; }
    *out = hs->cam_opts[i];
    return RT_OK;
}
int rt_host_camera_tonemap(const RtHostScene* hs, int i, RtTonemap* out) {
    if (!hs || !out || i < 0 || i >= (int)hs->cam_tonemap.size()) { SetLastError("rt_host_camera_tonemap: bad argument"); return RT_ERR_INVALID; }
    *out = hs->cam_tonemap[i];
    return RT_OK;
}
const char* rt_host_camera_image_name(const RtHostScene* hs, int i) { return (hs && i >= 0 && i < (int)hs->image_names.size()) ? hs->image_names[i].c_str() : nullptr; }

int rt_host_save_image(const char* image_name, const float* rgb, int nx, int ny) {
    if (!image_name || !rgb || nx <= 0 || ny <= 0) { SetLastError("rt_host_save_image: bad argument"); return RT_ERR_INVALID; }
    std::string err;
    bool ok = NameHasExt(image_name, "png") ? SavePpmP3(image_name, rgb, nx, ny, &err)      // src/Image.cpp:26-34
                                            : SaveExrHalfBGR(image_name, rgb, nx, ny, &err);
    if (!ok) { SetLastError("rt_host_save_image: " + err); return RT_ERR_IO; }
    return RT_OK;
}
int rt_host_save_f32(const char* path, const float* rgb, int nx, int ny) {
    if (!path || !rgb || nx <= 0 || ny <= 0) { SetLastError("rt_host_save_f32: bad argument"); return RT_ERR_INVALID; }
    std::string err;
    if (!SaveF32(path, rgb, nx, ny, &err)) { SetLastError("rt_host_save_f32: " + err); return RT_ERR_IO; }
    return RT_OK;
}

}  // extern "C"
