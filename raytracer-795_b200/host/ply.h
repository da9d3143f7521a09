// PLY reader for the mesh surface the reference consumes through happly (src/Parser.h:1049-1097):
// element "vertex" with scalar properties x,y,z (+ optional u,v) and element "face" with one list property of
// vertex indices (triangles or quads).  ascii and binary_little_endian / binary_big_endian.  Own code.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <sstream>
#include <string>
#include <vector>

namespace rt795 {

struct PlyMesh {
    std::vector<float> xyz;                 // 3 per vertex (happly returns double; the reference narrows to float)
    std::vector<float> uv;                  // 2 per vertex, empty if the file has no "u" property
    std::vector<std::vector<int64_t>> faces;
};

namespace plydetail {
enum Ty { T_I8, T_U8, T_I16, T_U16, T_I32, T_U32, T_F32, T_F64, T_BAD };
struct Prop { std::string name; bool is_list = false; std::string type, count_type; Ty ty = T_BAD, cty = T_BAD; };
struct Elem { std::string name; size_t count = 0; std::vector<Prop> props; };

inline int TypeSize(const std::string& t) {
    if (t == "char" || t == "uchar" || t == "int8" || t == "uint8") return 1;
    if (t == "short" || t == "ushort" || t == "int16" || t == "uint16") return 2;
    if (t == "int" || t == "uint" || t == "float" || t == "int32" || t == "uint32" || t == "float32") return 4;
    if (t == "double" || t == "float64") return 8;
    return 0;
}
inline Ty TypeOf(const std::string& t) {
    if (t == "char" || t == "int8") return T_I8;
    if (t == "uchar" || t == "uint8") return T_U8;
    if (t == "short" || t == "int16") return T_I16;
    if (t == "ushort" || t == "uint16") return T_U16;
    if (t == "int" || t == "int32") return T_I32;
    if (t == "uint" || t == "uint32") return T_U32;
    if (t == "float" || t == "float32") return T_F32;
    if (t == "double" || t == "float64") return T_F64;
    return T_BAD;
}
inline int SizeOf(Ty t) { static const int sz[] = {1, 1, 2, 2, 4, 4, 4, 8, 0}; return sz[t]; }
// little-endian fast path (the common case): no per-value string compares
inline double ReadLE(const unsigned char* p, Ty t) {
    switch (t) {
        case T_I8: { int8_t v; memcpy(&v, p, 1); return v; }
        case T_U8: return *p;
        case T_I16: { int16_t v; memcpy(&v, p, 2); return v; }
        case T_U16: { uint16_t v; memcpy(&v, p, 2); return v; }
        case T_I32: { int32_t v; memcpy(&v, p, 4); return v; }
        case T_U32: { uint32_t v; memcpy(&v, p, 4); return v; }
        case T_F32: { float v; memcpy(&v, p, 4); return v; }
        case T_F64: { double v; memcpy(&v, p, 8); return v; }
        default: return 0;
    }
}
inline double ReadBin(const unsigned char* p, const std::string& t, bool swap) {
    unsigned char b[8];
    int n = TypeSize(t);
    for (int i = 0; i < n; i++) b[i] = swap ? p[n - 1 - i] : p[i];
    if (t == "char" || t == "int8") { int8_t v; memcpy(&v, b, 1); return v; }
    if (t == "uchar" || t == "uint8") { uint8_t v; memcpy(&v, b, 1); return v; }
    if (t == "short" || t == "int16") { int16_t v; memcpy(&v, b, 2); return v; }
    if (t == "ushort" || t == "uint16") { uint16_t v; memcpy(&v, b, 2); return v; }
    if (t == "int" || t == "int32") { int32_t v; memcpy(&v, b, 4); return v; }
    if (t == "uint" || t == "uint32") { uint32_t v; memcpy(&v, b, 4); return v; }
    if (t == "float" || t == "float32") { float v; memcpy(&v, b, 4); return v; }
    double v; memcpy(&v, b, 8); return v;
}
}  // namespace plydetail

inline bool LoadPly(const std::string& path, PlyMesh* out, std::string* err) {
    using namespace plydetail;
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) { *err = "cannot open PLY " + path; return false; }
    std::string data;
    char buf[1 << 16];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) data.append(buf, n);
    fclose(f);

    size_t pos = 0;
    auto getline = [&](std::string& line) {
        if (pos >= data.size()) return false;
        size_t e = data.find('\n', pos);
        if (e == std::string::npos) e = data.size();
        line = data.substr(pos, e - pos);
        if (!line.empty() && line.back() == '\r') line.pop_back();
        pos = e + 1;
        return true;
    };
    std::string line;
    if (!getline(line) || line != "ply") { *err = "not a PLY file: " + path; return false; }
    std::string format;
    std::vector<Elem> elems;
    while (getline(line)) {
        std::istringstream is(line);
        std::string tok;
        is >> tok;
        if (tok == "format") is >> format;
        else if (tok == "element") { Elem e; is >> e.name >> e.count; elems.push_back(e); }
        else if (tok == "property") {
            if (elems.empty()) { *err = "PLY property before element"; return false; }
            Prop p; std::string t; is >> t;
            if (t == "list") { p.is_list = true; is >> p.count_type >> p.type >> p.name; }
            else { p.type = t; is >> p.name; }
            p.ty = TypeOf(p.type); p.cty = TypeOf(p.count_type);
            elems.back().props.push_back(p);
        } else if (tok == "end_header") break;
    }
    bool ascii = format == "ascii";
    bool swap = format == "binary_big_endian";
    if (!ascii && format != "binary_little_endian" && !swap) { *err = "unsupported PLY format " + format; return false; }

    std::istringstream as;
    if (ascii) as.str(data.substr(pos));
    const unsigned char* p = (const unsigned char*)data.data() + pos;
    const unsigned char* end = (const unsigned char*)data.data() + data.size();

    for (const Elem& e : elems) {
        int ix = -1, iy = -1, iz = -1, iu = -1, iv = -1;
        for (size_t k = 0; k < e.props.size(); k++) {
            const std::string& nm = e.props[k].name;
            if (nm == "x") ix = k; else if (nm == "y") iy = k; else if (nm == "z") iz = k;
            else if (nm == "u") iu = k; else if (nm == "v") iv = k;
        }
        bool isv = e.name == "vertex", isf = e.name == "face";
        if (isv) { out->xyz.resize(e.count * 3); if (iu >= 0 && iv >= 0) out->uv.resize(e.count * 2); }
        if (isf) out->faces.resize(e.count);
        std::vector<double> scal(e.props.size());
        for (size_t i = 0; i < e.count; i++) {
            for (size_t k = 0; k < e.props.size(); k++) {
                const Prop& pr = e.props[k];
                if (pr.is_list) {
                    size_t cnt;
                    if (ascii) { double c; if (!(as >> c)) { *err = "PLY truncated"; return false; } cnt = (size_t)c; }
                    else { int cs = SizeOf(pr.cty); if (cs == 0 || p + cs > end) { *err = "PLY truncated"; return false; } cnt = (size_t)(swap ? ReadBin(p, pr.count_type, true) : ReadLE(p, pr.cty)); p += cs; }
                    // a list cannot be longer than what is left of the file (1 byte / 1 token per item at least): checked BEFORE allocating
                    if (!ascii && cnt > (size_t)(end - p)) { *err = "PLY truncated"; return false; }
                    if (ascii && cnt > (size_t)1 << 24) { *err = "PLY: unreasonable list length"; return false; }
                    std::vector<int64_t> lst(cnt);
                    int ts = SizeOf(pr.ty);
                    if (!ascii && (ts == 0 || p + (size_t)ts * cnt > end)) { *err = "PLY truncated"; return false; }
                    for (size_t j = 0; j < cnt; j++) {
                        if (ascii) { double v; if (!(as >> v)) { *err = "PLY truncated"; return false; } lst[j] = (int64_t)v; }
                        else { lst[j] = (int64_t)(swap ? ReadBin(p, pr.type, true) : ReadLE(p, pr.ty)); p += ts; }
                    }
                    if (isf && (pr.name == "vertex_indices" || pr.name == "vertex_index")) out->faces[i] = std::move(lst);
                } else {
                    if (ascii) { if (!(as >> scal[k])) { *err = "PLY truncated"; return false; } }
                    else { int ts = SizeOf(pr.ty); if (ts == 0 || p + ts > end) { *err = "PLY truncated / bad type"; return false; } scal[k] = swap ? ReadBin(p, pr.type, true) : ReadLE(p, pr.ty); p += ts; }
                }
            }
            if (isv) {
                if (ix < 0 || iy < 0 || iz < 0) { *err = "PLY vertex element lacks x/y/z"; return false; }
                out->xyz[i * 3 + 0] = (float)scal[ix]; out->xyz[i * 3 + 1] = (float)scal[iy]; out->xyz[i * 3 + 2] = (float)scal[iz];
                if (!out->uv.empty()) { out->uv[i * 2 + 0] = (float)scal[iu]; out->uv[i * 2 + 1] = (float)scal[iv]; }
            }
        }
    }
    return true;
}

}  // namespace rt795
