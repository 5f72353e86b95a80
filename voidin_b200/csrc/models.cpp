// models.cpp — asset loaders of libbvh_cuda.so (include/bvh_cuda_models.h): the producers of the vertex / index arrays
// that voidin hands to MeshPool::add and from there to BvhBuilder (crates/pools/src/mesh/mod.rs:309-321).
//
//   bvh_cuda_model_load_obj   <- ObjModel::import      crates/app/src/models/mod.rs:20-57     tobj 4.0.0, GPU_LOAD_OPTIONS
//   bvh_cuda_model_load_gltf  <- GltfDocument::import  crates/app/src/models/gltf_model/mod.rs:103-155 (make_meshes),
//                                                      :166-207 (get_scene_instances / gather_instances_recursive),
//                                                      :209-220 (data_of_accessor)            gltf 1.2.0
// tobj and gltf are crates.io dependencies that are not part of the voidin checkout (Cargo.lock pins them); what is
// restated here is their published behaviour for exactly the options voidin passes.  Host code only, no CUDA.
#include <ctype.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/bvh_cuda.h"
#include "../../include/bvh_cuda_models.h"

namespace {

thread_local std::string g_model_err;

struct Mesh {
    std::vector<float> positions, normals, texcoords, tangents;
    std::vector<uint32_t> indices;
    int32_t material = -1, gltf_mesh = -1, gltf_primitive = -1;
    std::string name;
};
struct Material {
    std::string name;
    float base_color[4] = {1.f, 1.f, 1.f, 1.f};
};
struct Inst {
    float m[16];
    uint32_t mesh;
    int32_t material;
};

}  // namespace

struct bvh_cuda_model {
    std::vector<Mesh> meshes;
    std::vector<Material> materials;
    std::vector<Inst> instances;
};

namespace {

bool read_file(const std::string& path, std::string& out) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    if (n < 0) { fclose(f); return false; }
    out.resize((size_t)n);
    size_t got = n ? fread(&out[0], 1, (size_t)n, f) : 0;
    fclose(f);
    return got == (size_t)n;
}
std::string dir_of(const std::string& path) {
    size_t p = path.find_last_of("/\\");
    return p == std::string::npos ? std::string() : path.substr(0, p + 1);
}
int fail(const std::string& msg) {
    g_model_err = msg;
    return BVH_CUDA_EINVAL;
}

// =====================================================================================================================
// Wavefront OBJ with tobj's GPU_LOAD_OPTIONS = { triangulate, single_index, ignore_points, ignore_lines }
// (crates/app/src/models/mod.rs:24).  tobj::load_obj_buf: `v` / `vt` / `vn` pools are global to the file; a model is
// closed by the next `o` / `g` (and by a `usemtl` that changes the material) when it has faces; f32 fields are parsed
// with str::parse::<f32>() (correctly rounded, = strtof); indices are 1-based, negative = relative to the pool size at
// the time the face is read.  export_faces (single index): every distinct (v, vt, vn) triple of a model becomes one
// output vertex, numbered in the order of first use; polygons become the fan (0, i-1, i).
// =====================================================================================================================
struct VIdx {
    size_t v, vt, vn;
    bool operator==(const VIdx& o) const { return v == o.v && vt == o.vt && vn == o.vn; }
};
struct VIdxHash {
    size_t operator()(const VIdx& k) const {
        uint64_t h = k.v * 0x9E3779B97F4A7C15ull;
        h ^= (k.vt + 0x7F4A7C15ull) * 0xC2B2AE3D27D4EB4Full + (h << 6) + (h >> 2);
        h ^= (k.vn + 0x165667B1ull) * 0x27D4EB2F165667C5ull + (h << 6) + (h >> 2);
        return (size_t)h;
    }
};
constexpr size_t MISSING = (size_t)-1;

struct Tok {  // whitespace tokenizer over one line
    const char* p;
    const char* e;
    bool next(const char*& b, const char*& en) {
        while (p < e && (*p == ' ' || *p == '\t' || *p == '\r')) ++p;
        if (p >= e) return false;
        b = p;
        while (p < e && *p != ' ' && *p != '\t' && *p != '\r') ++p;
        en = p;
        return true;
    }
};
bool parse_f32(const char* b, const char* e, float& out) {
    char buf[64];
    size_t n = (size_t)(e - b);
    if (n == 0 || n >= sizeof buf) return false;
    memcpy(buf, b, n);
    buf[n] = 0;
    char* end = nullptr;
    out = strtof(buf, &end);
    return end == buf + n;
}
// one of the up to three '/'-separated fields of a face vertex
bool parse_index(const char* b, const char* e, size_t pool, size_t& out) {
    if (b == e) { out = MISSING; return true; }
    char buf[32];
    size_t n = (size_t)(e - b);
    if (n >= sizeof buf) return false;
    memcpy(buf, b, n);
    buf[n] = 0;
    char* end = nullptr;
    long long x = strtoll(buf, &end, 10);
    if (end != buf + n) return false;
    out = x < 0 ? (size_t)((long long)pool + x) : (size_t)(x - 1);
    return true;
}

struct ObjState {
    std::vector<float> pos, tex, nrm;
    std::vector<std::vector<VIdx>> faces;  // of the model being collected
    std::string name = "unnamed_object";   // tobj's default model name
    int32_t mat = -1;
};

int obj_add_vertex(const ObjState& st, Mesh& m, std::unordered_map<VIdx, uint32_t, VIdxHash>& map, const VIdx& vi) {
    auto it = map.find(vi);
    if (it != map.end()) { m.indices.push_back(it->second); return 0; }
    if (vi.v == MISSING || vi.v * 3 + 2 >= st.pos.size()) return fail("obj: face references a position that does not exist");
    m.positions.insert(m.positions.end(), st.pos.begin() + vi.v * 3, st.pos.begin() + vi.v * 3 + 3);
    if (!st.tex.empty() && vi.vt != MISSING) {
        if (vi.vt * 2 + 1 >= st.tex.size()) return fail("obj: face references a texcoord that does not exist");
        m.texcoords.push_back(st.tex[vi.vt * 2]);
        m.texcoords.push_back(st.tex[vi.vt * 2 + 1]);
    }
    if (!st.nrm.empty() && vi.vn != MISSING) {
        if (vi.vn * 3 + 2 >= st.nrm.size()) return fail("obj: face references a normal that does not exist");
        m.normals.insert(m.normals.end(), st.nrm.begin() + vi.vn * 3, st.nrm.begin() + vi.vn * 3 + 3);
    }
    const uint32_t next = (uint32_t)map.size();
    m.indices.push_back(next);
    map.emplace(vi, next);
    return 0;
}

int obj_export(ObjState& st, bvh_cuda_model& model) {
    Mesh m;
    m.name = st.name;
    m.material = st.mat;
    std::unordered_map<VIdx, uint32_t, VIdxHash> map;
    size_t corners = 0;
    for (auto& f : st.faces) corners += f.size();
    map.reserve(corners);
    for (auto& f : st.faces) {
        if (f.size() < 3) continue;  // points and lines: ignore_points / ignore_lines
        const VIdx a = f[0];
        VIdx b = f[1];
        for (size_t k = 2; k < f.size(); ++k) {
            const VIdx c = f[k];
            int rc;
            if ((rc = obj_add_vertex(st, m, map, a)) || (rc = obj_add_vertex(st, m, map, b)) || (rc = obj_add_vertex(st, m, map, c)))
                return rc;
            b = c;
        }
    }
    st.faces.clear();
    // ObjModel::import (models/mod.rs:44): tangents = Vec4::ZERO per position
    m.tangents.assign(m.positions.size() / 3 * 4, 0.f);
    model.meshes.push_back(std::move(m));
    return 0;
}

// tobj::load_mtl_buf, reduced to what ObjModel::import reads (models/mod.rs:28-36): the order of `newmtl` blocks and Kd
void load_mtl(const std::string& path, bvh_cuda_model& model, std::map<std::string, int32_t>& by_name) {
    std::string text;
    if (!read_file(path, text)) return;  // a missing .mtl is not an error for ObjModel::import (`if let Ok(..)`)
    const char* p = text.data();
    const char* end = p + text.size();
    while (p < end) {
        const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
        const char* le = nl ? nl : end;
        Tok t{p, le};
        const char *b, *e;
        if (t.next(b, e)) {
            const std::string key(b, e);
            if (key == "newmtl") {
                const char* nb = t.p;
                while (nb < le && (*nb == ' ' || *nb == '\t')) ++nb;
                const char* ne = le;
                while (ne > nb && (ne[-1] == ' ' || ne[-1] == '\t' || ne[-1] == '\r')) --ne;
                Material m;
                m.name.assign(nb, ne);
                m.base_color[3] = 0.5f;  // base_color.extend(0.5), models/mod.rs:32
                by_name[m.name] = (int32_t)model.materials.size();
                model.materials.push_back(m);
            } else if (key == "Kd" && !model.materials.empty()) {
                float c[3];
                int k = 0;
                while (k < 3 && t.next(b, e) && parse_f32(b, e, c[k])) ++k;
                if (k == 3) memcpy(model.materials.back().base_color, c, sizeof c);
            }
        }
        p = nl ? nl + 1 : end;
    }
}

int load_obj(const std::string& path, bvh_cuda_model& model) {
    std::string text;
    if (!read_file(path, text)) return fail("obj: cannot read " + path);
    ObjState st;
    std::map<std::string, int32_t> mat_by_name;
    const char* p = text.data();
    const char* end = p + text.size();
    size_t line_no = 0;
    while (p < end) {
        ++line_no;
        const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
        const char* le = nl ? nl : end;
        Tok t{p, le};
        const char *b, *e;
        if (t.next(b, e) && *b != '#') {
            const size_t kl = (size_t)(e - b);
            auto bad = [&](const char* what) { return fail("obj: " + std::string(what) + " at line " + std::to_string(line_no)); };
            if (kl == 1 && *b == 'v') {
                float x[3];
                for (int k = 0; k < 3; ++k)
                    if (!t.next(b, e) || !parse_f32(b, e, x[k])) return bad("bad position");
                st.pos.insert(st.pos.end(), x, x + 3);
            } else if (kl == 2 && b[0] == 'v' && b[1] == 't') {
                float x[2] = {0.f, 0.f};
                for (int k = 0; k < 2; ++k)
                    if (!t.next(b, e) || !parse_f32(b, e, x[k])) return bad("bad texcoord");
                st.tex.insert(st.tex.end(), x, x + 2);
            } else if (kl == 2 && b[0] == 'v' && b[1] == 'n') {
                float x[3];
                for (int k = 0; k < 3; ++k)
                    if (!t.next(b, e) || !parse_f32(b, e, x[k])) return bad("bad normal");
                st.nrm.insert(st.nrm.end(), x, x + 3);
            } else if (kl == 1 && (*b == 'f' || *b == 'l' || *b == 'p')) {
                std::vector<VIdx> face;
                while (t.next(b, e)) {
                    VIdx vi{MISSING, MISSING, MISSING};
                    const char* s1 = (const char*)memchr(b, '/', (size_t)(e - b));
                    const char* s2 = s1 ? (const char*)memchr(s1 + 1, '/', (size_t)(e - s1 - 1)) : nullptr;
                    bool ok = parse_index(b, s1 ? s1 : e, st.pos.size() / 3, vi.v);
                    if (s1) ok = ok && parse_index(s1 + 1, s2 ? s2 : e, st.tex.size() / 2, vi.vt);
                    if (s2) ok = ok && parse_index(s2 + 1, e, st.nrm.size() / 3, vi.vn);
                    if (!ok || vi.v == MISSING) return bad("bad face vertex");
                    face.push_back(vi);
                }
                if (face.empty()) return bad("empty face");
                st.faces.push_back(std::move(face));
            } else if (kl == 1 && (*b == 'o' || *b == 'g')) {
                if (!st.faces.empty()) {
                    if (int rc = obj_export(st, model)) return rc;
                }
                const char* nb = t.p;
                while (nb < le && (*nb == ' ' || *nb == '\t')) ++nb;
                const char* ne = le;
                while (ne > nb && (ne[-1] == ' ' || ne[-1] == '\t' || ne[-1] == '\r')) --ne;
                st.name.assign(nb, ne);
                if (st.name.empty()) st.name = "unnamed_object";
            } else if (kl == 6 && !memcmp(b, "mtllib", 6)) {
                if (t.next(b, e)) load_mtl(dir_of(path) + std::string(b, e), model, mat_by_name);
            } else if (kl == 6 && !memcmp(b, "usemtl", 6)) {
                const char* nb = t.p;
                while (nb < le && (*nb == ' ' || *nb == '\t')) ++nb;
                const char* ne = le;
                while (ne > nb && (ne[-1] == ' ' || ne[-1] == '\t' || ne[-1] == '\r')) --ne;
                auto it = mat_by_name.find(std::string(nb, ne));
                const int32_t new_mat = it == mat_by_name.end() ? -1 : it->second;
                // a material change in the middle of a model closes it; the next one keeps the name
                if (new_mat != st.mat && !st.faces.empty()) {
                    if (int rc = obj_export(st, model)) return rc;
                }
                st.mat = new_mat;
            }
            // everything else (s, vp, comments, ...) is skipped, as tobj does
        }
        p = nl ? nl + 1 : end;
    }
    if (!st.faces.empty()) {
        if (int rc = obj_export(st, model)) return rc;
    }
    return 0;
}

// =====================================================================================================================
// Minimal JSON (RFC 8259) DOM: what a glTF 2.0 document needs.  Numbers are kept as double; f32 fields are produced by
// `(float)double`, which is what serde_json does for an f32 target (parse as f64, then `as f32`).
// =====================================================================================================================
struct J {
    enum T { Null, Bool, Num, Str, Arr, Obj } t = Null;
    bool b = false;
    double n = 0;
    std::string s;
    std::vector<J> a;
    std::vector<std::pair<std::string, J>> o;
    const J* get(const char* k) const {
        if (t != Obj) return nullptr;
        for (auto& kv : o)
            if (kv.first == k) return &kv.second;
        return nullptr;
    }
    size_t size() const { return t == Arr ? a.size() : 0; }
};
struct JParser {
    const char* p;
    const char* e;
    bool ok = true;
    void ws() { while (p < e && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) ++p; }
    static void utf8(std::string& s, uint32_t c) {
        if (c < 0x80) s += (char)c;
        else if (c < 0x800) { s += (char)(0xC0 | (c >> 6)); s += (char)(0x80 | (c & 63)); }
        else if (c < 0x10000) { s += (char)(0xE0 | (c >> 12)); s += (char)(0x80 | ((c >> 6) & 63)); s += (char)(0x80 | (c & 63)); }
        else { s += (char)(0xF0 | (c >> 18)); s += (char)(0x80 | ((c >> 12) & 63)); s += (char)(0x80 | ((c >> 6) & 63)); s += (char)(0x80 | (c & 63)); }
    }
    bool hex4(uint32_t& v) {
        if (e - p < 4) return false;
        v = 0;
        for (int k = 0; k < 4; ++k) {
            char c = *p++;
            v <<= 4;
            if (c >= '0' && c <= '9') v |= (uint32_t)(c - '0');
            else if (c >= 'a' && c <= 'f') v |= (uint32_t)(c - 'a' + 10);
            else if (c >= 'A' && c <= 'F') v |= (uint32_t)(c - 'A' + 10);
            else return false;
        }
        return true;
    }
    bool str(std::string& out) {
        if (p >= e || *p != '"') return false;
        ++p;
        while (p < e && *p != '"') {
            if (*p == '\\') {
                if (++p >= e) return false;
                char c = *p++;
                switch (c) {
                    case 'n': out += '\n'; break;
                    case 't': out += '\t'; break;
                    case 'r': out += '\r'; break;
                    case 'b': out += '\b'; break;
                    case 'f': out += '\f'; break;
                    case 'u': {
                        uint32_t v, lo;
                        if (!hex4(v)) return false;
                        if (v >= 0xD800 && v < 0xDC00 && e - p >= 6 && p[0] == '\\' && p[1] == 'u') {
                            p += 2;
                            if (!hex4(lo)) return false;
                            v = 0x10000 + ((v - 0xD800) << 10) + (lo - 0xDC00);
                        }
                        utf8(out, v);
                        break;
                    }
                    default: out += c;  // \" \\ \/
                }
            } else out += *p++;
        }
        if (p >= e) return false;
        ++p;
        return true;
    }
    J value(int depth = 0) {
        J v;
        ws();
        if (p >= e || depth > 256) { ok = false; return v; }
        if (*p == '{') {
            v.t = J::Obj;
            ++p;
            ws();
            if (p < e && *p == '}') { ++p; return v; }
            while (ok) {
                ws();
                std::string k;
                if (!str(k)) { ok = false; break; }
                ws();
                if (p >= e || *p != ':') { ok = false; break; }
                ++p;
                v.o.emplace_back(std::move(k), value(depth + 1));
                ws();
                if (p < e && *p == ',') { ++p; continue; }
                if (p < e && *p == '}') { ++p; break; }
                ok = false;
            }
        } else if (*p == '[') {
            v.t = J::Arr;
            ++p;
            ws();
            if (p < e && *p == ']') { ++p; return v; }
            while (ok) {
                v.a.push_back(value(depth + 1));
                ws();
                if (p < e && *p == ',') { ++p; continue; }
                if (p < e && *p == ']') { ++p; break; }
                ok = false;
            }
        } else if (*p == '"') {
            v.t = J::Str;
            if (!str(v.s)) ok = false;
        } else if (e - p >= 4 && !memcmp(p, "true", 4)) { v.t = J::Bool; v.b = true; p += 4; }
        else if (e - p >= 5 && !memcmp(p, "false", 5)) { v.t = J::Bool; v.b = false; p += 5; }
        else if (e - p >= 4 && !memcmp(p, "null", 4)) { p += 4; }
        else {
            const char* q = p;
            while (q < e && (*q == '-' || *q == '+' || *q == '.' || *q == 'e' || *q == 'E' || (*q >= '0' && *q <= '9'))) ++q;
            if (q == p || q - p > 63) { ok = false; return v; }
            char buf[64];
            memcpy(buf, p, (size_t)(q - p));
            buf[q - p] = 0;
            char* end = nullptr;
            v.n = strtod(buf, &end);
            if (end != buf + (q - p)) ok = false;
            v.t = J::Num;
            p = q;
        }
        return v;
    }
};
long long jint(const J* j, long long dflt) { return j && j->t == J::Num ? (long long)j->n : dflt; }

bool base64_decode(const char* b, const char* e, std::string& out) {
    static int8_t tbl[256];
    static bool init = false;
    if (!init) {
        memset(tbl, -1, sizeof tbl);
        const char* abc = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
        for (int i = 0; i < 64; ++i) tbl[(uint8_t)abc[i]] = (int8_t)i;
        init = true;
    }
    uint32_t acc = 0;
    int bits = 0;
    for (; b < e; ++b) {
        if (*b == '=' || *b == '\n' || *b == '\r') continue;
        int8_t v = tbl[(uint8_t)*b];
        if (v < 0) return false;
        acc = (acc << 6) | (uint32_t)v;
        bits += 6;
        if (bits >= 8) { bits -= 8; out += (char)((acc >> bits) & 0xFF); }
    }
    return true;
}

// =====================================================================================================================
// glTF 2.0
// =====================================================================================================================
struct Gltf {
    J doc;
    std::vector<std::string> buffers;
};
struct Acc {  // a resolved accessor
    const uint8_t* base = nullptr;  // first element (bufferView.byteOffset + accessor.byteOffset)
    size_t avail = 0;               // bytes from base to the end of the bufferView
    size_t count = 0, stride = 0, comp_size = 0, ncomp = 0;
    int comp_type = 0;
    bool normalized = false;
};
size_t comp_size_of(int ct) {
    switch (ct) {
        case 5120: case 5121: return 1;
        case 5122: case 5123: return 2;
        case 5125: case 5126: return 4;
    }
    return 0;
}
size_t ncomp_of(const std::string& t) {
    if (t == "SCALAR") return 1;
    if (t == "VEC2") return 2;
    if (t == "VEC3") return 3;
    if (t == "VEC4" || t == "MAT2") return 4;
    if (t == "MAT3") return 9;
    if (t == "MAT4") return 16;
    return 0;
}
// false = no such accessor / no bufferView (`accessor.view()?` in data_of_accessor, gltf_model/mod.rs:213) / out of range
bool resolve(const Gltf& g, long long idx, Acc& out) {
    const J* accs = g.doc.get("accessors");
    if (!accs || idx < 0 || (size_t)idx >= accs->size()) return false;
    const J& a = accs->a[(size_t)idx];
    const long long bvi = jint(a.get("bufferView"), -1);
    const J* views = g.doc.get("bufferViews");
    if (bvi < 0 || !views || (size_t)bvi >= views->size()) return false;
    const J& bv = views->a[(size_t)bvi];
    const long long bi = jint(bv.get("buffer"), -1);
    if (bi < 0 || (size_t)bi >= g.buffers.size()) return false;
    const std::string& buf = g.buffers[(size_t)bi];
    const size_t voff = (size_t)jint(bv.get("byteOffset"), 0), vlen = (size_t)jint(bv.get("byteLength"), 0);
    const size_t aoff = (size_t)jint(a.get("byteOffset"), 0);
    if (voff > buf.size() || vlen > buf.size() - voff || aoff > vlen) return false;
    out.comp_type = (int)jint(a.get("componentType"), 0);
    out.comp_size = comp_size_of(out.comp_type);
    const J* ty = a.get("type");
    out.ncomp = ty && ty->t == J::Str ? ncomp_of(ty->s) : 0;
    if (!out.comp_size || !out.ncomp) return false;
    out.count = (size_t)jint(a.get("count"), 0);
    const size_t bs = (size_t)jint(bv.get("byteStride"), 0);
    out.stride = bs ? bs : out.comp_size * out.ncomp;
    const J* nm = a.get("normalized");
    out.normalized = nm && nm->t == J::Bool && nm->b;
    out.base = (const uint8_t*)buf.data() + voff + aoff;
    out.avail = vlen - aoff;
    return true;
}
// data_of_accessor (gltf_model/mod.rs:209-220): `count * size` CONTIGUOUS bytes from the accessor's start — the
// reference ignores byteStride for POSITION and NORMAL, and so does this.
bool raw_vec3(const Acc& a, std::vector<float>& out) {
    const size_t bytes = a.count * a.comp_size * a.ncomp;
    if (bytes > a.avail || bytes % 12) return false;  // the slice index / bytemuck::cast_slice would panic
    out.resize(bytes / 4);
    if (bytes) memcpy(out.data(), a.base, bytes);
    return true;
}
bool elem_ok(const Acc& a) { return a.count == 0 || (a.count - 1) * a.stride + a.comp_size * a.ncomp <= a.avail; }
// gltf::mesh::Reader::read_tex_coords(0).into_f32(): f32 as is, u8 / 255, u16 / 65535
bool read_texcoords(const Acc& a, std::vector<float>& out) {
    if (a.ncomp != 2 || !elem_ok(a)) return false;
    out.resize(a.count * 2);
    for (size_t i = 0; i < a.count; ++i)
        for (size_t k = 0; k < 2; ++k) {
            const uint8_t* p = a.base + i * a.stride + k * a.comp_size;
            float v;
            if (a.comp_type == 5126) memcpy(&v, p, 4);
            else if (a.comp_type == 5121) v = (float)p[0] / 255.0f;
            else if (a.comp_type == 5123) { uint16_t u; memcpy(&u, p, 2); v = (float)u / 65535.0f; }
            else return false;
            out[i * 2 + k] = v;
        }
    return true;
}
bool read_tangents(const Acc& a, std::vector<float>& out) {
    if (a.ncomp != 4 || a.comp_type != 5126 || !elem_ok(a)) return false;
    out.resize(a.count * 4);
    for (size_t i = 0; i < a.count; ++i) memcpy(&out[i * 4], a.base + i * a.stride, 16);
    return true;
}
// read_indices().into_u32(): u8 / u16 / u32 widened
bool read_indices(const Acc& a, std::vector<uint32_t>& out) {
    if (a.ncomp != 1 || !elem_ok(a)) return false;
    out.resize(a.count);
    for (size_t i = 0; i < a.count; ++i) {
        const uint8_t* p = a.base + i * a.stride;
        if (a.comp_type == 5121) out[i] = p[0];
        else if (a.comp_type == 5123) { uint16_t u; memcpy(&u, p, 2); out[i] = u; }
        else if (a.comp_type == 5125) memcpy(&out[i], p, 4);
        else return false;
    }
    return true;
}

// ---- transforms (column-major float[16]) ------------------------------------------------------------------------------
// Every product is ((a*x + b*y) + c*z) + d*w per column, in f32, unfused: the order of both gltf's own Matrix4 (used by
// Node::transform().matrix() for T*R*S nodes) and glam 0.24's Mat4 * Mat4 (gather_instances_recursive, :173-174).
#if defined(__GNUC__)
#pragma GCC push_options
#pragma GCC optimize("fp-contract=off")
#endif
void mat_mul(const float* A, const float* B, float* out) {
    float r[16];
    for (int c = 0; c < 4; ++c)
        for (int k = 0; k < 4; ++k) {
            float acc = A[0 * 4 + k] * B[c * 4 + 0];
            acc = acc + A[1 * 4 + k] * B[c * 4 + 1];
            acc = acc + A[2 * 4 + k] * B[c * 4 + 2];
            acc = acc + A[3 * 4 + k] * B[c * 4 + 3];
            r[c * 4 + k] = acc;
        }
    memcpy(out, r, sizeof r);
}
// gltf::scene::Transform::Decomposed -> matrix(): T * R * S with the cgmath quaternion formula
void trs_matrix(const float* t, const float* q /* x y z w */, const float* s, float* out) {
    const float x = q[0], y = q[1], z = q[2], w = q[3];
    const float x2 = x + x, y2 = y + y, z2 = z + z;
    const float xx2 = x2 * x, xy2 = x2 * y, xz2 = x2 * z, yy2 = y2 * y, yz2 = y2 * z, zz2 = z2 * z;
    const float sy2 = y2 * w, sz2 = z2 * w, sx2 = x2 * w;
    const float R[16] = {1.f - yy2 - zz2, xy2 + sz2, xz2 - sy2, 0.f, xy2 - sz2, 1.f - xx2 - zz2, yz2 + sx2, 0.f,
                         xz2 + sy2, yz2 - sx2, 1.f - xx2 - yy2, 0.f, 0.f, 0.f, 0.f, 1.f};
    const float T[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, t[0], t[1], t[2], 1};
    const float S[16] = {s[0], 0, 0, 0, 0, s[1], 0, 0, 0, 0, s[2], 0, 0, 0, 0, 1};
    float TR[16];
    mat_mul(T, R, TR);
    mat_mul(TR, S, out);
}
#if defined(__GNUC__)
#pragma GCC pop_options
#endif
bool f32_array(const J* j, size_t n, float* out) {
    if (!j || j->t != J::Arr || j->a.size() != n) return false;
    for (size_t i = 0; i < n; ++i) {
        if (j->a[i].t != J::Num) return false;
        out[i] = (float)j->a[i].n;
    }
    return true;
}
void node_matrix(const J& node, float* out) {
    if (f32_array(node.get("matrix"), 16, out)) return;
    float t[3] = {0, 0, 0}, r[4] = {0, 0, 0, 1}, s[3] = {1, 1, 1};
    f32_array(node.get("translation"), 3, t);
    f32_array(node.get("rotation"), 4, r);
    f32_array(node.get("scale"), 3, s);
    trs_matrix(t, r, s, out);
}

// gather_instances_recursive (gltf_model/mod.rs:166-207): children first, then the node's own primitives
void gather(const Gltf& g, const std::map<std::pair<long long, long long>, uint32_t>& mesh_of, size_t node_idx, const float* parent,
            std::vector<Inst>& out, int depth) {
    const J* nodes = g.doc.get("nodes");
    if (!nodes || node_idx >= nodes->size() || depth > 512) return;
    const J& node = nodes->a[node_idx];
    float local[16], world[16];
    node_matrix(node, local);
    mat_mul(parent, local, world);
    if (const J* ch = node.get("children"))
        for (auto& c : ch->a) gather(g, mesh_of, (size_t)jint(&c, -1), world, out, depth + 1);
    const long long mi = jint(node.get("mesh"), -1);
    const J* meshes = g.doc.get("meshes");
    if (mi < 0 || !meshes || (size_t)mi >= meshes->size()) return;
    const J* prims = meshes->a[(size_t)mi].get("primitives");
    if (!prims) return;
    for (size_t pi = 0; pi < prims->size(); ++pi) {
        auto it = mesh_of.find({mi, (long long)pi});
        if (it == mesh_of.end()) continue;
        Inst in;
        memcpy(in.m, world, sizeof world);
        in.mesh = it->second;
        in.material = (int32_t)jint(prims->a[pi].get("material"), -1);
        out.push_back(in);
    }
}

int load_gltf(const std::string& path, bvh_cuda_model& model) {
    std::string raw;
    if (!read_file(path, raw)) return fail("gltf: cannot read " + path);
    Gltf g;
    std::string json_text, glb_bin;
    bool has_glb_bin = false;
    if (raw.size() >= 12 && !memcmp(raw.data(), "glTF", 4)) {
        uint32_t length;
        memcpy(&length, raw.data() + 8, 4);
        size_t off = 12, lim = length < raw.size() ? length : raw.size();
        while (off + 8 <= lim) {
            uint32_t clen, ctype;
            memcpy(&clen, raw.data() + off, 4);
            memcpy(&ctype, raw.data() + off + 4, 4);
            if (off + 8 + (size_t)clen > lim) return fail("glb: truncated chunk");
            if (ctype == 0x4E4F534Au) json_text.assign(raw, off + 8, clen);
            else if (ctype == 0x004E4942u && !has_glb_bin) { glb_bin.assign(raw, off + 8, clen); has_glb_bin = true; }
            off += 8 + (size_t)clen;
        }
        if (json_text.empty()) return fail("glb: no JSON chunk");
    } else {
        json_text.swap(raw);
    }
    JParser jp{json_text.data(), json_text.data() + json_text.size()};
    g.doc = jp.value();
    if (!jp.ok || g.doc.t != J::Obj) return fail("gltf: malformed JSON in " + path);

    if (const J* bufs = g.doc.get("buffers"))
        for (size_t i = 0; i < bufs->size(); ++i) {
            const J* uri = bufs->a[i].get("uri");
            std::string data;
            if (!uri || uri->t != J::Str) {
                if (i != 0 || !has_glb_bin) return fail("gltf: buffer " + std::to_string(i) + " has no uri and there is no GLB chunk");
                data = glb_bin;
            } else if (uri->s.compare(0, 5, "data:") == 0) {
                size_t comma = uri->s.find(',');
                if (comma == std::string::npos || !base64_decode(uri->s.data() + comma + 1, uri->s.data() + uri->s.size(), data))
                    return fail("gltf: bad data uri in buffer " + std::to_string(i));
            } else {
                std::string rel;  // percent-decoding, as gltf::import does for file uris
                for (size_t k = 0; k < uri->s.size(); ++k) {
                    if (uri->s[k] == '%' && k + 2 < uri->s.size() && isxdigit((unsigned char)uri->s[k + 1]) && isxdigit((unsigned char)uri->s[k + 2])) {
                        rel += (char)strtol(uri->s.substr(k + 1, 2).c_str(), nullptr, 16);
                        k += 2;
                    } else rel += uri->s[k];
                }
                if (!read_file(dir_of(path) + rel, data)) return fail("gltf: cannot read buffer " + dir_of(path) + rel);
            }
            const size_t want = (size_t)jint(bufs->a[i].get("byteLength"), 0);
            if (data.size() < want) return fail("gltf: buffer " + std::to_string(i) + " is shorter than its byteLength");
            g.buffers.push_back(std::move(data));
        }

    // materials: only what an instance refers to (index) and the base colour are kept
    if (const J* mats = g.doc.get("materials"))
        for (auto& m : mats->a) {
            Material mm;
            if (const J* n = m.get("name")) mm.name = n->s;
            if (const J* pbr = m.get("pbrMetallicRoughness")) f32_array(pbr->get("baseColorFactor"), 4, mm.base_color);
            model.materials.push_back(mm);
        }

    // make_meshes (gltf_model/mod.rs:103-155)
    std::map<std::pair<long long, long long>, uint32_t> mesh_of;
    if (const J* meshes = g.doc.get("meshes"))
        for (size_t mi = 0; mi < meshes->size(); ++mi) {
            const J* prims = meshes->a[mi].get("primitives");
            if (!prims) continue;
            for (size_t pi = 0; pi < prims->size(); ++pi) {
                const J& prim = prims->a[pi];
                const J* attrs = prim.get("attributes");
                if (!attrs) continue;
                Acc apos, anrm, a;
                if (!resolve(g, jint(attrs->get("POSITION"), -1), apos)) continue;  // `let Some(..) else { continue }` :117
                if (!resolve(g, jint(attrs->get("NORMAL"), -1), anrm)) continue;    // :120
                Mesh m;
                if (!raw_vec3(apos, m.positions) || !raw_vec3(anrm, m.normals))
                    return fail("gltf: POSITION / NORMAL of mesh " + std::to_string(mi) + " do not fit their buffer view");
                const size_t nv = m.positions.size() / 3;
                // tangents: read_tangents() padded with [0,1,0,1] to the vertex count (:125-131)
                std::vector<float> tan;
                if (resolve(g, jint(attrs->get("TANGENT"), -1), a)) read_tangents(a, tan);
                tan.resize(nv * 4 < tan.size() ? nv * 4 : tan.size());
                for (size_t i = tan.size() / 4; i < nv; ++i) { const float d[4] = {0.f, 1.f, 0.f, 1.f}; tan.insert(tan.end(), d, d + 4); }
                m.tangents.swap(tan);
                // tex_coords: set 0 as f32, `unwrap_repeat` pads with the default, truncated to the vertex count (:132-137)
                std::vector<float> uv;
                if (resolve(g, jint(attrs->get("TEXCOORD_0"), -1), a)) read_texcoords(a, uv);
                uv.resize(nv * 2, 0.f);
                m.texcoords.swap(uv);
                // indices: widened to u32, or 0..n when absent (:138-141)
                if (resolve(g, jint(prim.get("indices"), -1), a)) {
                    if (!read_indices(a, m.indices)) return fail("gltf: unreadable indices in mesh " + std::to_string(mi));
                } else {
                    m.indices.resize(nv);
                    for (size_t i = 0; i < nv; ++i) m.indices[i] = (uint32_t)i;
                }
                if (const J* n = meshes->a[mi].get("name")) m.name = n->s;
                m.material = (int32_t)jint(prim.get("material"), -1);
                m.gltf_mesh = (int32_t)mi;
                m.gltf_primitive = (int32_t)pi;
                mesh_of[{(long long)mi, (long long)pi}] = (uint32_t)model.meshes.size();
                model.meshes.push_back(std::move(m));
            }
        }

    // get_scene_instances(Mat4::IDENTITY) (:157-163)
    const float ident[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    if (const J* scenes = g.doc.get("scenes"))
        for (auto& sc : scenes->a)
            if (const J* roots = sc.get("nodes"))
                for (auto& r : roots->a) gather(g, mesh_of, (size_t)jint(&r, -1), ident, model.instances, 0);
    return 0;
}

template <class F>
int guarded_load(const char* path, bvh_cuda_model** out, F&& f) {
    if (!path || !out) return fail("null argument");
    *out = nullptr;
    try {
        std::unique_ptr<bvh_cuda_model> m(new bvh_cuda_model);
        if (int rc = f(std::string(path), *m)) return rc;
        g_model_err.clear();
        *out = m.release();
        return BVH_CUDA_OK;
    } catch (const std::bad_alloc&) {
        g_model_err = "out of host memory";
        return BVH_CUDA_ENOMEM;
    } catch (...) {
        return fail("unexpected failure while loading");
    }
}

}  // namespace

extern "C" {

int bvh_cuda_model_load_obj(const char* path, bvh_cuda_model** out) { return guarded_load(path, out, load_obj); }
int bvh_cuda_model_load_gltf(const char* path, bvh_cuda_model** out) { return guarded_load(path, out, load_gltf); }
void bvh_cuda_model_free(bvh_cuda_model* model) { delete model; }
const char* bvh_cuda_model_last_error(void) { return g_model_err.c_str(); }

size_t bvh_cuda_model_mesh_count(const bvh_cuda_model* model) { return model ? model->meshes.size() : 0; }
int bvh_cuda_model_mesh(const bvh_cuda_model* model, size_t i, BvhCudaMeshView* out) {
    if (!model || !out || i >= model->meshes.size()) return BVH_CUDA_EINVAL;
    const Mesh& m = model->meshes[i];
    out->positions = m.positions.data();
    out->normals = m.normals.empty() ? nullptr : m.normals.data();
    out->texcoords = m.texcoords.empty() ? nullptr : m.texcoords.data();
    out->tangents = m.tangents.empty() ? nullptr : m.tangents.data();
    out->indices = m.indices.data();
    out->n_vertices = m.positions.size() / 3;
    out->n_normals = m.normals.size() / 3;
    out->n_texcoords = m.texcoords.size() / 2;
    out->n_indices = m.indices.size();
    out->material = m.material;
    out->gltf_mesh = m.gltf_mesh;
    out->gltf_primitive = m.gltf_primitive;
    out->reserved = 0;
    out->name = m.name.c_str();
    return BVH_CUDA_OK;
}
size_t bvh_cuda_model_material_count(const bvh_cuda_model* model) { return model ? model->materials.size() : 0; }
int bvh_cuda_model_material(const bvh_cuda_model* model, size_t i, BvhCudaMaterialView* out) {
    if (!model || !out || i >= model->materials.size()) return BVH_CUDA_EINVAL;
    memcpy(out->base_color, model->materials[i].base_color, sizeof out->base_color);
    out->name = model->materials[i].name.c_str();
    return BVH_CUDA_OK;
}
size_t bvh_cuda_model_instance_count(const bvh_cuda_model* model) { return model ? model->instances.size() : 0; }
int bvh_cuda_model_instance(const bvh_cuda_model* model, size_t i, BvhCudaInstanceView* out) {
    if (!model || !out || i >= model->instances.size()) return BVH_CUDA_EINVAL;
    memcpy(out->transform, model->instances[i].m, sizeof out->transform);
    out->mesh = model->instances[i].mesh;
    out->material = model->instances[i].material;
    return BVH_CUDA_OK;
}

}  // extern "C"
