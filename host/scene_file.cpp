// scene_file.cpp -- see scene_file.hpp.  Reference behaviour followed: sceneLoader.cpp:47-308 (LoadScene) and
// tiny_obj_loader.h (LoadObj 813-1035, exportFaceGroupToShape 478-528, updateVertex 416-452, parseTriple 382-414,
// tryParseDouble 213-344).
#include "scene_file.hpp"

#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <tuple>

namespace spchost {

std::string normalize_path(const std::string& p) {
    std::string r = p;
    for (char& c : r)
        if (c == '\\') c = '/';
    return r;
}

static std::string dir_of(const std::string& path) {
    const size_t k = path.find_last_of('/');
    return k == std::string::npos ? std::string(".") : (k == 0 ? std::string("/") : path.substr(0, k));
}

// ---------------------------------------------------------------------------------------------------------
// .scene key scanning.  The reference scans every line of a block with sscanf(" key %f ..."): leading blanks are
// skipped, the key must match as a prefix, conversions stop at the first token that is not a number (values
// converted so far are kept).  key_floats / key_token / key_int behave the same way.
// ---------------------------------------------------------------------------------------------------------
static const char* after_key(const char* line, const char* key) {
    while (*line && isspace((unsigned char)*line)) line++;
    const size_t n = strlen(key);
    if (strncmp(line, key, n) != 0) return nullptr;
    return line + n;
}

static int key_floats(const char* line, const char* key, int count, float* dst) {
    const char* p = after_key(line, key);
    if (!p) return 0;
    int got = 0;
    for (; got < count; got++) {
        char* e = nullptr;
        const float v = strtof(p, &e);
        if (e == p) break;
        dst[got] = v;
        p = e;
    }
    return got;
}

static bool key_int(const char* line, const char* key, int* dst) {
    const char* p = after_key(line, key);
    if (!p) return false;
    char* e = nullptr;
    const long v = strtol(p, &e, 10);   // %d and %i agree on decimal input, which is all the format uses
    if (e == p) return false;
    *dst = (int)v;
    return true;
}

static bool key_token(const char* line, const char* key, std::string* dst) {
    const char* p = after_key(line, key);
    if (!p) return false;
    while (*p && isspace((unsigned char)*p)) p++;
    if (!*p) return false;
    const char* e = p;
    while (*e && !isspace((unsigned char)*e)) e++;
    dst->assign(p, e);
    return true;
}

namespace {
struct LineReader {
    FILE* f;
    char  buf[2048];   // kMaxLineLength (sceneLoader.cpp:20): longer lines arrive in pieces, as with fgets there
    bool next() { return fgets(buf, sizeof(buf), f) != nullptr; }
    bool closes_block() const { return strchr(buf, '}') != nullptr; }
};

inline void sub3(const float* a, const float* b, float* r) { r[0] = a[0] - b[0]; r[1] = a[1] - b[1]; r[2] = a[2] - b[2]; }
inline void cross3(const float* a, const float* b, float* r) {   // sutil/vec_math.h cross()
    const float x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
    r[0] = x; r[1] = y; r[2] = z;
}
inline float len3(const float* a) { return sqrtf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }
inline void normalize3(float* a) {   // sutil/vec_math.h normalize(): v * (1 / sqrtf(dot(v, v)))
    const float inv = 1.0f / sqrtf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
    a[0] *= inv; a[1] *= inv; a[2] *= inv;
}
}  // namespace

bool load_scene_file(const std::string& filename_in, const std::string& data_root_in, SceneFile& out, std::string& err) {
    const std::string filename = normalize_path(filename_in);
    FILE* file = fopen(filename.c_str(), "r");
    if (!file) {
        err = "cannot open scene file " + filename;
        return false;
    }
    out = SceneFile();
    std::string root = normalize_path(data_root_in);
    if (root.empty()) root = dir_of(dir_of(filename));
    out.data_root = root;
    auto resolve = [&](const std::string& rel) { return root + "/" + normalize_path(rel); };

    std::map<std::string, MaterialParameter> materials_by_name;
    std::map<std::string, int> texture_ids;
    int n_textures = 0;
    LineReader in{file, {0}};

    while (in.next()) {
        if (in.buf[0] == '#') continue;   // only a '#' in column 0 comments a line out (sceneLoader.cpp:68)
        std::string name;

        // ---- material <name> { ... } ------------------------------------------------------------------
        if (key_token(in.buf, "material", &name)) {
            MaterialParameter m;
            std::string tex = "None";
            while (in.next()) {
                if (in.closes_block()) break;
                key_token(in.buf, "name", &name);
                key_floats(in.buf, "color", 3, m.color);
                key_token(in.buf, "albedoTex", &tex);
                key_floats(in.buf, "emission", 3, m.emission);
                key_floats(in.buf, "metallic", 1, &m.metallic);
                key_floats(in.buf, "subsurface", 1, &m.subsurface);
                key_floats(in.buf, "specular", 1, &m.specular);          // "specularTint x" fails the %f and leaves it alone
                key_floats(in.buf, "specularTint", 1, &m.specularTint);
                key_floats(in.buf, "roughness", 1, &m.roughness);
                key_floats(in.buf, "anisotropic", 1, &m.anisotropic);
                key_floats(in.buf, "sheen", 1, &m.sheen);
                key_floats(in.buf, "sheenTint", 1, &m.sheenTint);
                key_floats(in.buf, "clearcoat", 1, &m.clearcoat);
                key_floats(in.buf, "clearcoatGloss", 1, &m.clearcoatGloss);
                key_int(in.buf, "brdf", &m.brdf);
            }
            auto it = texture_ids.find(tex);
            if (it != texture_ids.end()) {
                m.albedoID = it->second;
            } else if (tex != "None") {
                n_textures++;
                texture_ids[tex] = n_textures;
                out.texture_map[n_textures - 1] = normalize_path(tex);
                m.albedoID = n_textures;
            }
            materials_by_name[name] = m;
        }

        // ---- light { ... }: any line that contains "light" opens one (sceneLoader.cpp:128) --------------
        if (strstr(in.buf, "light")) {
            LightParameter L;
            float v1[3] = {0, 0, 0}, v2[3] = {0, 0, 0};
            std::string type = "None";
            int div = 1;
            while (in.next()) {
                if (in.closes_block()) break;
                key_floats(in.buf, "position", 3, L.position);
                key_floats(in.buf, "emission", 3, L.emission);
                key_floats(in.buf, "normal", 3, L.normal);
                key_floats(in.buf, "direction", 3, L.direction);
                key_floats(in.buf, "radius", 1, &L.radius);
                key_floats(in.buf, "v1", 3, v1);
                key_floats(in.buf, "v2", 3, v2);
                key_token(in.buf, "type", &type);
                key_int(in.buf, "divLevel", &div);
            }
            L.divLevel = div;
            if (type == "Quad") {
                L.lightType = LK_QUAD;
                sub3(v1, L.position, L.u);
                sub3(v2, L.position, L.v);
                float n[3];
                cross3(L.u, L.v, n);
                L.area = len3(n);
                normalize3(n);
                memcpy(L.normal, n, sizeof(n));
            } else if (type == "Sphere") {
                L.lightType = LK_SPHERE;
                normalize3(L.normal);
                L.area = 4.0f * 3.14159265358979323846f * L.radius * L.radius;
            } else if (type == "Direction") {
                L.lightType = LK_DIRECTION;
                normalize3(L.direction);
            } else if (type == "Env") {
                L.lightType = LK_ENV;
            }
            out.lights.push_back(L);
        }

        // ---- properties { width, height } (the application ignores them: optixPathTracer.cpp:700-701) ------
        if (strstr(in.buf, "properties")) {
            while (in.next()) {
                if (in.closes_block()) break;
                key_int(in.buf, "width", &out.width);
                key_int(in.buf, "height", &out.height);
            }
        }

        // ---- cameraSetting { ... } ------------------------------------------------------------------------
        if (strstr(in.buf, "cameraSetting")) {
            out.use_camera = true;
            float eye[3] = {0, 0, 0}, lookat[3] = {0, 0, 0}, up[3] = {0, 1, 0}, fov = 35.0f, env_lum = 1.0f;
            int geo_normal = 0;
            std::string env;
            while (in.next()) {
                if (in.closes_block()) break;
                key_floats(in.buf, "eye", 3, eye);
                key_floats(in.buf, "lookat", 3, lookat);
                key_floats(in.buf, "up", 3, up);
                key_floats(in.buf, "fov", 1, &fov);
                key_int(in.buf, "geo_normal", &geo_normal);
                key_floats(in.buf, "env_lum", 1, &env_lum);
                key_token(in.buf, "env_file", &env);
            }
            out.use_geometry_normal = geo_normal == 1;
            memcpy(out.eye, eye, sizeof(eye));
            memcpy(out.lookat, lookat, sizeof(lookat));
            memcpy(out.up, up, sizeof(up));
            out.fov = fov;
            out.env_file = env;
            out.env_factor = env_lum;
        }

        // ---- mesh { file, uv_file, material, transform } ---------------------------------------------------
        if (strstr(in.buf, "mesh")) {
            bool has_material = false;
            const size_t meshes_before = out.mesh_names.size();
            while (in.next()) {
                if (in.closes_block()) break;
                std::string tok;
                if (key_token(in.buf, "file", &tok)) {
                    out.mesh_names.push_back(resolve(tok));
                    out.uv_mesh_names.push_back(resolve(tok));
                }
                if (key_token(in.buf, "uv_file", &tok) && !out.uv_mesh_names.empty()) out.uv_mesh_names.back() = resolve(tok);
                if (key_token(in.buf, "material", &tok)) {
                    auto it = materials_by_name.find(tok);
                    if (it != materials_by_name.end()) {
                        out.materials.push_back(it->second);
                    } else {
                        // The reference prints "Could not find material" and pushes nothing, which shifts the material of
                        // every later mesh by one.  Keep mesh k <-> material k instead and say so.
                        out.warnings.push_back("material '" + tok + "' not found: default material used");
                        out.materials.push_back(MaterialParameter());
                    }
                    has_material = true;
                }
                if (strstr(in.buf, "transform")) out.warnings.push_back("mesh transform ignored (the reference parses and drops it, sceneLoader.cpp:295-306)");
            }
            if (out.mesh_names.size() > meshes_before && !has_material) {
                out.warnings.push_back("mesh '" + out.mesh_names.back() + "' has no material line: default material used");
                out.materials.push_back(MaterialParameter());
            }
        }
    }
    fclose(file);
    return true;
}

// ---------------------------------------------------------------------------------------------------------
// OBJ
// ---------------------------------------------------------------------------------------------------------
float parse_obj_float(const char* s, const char* s_end) {
    if (s >= s_end) return 0.0f;
    double mantissa = 0.0;
    int    exponent = 0;
    bool   negative = false, exp_negative = false;
    const char* c = s;
    if (*c == '+' || *c == '-') {
        negative = *c == '-';
        c++;
    } else if (!isdigit((unsigned char)*c)) {
        return 0.0f;
    }
    int read = 0;
    for (; c != s_end && isdigit((unsigned char)*c); c++, read++) {
        mantissa *= 10;
        mantissa += (int)(*c - '0');
    }
    if (read == 0) return 0.0f;
    bool at_end = c == s_end;
    if (!at_end) {
        if (*c == '.') {
            c++;
            read = 1;
            for (; c != s_end && isdigit((unsigned char)*c); c++, read++) mantissa += (int)(*c - '0') * pow(10.0, -read);
            at_end = c == s_end;
        } else if (*c != 'e' && *c != 'E') {
            at_end = true;   // "goto assemble"
        }
    }
    if (!at_end && (*c == 'e' || *c == 'E')) {
        c++;
        if (c != s_end && (*c == '+' || *c == '-')) {
            exp_negative = *c == '-';
            c++;
        } else if (!(c != s_end && isdigit((unsigned char)*c))) {
            return 0.0f;   // empty exponent: the whole number is rejected
        }
        read = 0;
        for (; c != s_end && isdigit((unsigned char)*c); c++, read++) exponent = exponent * 10 + (int)(*c - '0');
        if (exp_negative) exponent = -exponent;
        if (read == 0) return 0.0f;
    }
    const double val = (negative ? -1 : 1) * ldexp(mantissa * pow(5.0, exponent), exponent);
    return (float)val;
}

namespace {
struct VIdx {
    int v, vt, vn;
    bool operator<(const VIdx& o) const { return std::tie(v, vn, vt) < std::tie(o.v, o.vn, o.vt); }
};

inline bool is_blank(char c) { return c == ' ' || c == '\t'; }
inline bool is_eol(char c) { return c == '\r' || c == '\n' || c == '\0'; }

float next_float(const char*& t) {
    t += strspn(t, " \t");
    const char* e = t + strcspn(t, " \t\r");
    const float f = parse_obj_float(t, e);
    t = e;
    return f;
}

// 1-based -> 0-based, 0 stays 0, negative = relative to the count so far (tiny_obj_loader.h fixIndex)
inline int fix_index(int idx, int n) { return idx > 0 ? idx - 1 : (idx == 0 ? 0 : n + idx); }

VIdx parse_corner(const char*& t, int nv, int nvn, int nvt) {
    VIdx r{-1, -1, -1};
    r.v = fix_index(atoi(t), nv);
    t += strcspn(t, "/ \t\r");
    if (*t != '/') return r;
    t++;
    if (*t == '/') {   // v//vn
        t++;
        r.vn = fix_index(atoi(t), nvn);
        t += strcspn(t, "/ \t\r");
        return r;
    }
    r.vt = fix_index(atoi(t), nvt);
    t += strcspn(t, "/ \t\r");
    if (*t != '/') return r;
    t++;
    r.vn = fix_index(atoi(t), nvn);
    t += strcspn(t, "/ \t\r");
    return r;
}

struct ObjState {
    std::vector<float> v, vn, vt;
    std::vector<std::vector<VIdx>> faces;
    std::string name;
};

// Flatten the pending faces into one shape.  Vertices are de-duplicated per shape on the (v, vt, vn) triple in
// first-use order; the reference passes its cache by value, so it never carries over between shapes.
bool flush_shape(ObjState& st, std::vector<ObjShape>& shapes) {
    if (st.faces.empty()) return false;
    ObjShape sh;
    std::map<VIdx, unsigned> cache;
    auto vertex = [&](const VIdx& k) -> unsigned {
        auto it = cache.find(k);
        if (it != cache.end()) return it->second;
        for (int c = 0; c < 3; c++) sh.positions.push_back(st.v[3 * (size_t)k.v + c]);
        if (k.vn >= 0)
            for (int c = 0; c < 3; c++) sh.normals.push_back(st.vn[3 * (size_t)k.vn + c]);
        if (k.vt >= 0)
            for (int c = 0; c < 2; c++) sh.texcoords.push_back(st.vt[2 * (size_t)k.vt + c]);
        const unsigned id = (unsigned)(sh.positions.size() / 3 - 1);
        cache[k] = id;
        return id;
    };
    for (const auto& f : st.faces) {
        if (f.size() < 3) continue;
        for (size_t k = 2; k < f.size(); k++) {   // triangle fan (0, k-1, k)
            const VIdx c[3] = {f[0], f[k - 1], f[k]};
            bool ok = true;
            for (const VIdx& q : c)
                ok = ok && q.v >= 0 && 3 * (size_t)q.v + 2 < st.v.size() && (q.vn < 0 || 3 * (size_t)q.vn + 2 < st.vn.size()) &&
                     (q.vt < 0 || 2 * (size_t)q.vt + 1 < st.vt.size());
            if (!ok) continue;   // the reference asserts / reads out of bounds here
            const unsigned a = vertex(c[0]), b = vertex(c[1]), d = vertex(c[2]);
            sh.indices.push_back(a);
            sh.indices.push_back(b);
            sh.indices.push_back(d);
        }
    }
    sh.name = st.name;
    shapes.push_back(std::move(sh));
    return true;
}
}  // namespace

bool load_obj(const std::string& filename, std::vector<ObjShape>& shapes, std::string& err) {
    shapes.clear();
    std::ifstream ifs(filename.c_str());
    if (!ifs) {
        err = "cannot open file [" + filename + "]";
        return false;
    }
    ObjState st;
    std::string line;
    while (std::getline(ifs, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (line.empty()) continue;
        const char* t = line.c_str();
        t += strspn(t, " \t");
        if (*t == '\0' || *t == '#') continue;

        if (t[0] == 'v' && is_blank(t[1])) {
            t += 2;
            for (int c = 0; c < 3; c++) st.v.push_back(next_float(t));
            continue;
        }
        if (t[0] == 'v' && t[1] == 'n' && is_blank(t[2])) {
            t += 3;
            for (int c = 0; c < 3; c++) st.vn.push_back(next_float(t));
            continue;
        }
        if (t[0] == 'v' && t[1] == 't' && is_blank(t[2])) {
            t += 3;
            for (int c = 0; c < 2; c++) st.vt.push_back(next_float(t));
            continue;
        }
        if (t[0] == 'f' && is_blank(t[1])) {
            t += 2;
            t += strspn(t, " \t");
            std::vector<VIdx> face;
            while (!is_eol(*t)) {
                face.push_back(parse_corner(t, (int)(st.v.size() / 3), (int)(st.vn.size() / 3), (int)(st.vt.size() / 2)));
                t += strspn(t, " \t\r");
            }
            st.faces.push_back(std::move(face));
            continue;
        }
        // usemtl, g and o all close the running shape; only the last two rename it.  mtllib is read by the reference
        // but nothing of the .mtl reaches the renderer (materials come from the .scene), so it is skipped here.
        if (strncmp(t, "usemtl", 6) == 0 && is_blank(t[6])) {
            flush_shape(st, shapes);
            st.faces.clear();
            continue;
        }
        if (t[0] == 'g' && is_blank(t[1])) {
            flush_shape(st, shapes);
            st.faces.clear();
            std::vector<std::string> names;
            while (!is_eol(*t)) {
                t += strspn(t, " \t");
                const size_t e = strcspn(t, " \t\r");
                names.emplace_back(t, t + e);
                t += e;
                t += strspn(t, " \t\r");
            }
            st.name = names.size() > 1 ? names[1] : std::string();
            continue;
        }
        if (t[0] == 'o' && is_blank(t[1])) {
            flush_shape(st, shapes);
            st.faces.clear();
            t += 2;
            t += strspn(t, " \t");
            st.name.assign(t, t + strcspn(t, " \t\r"));
            continue;
        }
    }
    flush_shape(st, shapes);
    return true;
}

}  // namespace spchost
