// host_scene.cpp -- see host_scene.hpp.  Reference behaviour followed: scene_shift.cpp:32-329, sutil/Camera.cpp:32-43.
#include "host_scene.hpp"

#include <cmath>
#include <cstdio>
#include <cstring>

namespace spchost {

size_t HostScene::n_triangles() const {
    size_t n = 0;
    for (const auto& m : meshes) n += m.indices.size() / 3;
    return n;
}

size_t HostScene::upload_bytes() const {
    size_t n = materials.size() * sizeof(spc_pbr) + lights.size() * sizeof(spc_light);
    for (const auto& m : meshes) n += (m.positions.size() + m.texcoords.size()) * sizeof(float) + m.indices.size() * sizeof(uint32_t);
    for (const auto& t : textures) n += (size_t)t.width * t.height * 4;
    return n;
}

void HostScene::compute_aabb() {
    bool first = true;
    for (const auto& m : meshes)
        for (size_t i = 0; i + 2 < m.positions.size(); i += 3) {
            for (int c = 0; c < 3; c++) {
                const float x = m.positions[i + c];
                if (first || x < aabb_min[c]) aabb_min[c] = x;
                if (first || x > aabb_max[c]) aabb_max[c] = x;
            }
            first = false;
        }
}

void HostScene::abi_views(std::vector<spc_mesh>& mo, std::vector<spc_texture>& to) const {
    mo.clear();
    to.clear();
    for (const auto& m : meshes) {
        spc_mesh v;
        v.positions = m.positions.data();
        v.indices = m.indices.data();
        v.texcoords = m.texcoords.empty() ? nullptr : m.texcoords.data();
        v.n_vertices = (uint32_t)(m.positions.size() / 3);
        v.n_triangles = (uint32_t)(m.indices.size() / 3);
        v.material_id = m.material_id;
        v.light_id = m.light_id;
        mo.push_back(v);
    }
    for (const auto& t : textures) to.push_back(spc_texture{t.rgba.data(), t.width, t.height});
}

static inline void cross3(const float* a, const float* b, float* r) {
    const float x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
    r[0] = x; r[1] = y; r[2] = z;
}
static inline float dot3(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void normalize3(float* a) {
    const float inv = 1.0f / sqrtf(dot3(a, a));
    a[0] *= inv; a[1] *= inv; a[2] *= inv;
}

void HostScene::camera_frame(int width, int height, float U[3], float V[3], float W[3]) const {
    const float aspect = (float)width / (float)height;
    for (int c = 0; c < 3; c++) W[c] = lookat[c] - eye[c];   // not normalised: its length is the focal distance
    const float wlen = sqrtf(dot3(W, W));
    cross3(W, up, U);
    normalize3(U);
    cross3(U, W, V);
    normalize3(V);
    const float vlen = wlen * tanf(0.5f * fov * 3.14159265358979323846f / 180.0f);
    for (int c = 0; c < 3; c++) V[c] *= vlen;
    const float ulen = vlen * aspect;
    for (int c = 0; c < 3; c++) U[c] *= ulen;
}

static spc_pbr default_pbr() {   // MaterialData() (src/cuda/MaterialData.h:41-58)
    spc_pbr p;
    memset(&p, 0, sizeof(p));
    p.base_color[0] = p.base_color[1] = p.base_color[2] = p.base_color[3] = 1.0f;
    p.metallic = 1.0f;
    p.roughness = 1.0f;
    p.specular = 0.5f;
    p.sheenTint = 0.5f;
    p.clearcoatGloss = 1.0f;
    return p;
}

bool build_host_scene(const SceneFile& src, int K_light, HostScene& dst) {
    dst = HostScene();
    dst.warnings = src.warnings;

    // ---- Material_shift: textures, then one material per mesh block ------------------------------------------
    std::vector<int> tex_slot(src.texture_map.size() + 1, 0);   // albedoID -> 1 + index into dst.textures (0 = none)
    for (size_t i = 0; i < src.texture_map.size(); i++) {
        auto it = src.texture_map.find((int)i);
        if (it == src.texture_map.end()) continue;
        ImageRGBA8 img;
        std::string err;
        if (!load_image_rgba8(src.data_root + "/" + it->second, img, err)) {
            dst.warnings.push_back("texture skipped (" + err + "): materials using it fall back to their colour");
            continue;
        }
        dst.textures.push_back(std::move(img));
        tex_slot[i + 1] = (int)dst.textures.size();
    }
    for (const MaterialParameter& p : src.materials) {
        spc_pbr m = default_pbr();
        m.base_color[0] = p.color[0];
        m.base_color[1] = p.color[1];
        m.base_color[2] = p.color[2];
        m.base_color[3] = 1.0f;
        m.metallic = p.metallic;     // only colour, metallic, roughness, brdf and the albedo texture cross over
        m.roughness = p.roughness;   // (scene_shift.cpp:70-86); the other Disney terms keep the MaterialData defaults
        m.brdf = (uint8_t)(p.brdf != 0);
        if (p.albedoID > 0 && p.albedoID < (int)tex_slot.size() && tex_slot[p.albedoID] > 0) {
            m.base_color_tex.tex = (uint64_t)tex_slot[p.albedoID];
            m.base_color_tex.texcoord = 0;
            m.base_color_tex.texcoord_rotation[0] = sinf(0.0f);
            m.base_color_tex.texcoord_rotation[1] = cosf(0.0f);
            m.base_color_tex.texcoord_scale[0] = m.base_color_tex.texcoord_scale[1] = 1.0f;
        }
        dst.materials.push_back(m);
    }
    if (dst.materials.empty()) dst.materials.push_back(default_pbr());

    // ---- LightSource_shift: quad lights only (directional lights go to a list the SPCBPT path never reads) -----
    int ss_base = !src.env_file.empty() ? (int)(0.5 * K_light) : 0;
    std::vector<int> light_of(src.lights.size(), -1);
    for (size_t i = 0; i < src.lights.size(); i++) {
        const LightParameter& s = src.lights[i];
        if (s.lightType != LK_QUAD) {
            if (s.lightType != LK_NONE) dst.warnings.push_back("non-quad light ignored (environment / directional / sphere lights are unfinished in the reference)");
            continue;
        }
        spc_light L;
        memset(&L, 0, sizeof(L));
        L.type = SPC_LIGHT_QUAD;
        L.id = (int)dst.lights.size();
        L.ssBase = ss_base;
        L.divLevel = s.divLevel;
        ss_base += s.divLevel * s.divLevel;
        float n[3];
        cross3(s.u, s.v, n);
        L.area = sqrtf(dot3(n, n));
        normalize3(n);
        L.corner = {s.position[0], s.position[1], s.position[2]};
        L.u = {s.position[0] + s.u[0], s.position[1] + s.u[1], s.position[2] + s.u[2]};   // sic: corner + u, corner + v
        L.v = {s.position[0] + s.v[0], s.position[1] + s.v[1], s.position[2] + s.v[2]};
        L.emission = {s.emission[0], s.emission[1], s.emission[2]};
        L.normal = {n[0], n[1], n[2]};
        light_of[i] = L.id;
        dst.lights.push_back(L);
    }
    if (!src.env_file.empty()) dst.warnings.push_back("env_file ignored: environment lighting is unfinished in the reference (readme.md:28)");

    // ---- Camera_shift ---------------------------------------------------------------------------------------------
    if (src.use_camera) {
        memcpy(dst.eye, src.eye, sizeof(dst.eye));
        memcpy(dst.lookat, src.lookat, sizeof(dst.lookat));
        memcpy(dst.up, src.up, sizeof(dst.up));
        dst.fov = src.fov;
    }

    // ---- Geometry_shift: shapes of every mesh file, then the light quads ----------------------------------------------
    for (size_t k = 0; k < src.mesh_names.size(); k++) {
        std::vector<ObjShape> shapes;
        std::string err;
        if (!load_obj(src.mesh_names[k], shapes, err)) {
            dst.warnings.push_back("mesh skipped: " + err);
            continue;
        }
        for (ObjShape& sh : shapes) {
            HostMesh m;
            m.name = sh.name;
            m.positions = std::move(sh.positions);
            m.indices.assign(sh.indices.begin(), sh.indices.end());
            m.texcoords = std::move(sh.texcoords);
            const size_t nv = m.positions.size() / 3;
            if (m.texcoords.size() < nv * 2) m.texcoords.resize(nv * 2, 0.0f);
            m.texcoords.resize(nv * 2);
            m.material_id = k < dst.materials.size() ? (int)k : 0;
            dst.meshes.push_back(std::move(m));
        }
    }
    for (size_t i = 0; i < src.lights.size(); i++) {
        if (light_of[i] < 0) continue;
        const spc_light& L = dst.lights[light_of[i]];
        HostMesh m;
        m.name = "quad_light";
        const float c[3] = {L.corner.x, L.corner.y, L.corner.z}, u[3] = {L.u.x, L.u.y, L.u.z}, v[3] = {L.v.x, L.v.y, L.v.z};
        m.positions = {c[0], c[1], c[2], u[0], u[1], u[2], v[0], v[1], v[2], u[0] + v[0] - c[0], u[1] + v[1] - c[1], u[2] + v[2] - c[2]};
        m.indices = {0, 1, 3, 0, 3, 2};
        m.texcoords = {0, 0, 1, 0, 0, 1, 1, 1};
        m.light_id = L.id;
        dst.meshes.push_back(std::move(m));
    }

    dst.compute_aabb();
    return true;
}

// ---------------------------------------------------------------------------------------------------------------------
// .spcscene: little endian.
//   char magic[8] = "SPCSCN01"; u32 n_meshes, n_materials, n_lights, n_textures; f32 eye[3], lookat[3], up[3], fov;
//   per mesh:    u32 n_vertices, n_triangles; i32 material_id, light_id; f32 pos[3 nv]; u32 idx[3 nt]; f32 uv[2 nv]
//   spc_pbr materials[n_materials] (144 B each); spc_light lights[n_lights] (80 B each)
//   per texture: i32 width, height; u8 rgba[4 w h]
// ---------------------------------------------------------------------------------------------------------------------
namespace {
struct Writer {
    FILE* f;
    bool ok = true;
    void put(const void* p, size_t n) { ok = ok && (n == 0 || fwrite(p, 1, n, f) == n); }
    template <class T> void val(T v) { put(&v, sizeof(T)); }
};
struct Reader {
    FILE* f;
    bool ok = true;
    long size = -1;     // of the whole file, when known: a count read from the file may not announce more than the file still holds
    void get(void* p, size_t n) { ok = ok && (n == 0 || fread(p, 1, n, f) == n); }
    template <class T> T val() { T v{}; get(&v, sizeof(T)); return v; }
    bool holds(uint64_t bytes) {   // call before sizing a buffer by a count from the file
        if (size >= 0) {
            const long at = ftell(f);
            if (at < 0 || bytes > (uint64_t)(size - at)) ok = false;
        }
        return ok;
    }
};
}  // namespace

bool save_scene_cache(const std::string& path, const HostScene& s) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) return false;
    Writer w{f};
    w.put("SPCSCN01", 8);
    w.val<uint32_t>((uint32_t)s.meshes.size());
    w.val<uint32_t>((uint32_t)s.materials.size());
    w.val<uint32_t>((uint32_t)s.lights.size());
    w.val<uint32_t>((uint32_t)s.textures.size());
    w.put(s.eye, 12); w.put(s.lookat, 12); w.put(s.up, 12); w.val<float>(s.fov);
    for (const auto& m : s.meshes) {
        w.val<uint32_t>((uint32_t)(m.positions.size() / 3));
        w.val<uint32_t>((uint32_t)(m.indices.size() / 3));
        w.val<int32_t>(m.material_id);
        w.val<int32_t>(m.light_id);
        w.put(m.positions.data(), m.positions.size() * 4);
        w.put(m.indices.data(), m.indices.size() * 4);
        w.put(m.texcoords.data(), m.texcoords.size() * 4);
    }
    w.put(s.materials.data(), s.materials.size() * sizeof(spc_pbr));
    w.put(s.lights.data(), s.lights.size() * sizeof(spc_light));
    for (const auto& t : s.textures) {
        w.val<int32_t>(t.width);
        w.val<int32_t>(t.height);
        w.put(t.rgba.data(), t.rgba.size());
    }
    fclose(f);
    return w.ok;
}

bool load_scene_cache(const std::string& path, HostScene& s, std::string& err) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) {
        err = "cannot open " + path;
        return false;
    }
    Reader r{f};
    if (fseek(f, 0, SEEK_END) == 0) r.size = ftell(f);
    fseek(f, 0, SEEK_SET);
    char magic[8];
    r.get(magic, 8);
    if (!r.ok || memcmp(magic, "SPCSCN01", 8) != 0) {
        fclose(f);
        err = "not a .spcscene file: " + path;
        return false;
    }
    s = HostScene();
    const uint32_t nm = r.val<uint32_t>(), nmat = r.val<uint32_t>(), nl = r.val<uint32_t>(), nt = r.val<uint32_t>();
    r.get(s.eye, 12); r.get(s.lookat, 12); r.get(s.up, 12); s.fov = r.val<float>();
    for (uint32_t i = 0; i < nm && r.ok; i++) {
        HostMesh m;
        const uint32_t nv = r.val<uint32_t>(), ntri = r.val<uint32_t>();
        m.material_id = r.val<int32_t>();
        m.light_id = r.val<int32_t>();
        if (!r.ok || nv > (1u << 30) || ntri > (1u << 30) || !r.holds((uint64_t)nv * 20 + (uint64_t)ntri * 12)) { r.ok = false; break; }
        m.positions.resize((size_t)nv * 3);
        m.indices.resize((size_t)ntri * 3);
        m.texcoords.resize((size_t)nv * 2);
        r.get(m.positions.data(), m.positions.size() * 4);
        r.get(m.indices.data(), m.indices.size() * 4);
        r.get(m.texcoords.data(), m.texcoords.size() * 4);
        s.meshes.push_back(std::move(m));
    }
    if (r.ok && nmat < (1u << 24) && nl < (1u << 24) && r.holds((uint64_t)nmat * sizeof(spc_pbr) + (uint64_t)nl * sizeof(spc_light))) {
        s.materials.resize(nmat);
        s.lights.resize(nl);
        r.get(s.materials.data(), nmat * sizeof(spc_pbr));
        r.get(s.lights.data(), nl * sizeof(spc_light));
    } else {
        r.ok = false;
    }
    for (uint32_t i = 0; i < nt && r.ok; i++) {
        ImageRGBA8 t;
        t.width = r.val<int32_t>();
        t.height = r.val<int32_t>();
        if (!r.ok || t.width <= 0 || t.height <= 0 || (size_t)t.width * t.height > (1u << 28) || !r.holds((uint64_t)t.width * t.height * 4)) { r.ok = false; break; }
        t.rgba.resize((size_t)t.width * t.height * 4);
        r.get(t.rgba.data(), t.rgba.size());
        s.textures.push_back(std::move(t));
    }
    fclose(f);
    s.compute_aabb();
    if (!r.ok) err = "truncated or corrupt .spcscene file: " + path;
    return r.ok;
}

}  // namespace spchost
