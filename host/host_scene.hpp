// host_scene.hpp -- the scene as the render core wants it: the arrays spc_scene_upload() takes.
//
// Stands in for src/OptiXPathTracer/scene_shift.{h,cpp} (Scene_shift = Material_shift + Camera_shift + Geometry_shift,
// and LightSource_shift), which translate the parsed .scene into a sutil::Scene, and for the camera set-up of
// optixPathTracer.cpp (initCameraState :670-680, handleCameraUpdate :352-369, sutil/Camera.cpp:32-43).
// Mesh order = every shape of every mesh file in file order, then one two-triangle mesh per quad light
// (scene_shift.cpp:187-328); this order defines the global primitive ids.
#pragma once
#include <string>
#include <vector>

#include "image_io.hpp"
#include "scene_file.hpp"
#include "spcbpt_b200.h"

namespace spchost {

struct HostMesh {
    std::string           name;
    std::vector<float>    positions;   // 3 per vertex
    std::vector<uint32_t> indices;     // 3 per triangle
    std::vector<float>    texcoords;   // 2 per vertex, zero-padded like scene_shift.cpp:203-206
    int material_id = -1, light_id = -1;
};

struct HostScene {
    std::vector<HostMesh>   meshes;
    std::vector<spc_pbr>    materials;
    std::vector<spc_light>  lights;
    std::vector<ImageRGBA8> textures;
    float eye[3] = {1, 1, 1}, lookat[3] = {0, 0, 0}, up[3] = {0, 1, 0}, fov = 35.0f;   // sutil::Camera() defaults
    float aabb_min[3] = {0, 0, 0}, aabb_max[3] = {0, 0, 0};
    std::vector<std::string> warnings;

    size_t n_triangles() const;
    size_t upload_bytes() const;   // host bytes spc_scene_upload reads: positions + indices + texcoords + materials + lights + textures
    void   compute_aabb();
    // views for spc_scene_upload; valid while *this is alive and unmodified
    void abi_views(std::vector<spc_mesh>& meshes_out, std::vector<spc_texture>& textures_out) const;
    // sutil::Camera::UVWFrame with aspect = width / height (float division, optixPathTracer.cpp:358)
    void camera_frame(int width, int height, float U[3], float V[3], float W[3]) const;
};

// K_light = NUM_SUBSPACE_LIGHTSOURCE: only used for ssBase when the scene has an environment map (scene_shift.cpp:110)
bool build_host_scene(const SceneFile& src, int K_light, HostScene& dst);

// binary scene cache (".spcscene"): the arrays above, so that a parsed + decoded scene can be reloaded without the
// OBJ / JPEG files.  Layout is documented in host_scene.cpp and read by scenes.py too.
bool save_scene_cache(const std::string& path, const HostScene& s);
bool load_scene_cache(const std::string& path, HostScene& s, std::string& err);

}  // namespace spchost
