// scene_file.hpp -- the reference's `.scene` text format and the OBJ meshes it names, read on the host.
//
// Stands in for src/OptiXPathTracer/sceneLoader.{h,cpp} (LoadScene, struct Scene) and for the subset of
// tinyobjloader 0.9.x the reference uses (tiny_obj_loader.h: LoadObj -> shapes with flattened, per-shape
// de-duplicated vertices and fan-triangulated faces).  Written from the observed behaviour of those files; the
// quirks that decide triangle order, vertex order and float bits are kept and named at the place they occur,
// because primitive ids and vertex bits are part of this repo's parity contract.
#pragma once
#include <map>
#include <string>
#include <vector>

namespace spchost {

// material_parameters.h:17-33 (MaterialParameter defaults)
struct MaterialParameter {
    int   albedoID = 0;   // 0 = no texture; else 1 + index into SceneFile::texture_map
    float color[3] = {1.0f, 1.0f, 1.0f};
    float emission[3] = {0.0f, 0.0f, 0.0f};
    float metallic = 0.0f, subsurface = 0.0f, specular = 0.5f, roughness = 0.5f, specularTint = 0.0f, anisotropic = 0.0f;
    float sheen = 0.0f, sheenTint = 0.5f, clearcoat = 0.0f, clearcoatGloss = 1.0f;
    int   brdf = 0;       // 0 DISNEY, 1 GLASS (unfinished in the reference, readme.md:28)
};

// light_parameters.h:8-32 (LightType / LightParameter)
enum LightKind { LK_SPHERE = 0, LK_QUAD = 1, LK_DIRECTION = 2, LK_ENV = 3, LK_NONE = -1 };
struct LightParameter {
    float position[3] = {0, 0, 0}, normal[3] = {0, 0, 0}, emission[3] = {0, 0, 0};
    float u[3] = {0, 0, 0}, v[3] = {0, 0, 0}, direction[3] = {0, 0, 0};
    int   lightType = LK_NONE;
    float area = 0.0f, radius = 0.0f;
    int   divLevel = 1;
};

// sceneLoader.h:44-71 (struct Scene)
struct SceneFile {
    std::vector<std::string>       mesh_names;      // resolved paths (data_root + "/" + name, '\\' -> '/')
    std::vector<std::string>       uv_mesh_names;   // parsed like the reference, unused by it as well
    std::vector<MaterialParameter> materials;       // one per mesh block, in mesh order (sceneLoader.cpp:283-294)
    std::vector<LightParameter>    lights;
    std::map<int, std::string>     texture_map;     // texture index -> file name relative to data_root
    std::string env_file;
    int   width = 1920, height = 1001;              // Properties defaults (sceneLoader.cpp:207-210)
    float eye[3] = {0, 0, 0}, lookat[3] = {0, 0, 0}, up[3] = {0, 1, 0};
    float fov = 35.0f;
    bool  use_camera = false, use_geometry_normal = false;
    float env_factor = 1.0f;
    std::string data_root;
    std::vector<std::string> warnings;
};

// LoadScene(filename) (sceneLoader.cpp:47-308).  `data_root` replaces SAMPLES_DIR "/data/"; when empty it is
// the parent of the directory that holds the .scene file (the shipped layout data/house/x.scene names files as
// house/geometry/...), falling back to the .scene's own directory.
bool load_scene_file(const std::string& filename, const std::string& data_root, SceneFile& out, std::string& err);

// one tinyobj shape_t::mesh (tiny_obj_loader.h:88-96): flattened arrays
struct ObjShape {
    std::string           name;
    std::vector<float>    positions;   // 3 per vertex
    std::vector<float>    normals;     // 3 per vertex that had a vn index
    std::vector<float>    texcoords;   // 2 per vertex that had a vt index (may be shorter than the vertex count)
    std::vector<unsigned> indices;     // 3 per triangle
};

// tinyobj::LoadObj(shapes, materials, err, filename) as Scene::getMeshData calls it (sceneLoader.cpp:331-341).
// Returns false (and no shapes) when the file cannot be opened -- the reference then simply has no shapes for
// that mesh (three of the shipped house meshes are absent).
bool load_obj(const std::string& filename, std::vector<ObjShape>& shapes, std::string& err);

// tinyobj's own decimal parser (tiny_obj_loader.h:213-344 tryParseDouble): digits accumulated in double, the
// fraction as sum(digit * 10^-k), then ldexp(m * 5^e, e).  It is NOT correctly rounded like strtod, so it is
// restated here to get the same float bits for every vertex.  Returns 0 on a parse failure, like parseFloat.
float parse_obj_float(const char* begin, const char* end);

std::string normalize_path(const std::string& p);   // '\\' -> '/'

}  // namespace spchost
