// ref_loader.cpp -- TEST INFRASTRUCTURE: the reference's own scene-ingest code (sceneLoader.cpp LoadScene, tinyobjloader,
// stb_image) compiled from /root/reference into oracle/_ref/ref_loader.  tests/test_host_loader.py and
// tests/golden/make_golden_loader.py run it to pin host/ (scene_file.cpp, jpeg_decode.cpp, png_decode.cpp).
// Output formats are those of host/scene_tool.cpp so the files can be compared byte for byte.
//   ref_loader obj <file.obj> <out.bin> | decode <image> <out.rgba8> | scene <file.scene> <out.txt>
#define TINYOBJLOADER_IMPLEMENTATION   // tiny_obj_loader.cc of the reference is exactly this define + include
#include "sceneLoader.cpp"
#define STB_IMAGE_IMPLEMENTATION
#include "stb_image.h"

#include <cstdint>

int main(int argc, char** argv) {
    if (argc < 4) return 2;
    const std::string cmd = argv[1];
    if (cmd == "obj") {
        std::vector<tinyobj::shape_t> shapes;
        std::vector<tinyobj::material_t> mats;
        std::string err;
        tinyobj::LoadObj(shapes, mats, err, argv[2]);
        FILE* f = fopen(argv[3], "wb");
        const uint32_t n = (uint32_t)shapes.size();
        fwrite(&n, 4, 1, f);
        for (auto& s : shapes) {
            const uint32_t h[3] = {(uint32_t)(s.mesh.positions.size() / 3), (uint32_t)(s.mesh.indices.size() / 3), (uint32_t)s.mesh.texcoords.size()};
            fwrite(h, 4, 3, f);
            fwrite(s.mesh.positions.data(), 4, s.mesh.positions.size(), f);
            fwrite(s.mesh.indices.data(), 4, s.mesh.indices.size(), f);
            fwrite(s.mesh.texcoords.data(), 4, s.mesh.texcoords.size(), f);
        }
        fclose(f);
        return 0;
    }
    if (cmd == "decode") {
        int w, h, ch;
        stbi_uc* px = stbi_load(argv[2], &w, &h, &ch, STBI_rgb_alpha);
        if (!px) return 1;
        FILE* f = fopen(argv[3], "wb");
        const int32_t wh[2] = {w, h};
        fwrite("SPCRGBA8", 1, 8, f);
        fwrite(wh, 4, 2, f);
        fwrite(px, 1, (size_t)w * h * 4, f);
        fclose(f);
        return 0;
    }
    if (cmd == "scene") {
        Scene* s = LoadScene(argv[2]);
        if (!s) return 1;
        FILE* f = fopen(argv[3], "w");
        fprintf(f, "camera %.9g %.9g %.9g  %.9g %.9g %.9g  %.9g %.9g %.9g  %.9g %d\n", s->eye.x, s->eye.y, s->eye.z, s->lookat.x, s->lookat.y, s->lookat.z, s->up.x, s->up.y,
                s->up.z, s->fov, (int)s->use_geometry_normal);
        for (size_t i = 0; i < s->mesh_names.size(); i++) fprintf(f, "mesh %s\n", s->mesh_names[i].c_str());
        for (auto& m : s->materials)
            fprintf(f, "material %d %.9g %.9g %.9g %.9g %.9g %.9g %.9g %d\n", m.albedoID, m.color.x, m.color.y, m.color.z, m.metallic, m.roughness, m.specular, m.clearcoatGloss, (int)m.brdf);
        for (auto& l : s->lights)
            fprintf(f, "light %d %d %.9g %.9g %.9g  %.9g %.9g %.9g  %.9g %.9g %.9g  %.9g %.9g %.9g  %.9g %.9g %.9g %.9g\n", (int)l.lightType, l.divLevel, l.position.x, l.position.y,
                    l.position.z, l.u.x, l.u.y, l.u.z, l.v.x, l.v.y, l.v.z, l.emission.x, l.emission.y, l.emission.z, l.normal.x, l.normal.y, l.normal.z, l.area);
        for (auto& t : s->texture_map) fprintf(f, "texture %d %s\n", t.first, t.second.c_str());
        fclose(f);
        return 0;
    }
    return 2;
}
