// scene_tool.cpp -- scene conversion and texture decoding without a GPU (no dependency on libspcbpt_b200.so):
//   spc_scene_tool convert <file.scene> <out.spcscene> [--data-root dir] [--K-light n]   .scene + OBJ + textures -> cache
//   spc_scene_tool info    <file.spcscene> <out.txt>                                    cache -> one summary line (loader check)
//   spc_scene_tool state   <in_prefix> <out_prefix> <K>                                 trained state (tree_eye/tree_light/Q/E .txt) read and re-written
//   spc_scene_tool camera  <file.spcscene> <out.txt> <width> <height>                   eye, U, V, W of the launch parameters (Camera::UVWFrame check)
//   spc_scene_tool relmse  <image.pfm> <reference.pfm>                                  relMSE of a render against a reference image (stdout)
//   spc_scene_tool decode  <image> <out.rgba8>                                          JPEG/PNG/PNM -> raw RGBA8 cache
//   spc_scene_tool png     <image> <out.png>                                            re-encode through the driver's PNG writer
//   spc_scene_tool scene   <file.scene> <out.txt> [data-root]                          parsed .scene as text (LoadScene check)
//   spc_scene_tool obj     <file.obj> <out.bin>      shapes as: u32 n_shapes; per shape u32 nv, nt, nuv; f32 pos[3nv]; u32 idx[3nt]; f32 uv[nuv]
// Used by tests/test_host_loader.py to compare the loaders with the reference's own tinyobj / stb_image / LoadScene.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "host_scene.hpp"
#include "train_state.hpp"

using namespace spchost;

int main(int argc, char** argv) {
    if (argc < 4) {
        fprintf(stderr, "usage: %s convert <file.scene> <out.spcscene> [--data-root dir] [--K-light n] | decode <image> <out.rgba8> | obj <file.obj> <out.bin>\n", argv[0]);
        return 2;
    }
    const std::string cmd = argv[1];
    std::string err;
    if (cmd == "convert") {
        std::string root;
        int k_light = 200;
        for (int i = 4; i + 1 < argc; i += 2) {
            if (!strcmp(argv[i], "--data-root")) root = argv[i + 1];
            if (!strcmp(argv[i], "--K-light")) k_light = atoi(argv[i + 1]);
        }
        SceneFile sf;
        HostScene hs;
        if (!load_scene_file(argv[2], root, sf, err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
        build_host_scene(sf, k_light, hs);
        for (const auto& w : hs.warnings) fprintf(stderr, "warning: %s\n", w.c_str());
        if (!save_scene_cache(argv[3], hs)) { fprintf(stderr, "cannot write %s\n", argv[3]); return 1; }
        printf("%zu meshes %zu triangles %zu materials %zu lights %zu textures\n", hs.meshes.size(), hs.n_triangles(), hs.materials.size(), hs.lights.size(), hs.textures.size());
        return 0;
    }
    if (cmd == "state") {   // host/train_state.cpp round trip: the reader and the writer of the reference's checkpoint text files
        if (argc < 5) { fprintf(stderr, "usage: %s state <in_prefix> <out_prefix> <K>\n", argv[0]); return 2; }
        TrainState st;
        if (!load_train_state(argv[2], atoi(argv[4]), st, err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
        if (!save_train_state(argv[3], st)) { fprintf(stderr, "cannot write %s*.txt\n", argv[3]); return 1; }
        printf("%zu + %zu tree nodes, %zu Q, %zu Gamma\n", st.eye_tree.size(), st.light_tree.size(), st.Q.size(), st.gamma.size());
        return 0;
    }
    if (cmd == "relmse") {   // the error metric of the equal-time comparisons: mean((I - R)^2 / (R^2 + 0.01)) over two .pfm files of the driver
        std::vector<float> a, b;
        int wa, ha, wb, hb;
        if (!read_pfm_rgb(argv[2], a, wa, ha, err) || !read_pfm_rgb(argv[3], b, wb, hb, err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
        if (wa != wb || ha != hb) { fprintf(stderr, "image sizes differ: %dx%d vs %dx%d\n", wa, ha, wb, hb); return 1; }
        size_t skipped = 0;
        const double e = rel_mse(a, b, &skipped);
        printf("relMSE %.9g over %dx%d pixels (%zu non-finite terms skipped)\n", e, wa, ha, skipped);
        return 0;
    }
    if (cmd == "camera") {
        if (argc < 6) { fprintf(stderr, "usage: %s camera <file.spcscene> <out.txt> <width> <height>\n", argv[0]); return 2; }
        HostScene hs;
        if (!load_scene_cache(argv[2], hs, err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
        float U[3], V[3], W[3];
        hs.camera_frame(atoi(argv[4]), atoi(argv[5]), U, V, W);
        FILE* f = fopen(argv[3], "w");
        if (!f) return 1;
        fprintf(f, "%.9g %.9g %.9g\n%.9g %.9g %.9g\n%.9g %.9g %.9g\n%.9g %.9g %.9g\n", hs.eye[0], hs.eye[1], hs.eye[2], U[0], U[1], U[2], V[0], V[1], V[2], W[0], W[1], W[2]);
        fclose(f);
        return 0;
    }
    if (cmd == "info") {
        HostScene hs;
        if (!load_scene_cache(argv[2], hs, err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
        FILE* f = fopen(argv[3], "w");
        if (!f) return 1;
        size_t texels = 0;
        for (const auto& t : hs.textures) texels += (size_t)t.width * t.height;
        fprintf(f, "%zu meshes %zu triangles %zu materials %zu lights %zu textures %zu texels %zu upload bytes\n", hs.meshes.size(), hs.n_triangles(), hs.materials.size(),
                hs.lights.size(), hs.textures.size(), texels, hs.upload_bytes());
        fclose(f);
        return 0;
    }
    if (cmd == "decode") {
        ImageRGBA8 img;
        if (!load_image_rgba8(argv[2], img, err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
        if (!write_rgba8_cache(argv[3], img)) { fprintf(stderr, "cannot write %s\n", argv[3]); return 1; }
        printf("%d %d\n", img.width, img.height);
        return 0;
    }
    if (cmd == "png") {   // image -> PNG through the driver's writer (frame-buffer convention: row 0 = bottom row)
        ImageRGBA8 img;
        if (!load_image_rgba8(argv[2], img, err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
        std::vector<uint32_t> frame((size_t)img.width * img.height);
        for (int y = 0; y < img.height; y++)
            memcpy(&frame[(size_t)(img.height - 1 - y) * img.width], &img.rgba[(size_t)y * img.width * 4], (size_t)img.width * 4);
        if (!write_png_from_uchar4(argv[3], frame.data(), img.width, img.height)) { fprintf(stderr, "cannot write %s\n", argv[3]); return 1; }
        printf("%d %d\n", img.width, img.height);
        return 0;
    }
    if (cmd == "obj") {
        std::vector<ObjShape> shapes;
        if (!load_obj(argv[2], shapes, err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
        FILE* f = fopen(argv[3], "wb");
        if (!f) return 1;
        const uint32_t n = (uint32_t)shapes.size();
        fwrite(&n, 4, 1, f);
        for (const auto& s : shapes) {
            const uint32_t h[3] = {(uint32_t)(s.positions.size() / 3), (uint32_t)(s.indices.size() / 3), (uint32_t)s.texcoords.size()};
            fwrite(h, 4, 3, f);
            if (!s.positions.empty()) fwrite(s.positions.data(), 4, s.positions.size(), f);
            if (!s.indices.empty()) fwrite(s.indices.data(), 4, s.indices.size(), f);
            if (!s.texcoords.empty()) fwrite(s.texcoords.data(), 4, s.texcoords.size(), f);
        }
        fclose(f);
        printf("%u shapes\n", n);
        return 0;
    }
    if (cmd == "scene") {   // the parsed .scene in the text form oracle/ref_shim/ref_loader.cpp prints for the reference's LoadScene
        SceneFile s;
        if (!load_scene_file(argv[2], argc > 4 ? argv[4] : "", s, err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
        FILE* f = fopen(argv[3], "w");
        if (!f) return 1;
        fprintf(f, "camera %.9g %.9g %.9g  %.9g %.9g %.9g  %.9g %.9g %.9g  %.9g %d\n", s.eye[0], s.eye[1], s.eye[2], s.lookat[0], s.lookat[1], s.lookat[2], s.up[0], s.up[1], s.up[2],
                s.fov, (int)s.use_geometry_normal);
        for (const auto& m : s.mesh_names) fprintf(f, "mesh %s\n", m.c_str());
        for (const auto& m : s.materials)
            fprintf(f, "material %d %.9g %.9g %.9g %.9g %.9g %.9g %.9g %d\n", m.albedoID, m.color[0], m.color[1], m.color[2], m.metallic, m.roughness, m.specular, m.clearcoatGloss, m.brdf);
        for (const auto& l : s.lights)
            fprintf(f, "light %d %d %.9g %.9g %.9g  %.9g %.9g %.9g  %.9g %.9g %.9g  %.9g %.9g %.9g  %.9g %.9g %.9g %.9g\n", l.lightType, l.divLevel, l.position[0], l.position[1],
                    l.position[2], l.u[0], l.u[1], l.u[2], l.v[0], l.v[1], l.v[2], l.emission[0], l.emission[1], l.emission[2], l.normal[0], l.normal[1], l.normal[2], l.area);
        for (const auto& t : s.texture_map) fprintf(f, "texture %d %s\n", t.first, t.second.c_str());
        fclose(f);
        return 0;
    }
    fprintf(stderr, "unknown command %s\n", cmd.c_str());
    return 2;
}
