// orc_scene.h -- CPU oracle: scene container + ray/triangle queries.  TEST INFRASTRUCTURE ONLY.
//
// Restates, on the host, what the reference gets from OptiX (closed source, OptiX 7.5.0, call
// sites cuProg.h:395,420,445,470 and sutil/Scene.cpp:1169,1232,1321).  OptiX's own arithmetic is
// not reproducible ("parity unpinned" at this boundary); the closest-hit
// contract is therefore the one written in csrc/traverse.cuh and restated here in orc_tri_test:
// nearest t in (tmin,tmax), ties -> lowest global prim id, back-face culling only on emitter quads.
#pragma once
#include <vector>
#include "../include/spcbpt_b200.h"
#include "orc_math.h"

namespace orc {

struct Tri {
    f3 v0, e1, e2;   // e1 = v1-v0, e2 = v2-v0 (one fp32 subtraction each)
    f3 v1, v2;
    float uv[3][2];
    int material;    // -1 for emitter quads
    int light;       // -1 for ordinary surfaces
    int mesh;
};

struct Hit {
    float t, u, v;
    int prim;
};

struct Bvh2Node {
    f3 lo, hi;
    int left, right;   // children (internal) or [first, count] (leaf)
    int first, count;
};

struct Texture {
    std::vector<uint8_t> rgba;
    int w, h;
};

struct Scene {
    std::vector<Tri> tris;
    std::vector<spc_pbr> materials;
    std::vector<spc_light> lights;
    std::vector<Texture> textures;
    std::vector<Bvh2Node> nodes;
    std::vector<int> order;     // BVH leaf order -> prim id
    float pad = 0.f;

    void build_bvh();
    bool closest(f3 o, f3 d, float tmin, float tmax, bool cull_back, Hit& h, bool brute = false) const;
    bool occluded(f3 o, f3 d, float tmin, float tmax, bool brute = false) const;
};

// the triangle test of the contract; returns true and fills t,u,v when tmin < t (no tmax test)
bool orc_tri_test(const Tri& tr, f3 o, f3 d, float tmin, bool cull_back, float& t, float& u, float& v);

}  // namespace orc
