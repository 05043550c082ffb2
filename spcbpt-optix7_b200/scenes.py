"""Synthetic scene fixtures in the reference's data model (host side, numpy only).

The reference ships one scene (src/data/house, 3 meshes missing) and no Cornell box (SURVEY.md
header), so BASELINE.json's configs are made concrete here:
  * cornell_scene()      config 1: Cornell-class box, ~49 k triangles, one quad light divLevel 2
  * heightfield_scene()  config 2: fractal height field 708x708 quads (1 002 528 triangles) in a
                         closed box with a 1x1 quad light -- the traversal microbench mesh
Meshes follow scene_shift.cpp: one mesh per OBJ shape, then one 2-triangle mesh per quad light with
vertices (corner, u, v, u+v-corner), triangles (0,1,3),(0,3,2), uv (0,0),(1,0),(0,1),(1,1)
(scene_shift.cpp:274-293); Light.u / Light.v are corner+u, corner+v (scene_shift.cpp:127-129).
"""
import ctypes
import ctypes.util

import numpy as np


def _tanf(x):
    """C tanf (what sutil/Camera.cpp:39 and host/host_scene.cpp call); numpy's float32 tan may differ in the last bit"""
    try:
        libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
        libm.tanf.restype = ctypes.c_float
        libm.tanf.argtypes = [ctypes.c_float]
        return np.float32(libm.tanf(float(x)))
    except OSError:
        return np.float32(np.tan(np.float32(x)))


class SceneData:
    def __init__(self):
        self.meshes = []        # dicts: positions (n,3) f4, indices (m,3) u4, texcoords (n,2) f4|None, material_id, light_id
        self.materials = None   # structured array PBR
        self.lights = None      # structured array LIGHT
        self.textures = []      # list of (h,w,4) uint8
        self.camera = dict(eye=(0, 0, -1), lookat=(0, 0, 0), up=(0, 1, 0), fov=35.0)
        self.name = ""

    @property
    def n_triangles(self):
        return int(sum(m["indices"].shape[0] for m in self.meshes))

    def camera_frame(self, width, height):
        """sutil::Camera::UVWFrame (sutil/Camera.cpp:32-43) in the reference's own fp32 operation order (vec_math.h dot / cross /
        normalize written out on scalars: numpy's dot may sum in another order); tests/test_host_loader.py compares it, and the C++
        driver's HostScene::camera_frame, with the reference's Camera.cpp bit for bit."""
        f = np.float32

        def dot(a, b):
            return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]

        def cross(a, b):
            return [a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]]

        def normalize(a):
            inv = f(1) / f(np.sqrt(dot(a, a)))
            return [a[0] * inv, a[1] * inv, a[2] * inv]
        eye = [f(x) for x in self.camera["eye"]]
        lookat = [f(x) for x in self.camera["lookat"]]
        up = [f(x) for x in self.camera["up"]]
        W = [lookat[k] - eye[k] for k in range(3)]          # not normalised: its length is the focal distance
        wlen = f(np.sqrt(dot(W, W)))
        U = normalize(cross(W, up))
        V = normalize(cross(U, W))
        vlen = f(wlen * _tanf(f(0.5) * f(self.camera["fov"]) * f(np.pi) / f(180.0)))
        V = [v * vlen for v in V]
        ulen = f(vlen * (f(width) / f(height)))
        U = [u * ulen for u in U]
        return np.asarray(eye, f), np.asarray(U, f), np.asarray(V, f), np.asarray(W, f)


def make_pbr(n):
    """MaterialData() defaults (src/cuda/MaterialData.h:41-58) as set up by Material_shift
    (scene_shift.cpp:64-91): only color / metallic / roughness / brdf come from the .scene."""
    from . import PBR
    m = np.zeros(n, PBR)
    m["base_color"] = (1, 1, 1, 1)
    m["metallic"] = 0.0      # MaterialParameter() default, material_parameters.h:19
    m["roughness"] = 0.5     # material_parameters.h:22
    m["specular"] = 0.5
    m["specularTint"] = 0.0
    m["subsurface"] = 0.0
    m["anisotropic"] = 0.0
    m["sheen"] = 0.0
    m["sheenTint"] = 0.5
    m["clearcoat"] = 0.0
    m["clearcoatGloss"] = 1.0
    m["base_color_tex"]["texcoord_scale"] = (1, 1)
    m["base_color_tex"]["texcoord_rotation"] = (0, 1)   # sin 0, cos 0
    return m


def make_quad_light(light_id, corner, v1, v2, emission, div_level, ss_base):
    """LightSource_shift for a `light{ type Quad }` block (sceneLoader.cpp:166-173 computes u=v1-pos,
    v=v2-pos; scene_shift.cpp:121-131 stores corner+u, corner+v, normal, area)."""
    from . import LIGHT, LIGHT_QUAD
    f = np.float32
    corner = np.asarray(corner, f)
    u = np.asarray(v1, f) - corner
    v = np.asarray(v2, f) - corner
    # cross / length in the scalar fp32 order of the reference's vec_math.h (and of host/host_scene.cpp): numpy's dot may sum differently
    n = np.asarray([u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0]], f)
    area = f(np.sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]))
    L = np.zeros(1, LIGHT)
    L["type"] = LIGHT_QUAD
    L["id"] = light_id
    L["divLevel"] = div_level
    L["ssBase"] = ss_base
    L["corner"] = corner
    L["u"] = corner + u
    L["v"] = corner + v
    L["emission"] = emission
    L["normal"] = n * (f(1) / area)
    L["area"] = area
    return L


def light_mesh(L, light_id):
    corner = L["corner"][0]
    pu = L["u"][0]
    pv = L["v"][0]
    pos = np.stack([corner, pu, pv, pu + pv - corner]).astype(np.float32)
    idx = np.array([[0, 1, 3], [0, 3, 2]], np.uint32)
    uv = np.array([[0, 0], [1, 0], [0, 1], [1, 1]], np.float32)
    return dict(positions=pos, indices=idx, texcoords=uv, material_id=-1, light_id=light_id)


def _grid_quad(p0, du, dv, m):
    """m x m cells over the parallelogram p0 + s*du + t*dv, 2 triangles per cell."""
    f = np.float32
    s = np.linspace(0.0, 1.0, m + 1, dtype=np.float64)
    S, T = np.meshgrid(s, s, indexing="xy")
    pos = (np.asarray(p0, np.float64)[None, None, :] + S[..., None] * np.asarray(du, np.float64)
           + T[..., None] * np.asarray(dv, np.float64)).reshape(-1, 3).astype(f)
    uv = np.stack([S, T], -1).reshape(-1, 2).astype(f)
    i = np.arange(m)
    I, J = np.meshgrid(i, i, indexing="xy")
    a = (J * (m + 1) + I).reshape(-1)
    b = a + 1
    c = a + (m + 1)
    d = c + 1
    idx = np.concatenate([np.stack([a, b, d], 1), np.stack([a, d, c], 1)], 0).astype(np.uint32)
    return pos, idx, uv


def _merge(parts):
    pos, idx, uv, off = [], [], [], 0
    for p, i, t in parts:
        pos.append(p)
        idx.append(i + np.uint32(off))
        uv.append(t)
        off += p.shape[0]
    return np.concatenate(pos), np.concatenate(idx), np.concatenate(uv)


def _box(center_xz, size_xz, height, angle_deg, m):
    """five faces (top + 4 sides) of a box standing on y=0, rotated about y."""
    a = np.deg2rad(angle_deg)
    ca, sa = np.cos(a), np.sin(a)
    hx, hz = size_xz[0] / 2, size_xz[1] / 2

    def P(x, y, z):
        return (center_xz[0] + ca * x + sa * z, y, center_xz[1] - sa * x + ca * z)
    c = [P(-hx, 0, -hz), P(hx, 0, -hz), P(hx, 0, hz), P(-hx, 0, hz)]
    t = [P(-hx, height, -hz), P(hx, height, -hz), P(hx, height, hz), P(-hx, height, hz)]
    faces = [_grid_quad(t[0], np.subtract(t[1], t[0]), np.subtract(t[3], t[0]), m)]
    for k in range(4):
        p0, p1 = c[k], c[(k + 1) % 4]
        faces.append(_grid_quad(p0, np.subtract(p1, p0), (0, height, 0), m))
    return _merge(faces)


def scaled(sc, k):
    """uniformly scale a scene (geometry, lights, camera) by k"""
    k = np.float32(k)
    for m in sc.meshes:
        m["positions"] = (m["positions"] * k).astype(np.float32)
    lights = []
    for i in range(sc.lights.shape[0]):
        L = sc.lights[i]
        lights.append(make_quad_light(int(L["id"]), L["corner"] * k, L["u"] * k, L["v"] * k, L["emission"], int(L["divLevel"]), int(L["ssBase"])))
    sc.lights = np.concatenate(lights)
    li = 0
    for j, m in enumerate(sc.meshes):
        if m["light_id"] >= 0:
            sc.meshes[j] = light_mesh(sc.lights[m["light_id"]:m["light_id"] + 1], m["light_id"])
    sc.camera = dict(sc.camera, eye=tuple(np.asarray(sc.camera["eye"], np.float32) * k), lookat=tuple(np.asarray(sc.camera["lookat"], np.float32) * k))
    return sc


def cornell_scene(wall_cells=48, box_cells=36, div_level=2, K=64):
    """Cornell-class fixture of config 1: 5 walls, 2 boxes, one quad light
    divLevel 2 with emission (17,12,4); Disney materials roughness .5 metallic 0; camera eye
    (278,273,-800) lookat (278,273,0) fov 39.3.  Default tessellation: 48 962 triangles."""
    sc = SceneData()
    sc.name = "cornell"
    X, Y, Z = 556.0, 548.8, 559.2
    mats = make_pbr(3)
    mats["base_color"][0] = (0.73, 0.73, 0.73, 1)
    mats["base_color"][1] = (0.65, 0.05, 0.05, 1)
    mats["base_color"][2] = (0.12, 0.45, 0.15, 1)
    sc.materials = mats
    m = wall_cells
    white = _merge([
        _grid_quad((0, 0, 0), (X, 0, 0), (0, 0, Z), m),      # floor
        _grid_quad((0, Y, 0), (X, 0, 0), (0, 0, Z), m),      # ceiling
        _grid_quad((0, 0, Z), (X, 0, 0), (0, Y, 0), m),      # back
    ])
    sc.meshes.append(dict(positions=white[0], indices=white[1], texcoords=white[2], material_id=0, light_id=-1))
    left = _grid_quad((X, 0, 0), (0, 0, Z), (0, Y, 0), m)
    sc.meshes.append(dict(positions=left[0], indices=left[1], texcoords=left[2], material_id=1, light_id=-1))
    right = _grid_quad((0, 0, 0), (0, 0, Z), (0, Y, 0), m)
    sc.meshes.append(dict(positions=right[0], indices=right[1], texcoords=right[2], material_id=2, light_id=-1))
    short = _box((185.0, 169.0), (165.0, 165.0), 165.0, -17.0, box_cells)
    sc.meshes.append(dict(positions=short[0], indices=short[1], texcoords=short[2], material_id=0, light_id=-1))
    tall = _box((368.0, 351.0), (165.0, 165.0), 330.0, 18.0, box_cells)
    sc.meshes.append(dict(positions=tall[0], indices=tall[1], texcoords=tall[2], material_id=0, light_id=-1))
    L = make_quad_light(0, (213.0, Y - 0.1, 227.0), (343.0, Y - 0.1, 227.0), (213.0, Y - 0.1, 332.0), (17.0, 12.0, 4.0), div_level, 0)
    sc.lights = L
    sc.meshes.append(light_mesh(L, 0))
    sc.camera = dict(eye=(278.0, 273.0, -800.0), lookat=(278.0, 273.0, 0.0), up=(0.0, 1.0, 0.0), fov=39.3)
    return sc


def _lcg_floats(seed, n):
    """the reference LCG (src/cuda/random.h:48-67), vectorised: n floats in [0,1)."""
    out = np.empty(n, np.float32)
    s = np.uint64(seed)
    for i in range(n):
        s = (np.uint64(1664525) * s + np.uint64(1013904223)) & np.uint64(0xffffffff)
        out[i] = np.float32(int(s) & 0x00ffffff) / np.float32(0x01000000)
    return out


def heightfield_scene(n=708, seed=1, octaves=6):
    """Traversal microbench mesh of config 2: deterministic fractal height
    field of n x n quads over [0,1]^2 (2 n^2 triangles; n=708 -> 1 002 528) inside a closed unit box
    (12 more triangles), LCG seed 1, plus a 1x1 quad light at the top facing down."""
    sc = SceneData()
    sc.name = "heightfield%d" % n
    r = _lcg_floats(seed, 4 * octaves)
    s = np.linspace(0.0, 1.0, n + 1, dtype=np.float64)
    Xg, Zg = np.meshgrid(s, s, indexing="xy")
    H = np.zeros_like(Xg)
    amp, freq = 0.12, 1.5
    for o in range(octaves):
        ph1, ph2, a1, a2 = (float(v) for v in r[4 * o:4 * o + 4])
        ang1, ang2 = 2 * np.pi * a1, 2 * np.pi * a2
        H += amp * np.sin(2 * np.pi * freq * (np.cos(ang1) * Xg + np.sin(ang1) * Zg) + 2 * np.pi * ph1) \
                 * np.cos(2 * np.pi * freq * 0.83 * (np.cos(ang2) * Xg + np.sin(ang2) * Zg) + 2 * np.pi * ph2)
        amp *= 0.55
        freq *= 2.03
    H = 0.3 + H
    pos = np.stack([Xg, H, Zg], -1).reshape(-1, 3).astype(np.float32)
    uv = np.stack([Xg, Zg], -1).reshape(-1, 2).astype(np.float32)
    i = np.arange(n)
    I, J = np.meshgrid(i, i, indexing="xy")
    a = (J * (n + 1) + I).reshape(-1)
    b, c = a + 1, a + (n + 1)
    d = c + 1
    idx = np.concatenate([np.stack([a, b, d], 1), np.stack([a, d, c], 1)], 0).astype(np.uint32)
    mats = make_pbr(2)
    mats["base_color"][0] = (0.6, 0.55, 0.5, 1)
    mats["base_color"][1] = (0.7, 0.7, 0.7, 1)
    sc.materials = mats
    sc.meshes.append(dict(positions=pos, indices=idx, texcoords=uv, material_id=0, light_id=-1))
    walls = _merge([
        _grid_quad((0, 0, 0), (1, 0, 0), (0, 0, 1), 1), _grid_quad((0, 1, 0), (1, 0, 0), (0, 0, 1), 1),
        _grid_quad((0, 0, 0), (1, 0, 0), (0, 1, 0), 1), _grid_quad((0, 0, 1), (1, 0, 0), (0, 1, 0), 1),
        _grid_quad((0, 0, 0), (0, 0, 1), (0, 1, 0), 1), _grid_quad((1, 0, 0), (0, 0, 1), (0, 1, 0), 1),
    ])
    sc.meshes.append(dict(positions=walls[0], indices=walls[1], texcoords=walls[2], material_id=1, light_id=-1))
    L = make_quad_light(0, (0.0, 0.999, 0.0), (1.0, 0.999, 0.0), (0.0, 0.999, 1.0), (5.0, 5.0, 5.0), 2, 0)
    sc.lights = L
    sc.meshes.append(light_mesh(L, 0))
    sc.camera = dict(eye=(0.5, 0.92, 0.02), lookat=(0.5, 0.3, 0.6), up=(0.0, 1.0, 0.0), fov=60.0)
    return sc


def random_soup_scene(n_tris=2000, seed=7, extent=10.0, n_lights=2):
    """Unstructured triangle soup with overlaps, degenerate and axis-aligned triangles: the
    adversarial parity fixture (ties, zero-area triangles, rays along box faces)."""
    rng = np.random.default_rng(seed)
    sc = SceneData()
    sc.name = "soup%d" % n_tris
    c = rng.uniform(-extent, extent, (n_tris, 1, 3))
    size = rng.choice([0.05, 0.5, 3.0], (n_tris, 1, 1))
    tri = (c + rng.normal(0, 1, (n_tris, 3, 3)) * size).astype(np.float32)
    k = n_tris // 20
    tri[:k, :, 1] = np.round(tri[:k, :1, 1])          # axis-aligned (flat in y)
    tri[k:2 * k, 2] = tri[k:2 * k, 1]                 # degenerate
    tri[2 * k:3 * k] = tri[3 * k:4 * k]               # exact duplicates -> equal-t ties
    pos = tri.reshape(-1, 3)
    idx = np.arange(3 * n_tris, dtype=np.uint32).reshape(-1, 3)
    half = n_tris // 2
    mats = make_pbr(2)
    mats["base_color"][1] = (0.2, 0.4, 0.9, 1)
    mats["metallic"][1] = 0.8
    mats["roughness"][1] = 0.2
    sc.materials = mats
    sc.meshes.append(dict(positions=pos[:3 * half], indices=idx[:half], texcoords=None, material_id=0, light_id=-1))
    sc.meshes.append(dict(positions=pos[3 * half:], indices=idx[half:] - np.uint32(3 * half), texcoords=None, material_id=1, light_id=-1))
    lights = []
    base = 0
    for li in range(n_lights):
        o = rng.uniform(-extent, extent, 3)
        L = make_quad_light(li, o, o + rng.normal(0, 2, 3), o + rng.normal(0, 2, 3), (10.0, 9.0, 8.0), 2, base)
        base += 4
        lights.append(L)
        sc.meshes.append(light_mesh(L, li))
    sc.lights = np.concatenate(lights)
    sc.camera = dict(eye=(0.0, 0.0, -3 * extent), lookat=(0.0, 0.0, 0.0), up=(0.0, 1.0, 0.0), fov=45.0)
    return sc


def camera_rays(scene, width, height, tmin=1e-3, tmax=1e16):
    """pinhole primaries at pixel centres (raygen.cu:335-344, subframe 0), numpy fp32."""
    from . import RAY
    eye, U, V, W = scene.camera_frame(width, height)
    f = np.float32
    x = (np.arange(width, dtype=f) + f(0.5)) / f(width)
    y = (np.arange(height, dtype=f) + f(0.5)) / f(height)
    dx = f(2) * x - f(1)
    dy = f(2) * y - f(1)
    D = dx[None, :, None] * U[None, None, :] + dy[:, None, None] * V[None, None, :] + W[None, None, :]
    D = (D / np.sqrt((D * D).sum(-1, keepdims=True))).astype(f).reshape(-1, 3)
    r = np.zeros(width * height, RAY)
    r["ox"], r["oy"], r["oz"] = eye
    r["dx"], r["dy"], r["dz"] = D[:, 0], D[:, 1], D[:, 2]
    r["tmin"] = tmin
    r["tmax"] = tmax
    return r


def random_rays(scene, n, seed=3, tmin=1e-3, tmax=1e16):
    """incoherent rays: origins uniform in the scene box, directions uniform on the sphere."""
    from . import RAY
    rng = np.random.default_rng(seed)
    P = np.concatenate([m["positions"] for m in scene.meshes])
    lo, hi = P.min(0), P.max(0)
    o = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    d = rng.normal(0, 1, (n, 3))
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    r = np.zeros(n, RAY)
    r["ox"], r["oy"], r["oz"] = o[:, 0], o[:, 1], o[:, 2]
    r["dx"], r["dy"], r["dz"] = d[:, 0], d[:, 1], d[:, 2]
    r["tmin"] = tmin
    r["tmax"] = tmax
    return r


# ------------------------------------------------------------------------------------------------------------
# on-disk formats: the reference's `.scene` + OBJ (sceneLoader.cpp, tiny_obj_loader.h) and this repo's binary
# `.spcscene` cache (host/host_scene.cpp).  The C++ host driver (host/spcbpt_main.cpp) reads both.
# ------------------------------------------------------------------------------------------------------------
def export_scene(scene, out_dir, name="scene"):
    """Write `scene` as <out_dir>/<name>/<name>.scene + one OBJ per mesh (+ PPM textures) in the reference's
    format, so that <out_dir> plays the role of SAMPLES_DIR/data.  Floats are printed with 9 significant digits,
    which round-trips binary32 through tinyobj's decimal parser.  Returns the path of the .scene file."""
    import os
    from . import LIGHT_QUAD
    d = os.path.join(out_dir, name)
    os.makedirs(os.path.join(d, "geometry"), exist_ok=True)
    os.makedirs(os.path.join(d, "textures"), exist_ok=True)
    g = lambda x: "%.9g" % float(x)   # noqa: E731
    lines = []
    cam = scene.camera
    lines += ["cameraSetting", "{", "    eye " + " ".join(g(v) for v in cam["eye"]), "    lookat " + " ".join(g(v) for v in cam["lookat"]),
              "    up " + " ".join(g(v) for v in cam["up"]), "    fov " + g(cam["fov"]), "    geo_normal 1", "}", ""]
    tex_files = []
    for i, t in enumerate(scene.textures):
        px = np.ascontiguousarray(t, np.uint8)
        rel = "%s/textures/tex%d.ppm" % (name, i)
        with open(os.path.join(out_dir, rel), "wb") as f:
            f.write(b"P6\n%d %d\n255\n" % (px.shape[1], px.shape[0]))
            f.write(px[:, :, :3].tobytes())
        tex_files.append(rel)
    for i in range(scene.materials.shape[0]):
        m = scene.materials[i]
        lines += ["material mat%d" % i, "{", "   color " + " ".join(g(v) for v in m["base_color"][:3]), "   roughness " + g(m["roughness"]),
                  "   metallic " + g(m["metallic"]), "   specular " + g(m["specular"])]
        tex = int(m["base_color_tex"]["tex"])
        if tex > 0:
            lines.append("   albedoTex " + tex_files[tex - 1])
        if int(m["brdf"]):
            lines.append("   brdf %d" % int(m["brdf"]))
        lines += ["}", ""]
    k = 0
    for m in scene.meshes:
        if m["light_id"] >= 0:
            continue
        rel = "%s/geometry/mesh%d.obj" % (name, k)
        with open(os.path.join(out_dir, rel), "w") as f:
            f.write("# exported by spcbpt-b200 scenes.export_scene\no mesh%d\n" % k)
            pos = np.asarray(m["positions"], np.float32)
            f.write("".join("v %.9g %.9g %.9g\n" % (p[0], p[1], p[2]) for p in pos.tolist()))
            uv = m.get("texcoords")
            if uv is not None:
                f.write("".join("vt %.9g %.9g\n" % (t[0], t[1]) for t in np.asarray(uv, np.float32).tolist()))
                f.write("".join("f %d/%d %d/%d %d/%d\n" % (a + 1, a + 1, b + 1, b + 1, c + 1, c + 1) for a, b, c in np.asarray(m["indices"]).tolist()))
            else:
                f.write("".join("f %d %d %d\n" % (a + 1, b + 1, c + 1) for a, b, c in np.asarray(m["indices"]).tolist()))
        # the reference writes Windows separators in its .scene files; keep one such line to exercise the path fix-up
        lines += ["mesh", "{", "    file " + (rel.replace("/", "\\") if k == 0 else rel), "    material mat%d" % int(m["material_id"]), "}", ""]
        k += 1
    for i in range(scene.lights.shape[0]):
        L = scene.lights[i]
        assert int(L["type"]) == LIGHT_QUAD
        lines += ["light", "{", "    position " + " ".join(g(v) for v in L["corner"]), "    v1 " + " ".join(g(v) for v in L["u"]),
                  "    v2 " + " ".join(g(v) for v in L["v"]), "    emission " + " ".join(g(v) for v in L["emission"]), "    type Quad",
                  "    divLevel %d" % int(L["divLevel"]), "}", ""]
    path = os.path.join(d, name + ".scene")
    with open(path, "w") as f:
        f.write("\n".join(lines))
    return path


def load_spcscene(path):
    """read a `.spcscene` cache written by host/host_scene.cpp save_scene_cache() (layout documented there)"""
    from . import LIGHT, PBR
    buf = np.fromfile(path, np.uint8)
    assert buf[:8].tobytes() == b"SPCSCN01", "not a .spcscene file"
    off = 8

    def take(dtype, count):
        nonlocal off
        dt = np.dtype(dtype)
        a = buf[off:off + dt.itemsize * count].view(dt)
        off += dt.itemsize * count
        return a
    nm, nmat, nl, nt = (int(v) for v in take("<u4", 4))
    cam = take("<f4", 10)
    sc = SceneData()
    sc.name = path
    sc.camera = dict(eye=tuple(cam[0:3]), lookat=tuple(cam[3:6]), up=tuple(cam[6:9]), fov=float(cam[9]))
    for _ in range(nm):
        nv, ntri = (int(v) for v in take("<u4", 2))
        mid, lid = (int(v) for v in take("<i4", 2))
        pos = take("<f4", 3 * nv).reshape(-1, 3).copy()
        idx = take("<u4", 3 * ntri).reshape(-1, 3).copy()
        uv = take("<f4", 2 * nv).reshape(-1, 2).copy()
        sc.meshes.append(dict(positions=pos, indices=idx, texcoords=uv, material_id=mid, light_id=lid))
    sc.materials = take(PBR, nmat).copy()
    sc.lights = take(LIGHT, nl).copy()
    for _ in range(nt):
        w, h = (int(v) for v in take("<i4", 2))
        sc.textures.append(take("u1", 4 * w * h).reshape(h, w, 4).copy())
    assert off == buf.shape[0], "trailing bytes in .spcscene"
    return sc


def large_scene(n=3160, emitters=16, seed=3):
    """Config 5 of BASELINE.json: large synthetic glossy scene -- a rough-metal fractal terrain of n x n quads
    (2 n^2 triangles: n = 3160 -> 19 971 200) with glossy ridges, inside a closed box, lit by emitters^2 small quad
    lights (16 x 16 = 256, divLevel 1 each -> 256 emitter subspaces: use K >= 1280, K_light = 256)."""
    sc = SceneData()
    sc.name = "large%d" % n
    r = _lcg_floats(seed, 32)
    s = np.linspace(0.0, 1.0, n + 1, dtype=np.float32)
    Xg, Zg = np.meshgrid(s, s, indexing="xy")
    H = np.zeros_like(Xg, dtype=np.float32)
    amp, freq = 0.10, 1.3
    for o in range(8):
        ph1, ph2, a1, a2 = (float(v) for v in r[4 * o:4 * o + 4])
        ang1, ang2 = 2 * np.pi * a1, 2 * np.pi * a2
        H += (amp * np.sin(2 * np.pi * freq * (np.cos(ang1) * Xg + np.sin(ang1) * Zg) + 2 * np.pi * ph1)
              * np.cos(2 * np.pi * freq * 0.83 * (np.cos(ang2) * Xg + np.sin(ang2) * Zg) + 2 * np.pi * ph2)).astype(np.float32)
        amp *= 0.55
        freq *= 2.03
    H = (0.3 + H).astype(np.float32)
    pos = np.stack([Xg, H, Zg], -1).reshape(-1, 3).astype(np.float32)
    uv = np.stack([Xg, Zg], -1).reshape(-1, 2).astype(np.float32)
    del Xg, Zg, H
    i = np.arange(n, dtype=np.int64)
    I, J = np.meshgrid(i, i, indexing="xy")
    a = (J * (n + 1) + I).reshape(-1)
    idx = np.empty((2 * n * n, 3), np.uint32)
    idx[:n * n, 0], idx[:n * n, 1], idx[:n * n, 2] = a, a + 1, a + n + 2
    idx[n * n:, 0], idx[n * n:, 1], idx[n * n:, 2] = a, a + n + 2, a + n + 1
    del a, I, J
    mats = make_pbr(2)
    mats["base_color"][0] = (0.9, 0.85, 0.8, 1)     # glossy metal terrain (caustic caster)
    mats["metallic"][0] = 1.0
    mats["roughness"][0] = 0.12
    mats["base_color"][1] = (0.7, 0.7, 0.72, 1)     # diffuse box
    sc.materials = mats
    sc.meshes.append(dict(positions=pos, indices=idx, texcoords=uv, material_id=0, light_id=-1))
    walls = _merge([
        _grid_quad((0, 0, 0), (1, 0, 0), (0, 0, 1), 1), _grid_quad((0, 1, 0), (1, 0, 0), (0, 0, 1), 1),
        _grid_quad((0, 0, 0), (1, 0, 0), (0, 1, 0), 1), _grid_quad((0, 0, 1), (1, 0, 0), (0, 1, 0), 1),
        _grid_quad((0, 0, 0), (0, 0, 1), (0, 1, 0), 1), _grid_quad((1, 0, 0), (0, 0, 1), (0, 1, 0), 1),
    ])
    sc.meshes.append(dict(positions=walls[0], indices=walls[1], texcoords=walls[2], material_id=1, light_id=-1))
    lights = []
    e = emitters
    for k in range(e * e):
        cx, cz = (k % e + 0.5) / e, (k // e + 0.5) / e
        hs = 0.15 / e
        y = 0.97
        # facing down: u x v must point to -y
        lights.append(make_quad_light(k, (cx - hs, y, cz - hs), (cx + hs, y, cz - hs), (cx - hs, y, cz + hs), (40.0, 36.0, 30.0), 1, k))
    sc.lights = np.concatenate(lights)
    for k in range(e * e):
        sc.meshes.append(light_mesh(sc.lights[k:k + 1], k))
    sc.camera = dict(eye=(0.5, 0.9, 0.03), lookat=(0.5, 0.3, 0.6), up=(0.0, 1.0, 0.0), fov=60.0)
    return sc
