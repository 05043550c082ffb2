"""Generates tests/golden/loader/*: inputs for the host scene-ingest code (host/scene_file.cpp, jpeg_decode.cpp,
png_decode.cpp) and the outputs of the REFERENCE's own loaders on them -- tinyobjloader, stb_image and LoadScene
compiled from /root/reference into oracle/_ref/ref_loader (oracle/ref_shim/ref_loader.cpp).  Run in the authoring
container only (needs /root/reference and PIL):  python tests/golden/make_golden_loader.py
"""
import hashlib
import io
import json
import os
import subprocess
import sys

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, "loader")
REF = os.path.join(ROOT, "oracle", "_ref", "ref_loader")
HOUSE = "/root/reference/src/data/house"

QUIRKS_OBJ = """# OBJ corner cases: relative indices, polygons, v//vn, v/vt, v/vt/vn, g / o / usemtl splits, exponents, CRLF
mtllib does_not_exist.mtl
v 0 0 0
v 1 0 0
v 1 1 0
v 0 1 0
v 0.5 0.5 1e-1
v -1.25E+1 3.14159265358979 .5
v +2.5e-3 -0.000123456789 12345678.9
vt 0 0
vt 1 0
vt 1 1
vt 0.25 0.75
vn 0 0 1
vn 0 1 0
f 1 2 3 4
f -1 -2 -3
o second_object
f 1/1 2/2 3/3
f 1/1 3/3 4/4 5/1
g group_a extra_name
f 1//1 2//1 3//2
usemtl whatever
f 1/1/1 2/2/1 3/3/2 4/4/2 5/1/1 6/2/2
f 5 6 7
g
f 7/4 6/3 5/2\r
f 1 2
v 9 9 9
f -1 1 2
"""


def ref(*args):
    subprocess.run([REF] + list(args), check=True)


def sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def main():
    assert os.path.exists(REF), "build oracle/_ref/ref_loader first (make -C oracle ref)"
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "quirks.obj"), "w", newline="") as f:
        f.write(QUIRKS_OBJ)
    ref("obj", os.path.join(OUT, "quirks.obj"), os.path.join(OUT, "quirks.bin"))

    # small synthetic images: smooth gradient + noise so that every DCT coefficient and chroma tap is exercised
    rng = np.random.default_rng(5)

    rng_prog = np.random.default_rng(6)    # (own stream for the progressive cases: the earlier fixtures keep their bytes)

    def picture(w, h, chan, rng=rng):
        y, x = np.mgrid[0:h, 0:w]
        base = np.stack([(x * 255 // max(w - 1, 1)), (y * 255 // max(h - 1, 1)), ((x + y) * 255 // max(w + h - 2, 1))], -1)
        img = np.clip(base + rng.integers(-40, 40, (h, w, 3)), 0, 255).astype(np.uint8)
        return img if chan == 3 else img[:, :, 0]
    golden = {}
    cases = []
    for name, w, h, mode, kw in [
        ("j444", 37, 29, "RGB", dict(subsampling=0, quality=90)),
        ("j422", 50, 19, "RGB", dict(subsampling=1, quality=75)),
        ("j420", 61, 43, "RGB", dict(subsampling=2, quality=60)),
        ("j420_odd", 17, 9, "RGB", dict(subsampling=2, quality=95)),
        ("j420_1px", 1, 1, "RGB", dict(subsampling=2, quality=95)),
        ("jgray", 33, 20, "L", dict(quality=80)),
        ("j420_restart", 64, 48, "RGB", dict(subsampling=2, quality=70, restart_marker_blocks=3)),
        ("j420_opt", 40, 40, "RGB", dict(subsampling=2, quality=30, optimize=True)),
        # progressive (SOF2): spectral selection + successive approximation, DC-only interleaved scans, end-of-band runs, restarts
        ("jprog444", 37, 29, "RGB", dict(subsampling=0, quality=90, progressive=True)),
        ("jprog420", 61, 43, "RGB", dict(subsampling=2, quality=60, progressive=True)),
        ("jprog420_lowq", 96, 80, "RGB", dict(subsampling=2, quality=8, progressive=True)),
        ("jprog422_restart", 50, 35, "RGB", dict(subsampling=1, quality=75, progressive=True, restart_marker_blocks=2)),
        ("jprog_gray", 33, 20, "L", dict(quality=80, progressive=True)),
        ("jprog_1px", 1, 1, "RGB", dict(subsampling=2, quality=95, progressive=True)),
    ]:
        im = Image.fromarray(picture(w, h, 3 if mode == "RGB" else 1, rng_prog if name.startswith("jprog") else rng), mode)
        p = os.path.join(OUT, name + ".jpg")
        im.save(p, "JPEG", **kw)
        cases.append(name + ".jpg")
    for name, w, h, mode, kw in [
        ("p_rgb", 23, 17, "RGB", {}),
        ("p_rgba", 16, 16, "RGBA", {}),
        ("p_gray", 19, 7, "L", {}),
        ("p_la", 9, 11, "LA", {}),
        ("p_pal", 32, 8, "P", {}),
        ("p_1bit", 21, 5, "1", {}),
        ("p_rgb16", 8, 6, "I;16", {}),
    ]:
        a = picture(w, h, 3)
        if mode == "RGB":
            im = Image.fromarray(a, "RGB")
        elif mode == "RGBA":
            im = Image.fromarray(np.concatenate([a, a[:, :, :1]], -1), "RGBA")
        elif mode == "L":
            im = Image.fromarray(a[:, :, 0], "L")
        elif mode == "LA":
            im = Image.fromarray(np.stack([a[:, :, 0], a[:, :, 1]], -1), "LA")
        elif mode == "P":
            im = Image.fromarray(a, "RGB").quantize(16)
        elif mode == "1":
            im = Image.fromarray(a[:, :, 0] > 128)
        else:
            im = Image.fromarray((a[:, :, 0].astype(np.uint16) * 257))
        p = os.path.join(OUT, name + ".png")
        im.save(p, "PNG", **kw)
        cases.append(name + ".png")
    for c in cases:
        tmp = os.path.join(OUT, "_tmp.rgba8")
        ref("decode", os.path.join(OUT, c), tmp)
        raw = np.fromfile(tmp, np.uint8)
        os.remove(tmp)
        golden[c] = raw
    np.savez_compressed(os.path.join(OUT, "stb_decodes.npz"), **golden)

    # digests of the reference loaders on the shipped scene (the data itself stays in /root/reference)
    dig = {"scene": None, "obj": {}, "tex": {}}
    tmp = os.path.join(OUT, "_tmp.bin")
    ref("scene", os.path.join(HOUSE, "house_uvrefine2.scene"), tmp)
    dig["scene"] = hashlib.sha256(open(tmp, "rb").read().replace(b"\\", b"/")).hexdigest()
    for f in sorted(os.listdir(os.path.join(HOUSE, "geometry"))):
        if f.endswith(".obj"):
            ref("obj", os.path.join(HOUSE, "geometry", f), tmp)
            dig["obj"][f] = sha(tmp)
    for f in sorted(os.listdir(os.path.join(HOUSE, "textures"))):
        if f.lower().endswith((".jpg", ".png")):
            ref("decode", os.path.join(HOUSE, "textures", f), tmp)
            dig["tex"][f] = sha(tmp)
    os.remove(tmp)
    dig["triangles"] = 119140
    json.dump(dig, open(os.path.join(OUT, "house_digest.json"), "w"), indent=1, sort_keys=True)
    print("wrote", OUT, len(cases), "images,", len(dig["obj"]), "house meshes,", len(dig["tex"]), "house textures")


if __name__ == "__main__":
    sys.exit(main())
