"""Mutation fuzzing of the host-side file readers (host/: JPEG, PNG, OBJ, .scene, .spcscene) under AddressSanitizer + UBSan.
Texture, mesh and cache files are untrusted input of the C++ driver; this script is how the size / table / file-type checks in
host/jpeg_decode.cpp, png_decode.cpp, image_io.cpp and host_scene.cpp were found (tests/test_host_loader.py keeps one regression
case of each).  CPU only:   python tests/fuzz_host_loaders.py [--runs 60] [--seed 2]
A finding = the tool dies on a signal, or a sanitizer report appears on stderr; rejected files (exit code 1) are the expected outcome."""
import argparse
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "loader")
SRCS = ["scene_file.cpp", "host_scene.cpp", "image_io.cpp", "jpeg_decode.cpp", "png_decode.cpp", "train_state.cpp", "scene_tool.cpp"]


def mutate(raw, rng, text):
    b = bytearray(raw)
    k = int(rng.integers(0, 4))
    if k == 0:
        b = b[:int(rng.integers(1, len(b)))]
    elif k == 1:
        for _ in range(int(rng.integers(1, 8))):
            b[int(rng.integers(0, len(b)))] = int(rng.integers(32, 127) if text else rng.integers(0, 256))
    elif k == 2:
        i = int(rng.integers(0, len(b)))
        b[i:i] = bytes(rng.integers(0, 256, int(rng.integers(1, 40)), dtype=np.uint8))
    elif text:
        lines = bytes(b).split(b"\n")
        rng.shuffle(lines)
        b = bytearray(b"\n".join(lines))
    else:
        i = int(rng.integers(0, max(1, len(b) - 4)))
        b[i:i + 4] = bytes([255, 255, 255, int(rng.integers(0, 256))])
    return bytes(b)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--runs", type=int, default=60, help="mutations per input file")
    ap.add_argument("--seed", type=int, default=2)
    args = ap.parse_args()
    tmp = tempfile.mkdtemp(prefix="spc_fuzz_")
    tool = os.path.join(tmp, "scene_tool_asan")
    subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fwrapv", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-I" + os.path.join(ROOT, "include")]
                   + [os.path.join(ROOT, "host", s) for s in SRCS] + ["-o", tool, "-lz"], check=True)
    # a small scene of its own: .scene + OBJ + one texture, and its cache
    sc = os.path.join(tmp, "sc")
    os.makedirs(sc)
    with open(os.path.join(sc, "t.obj"), "w") as f:
        f.write("v 0 0 0\nv 1 0 0\nv 0 1 0\nv 1 1 0\nvt 0 0\nvt 1 0\nvt 0 1\nf 1/1 2/2 3/3\nf 2 4 3\n")
    with open(os.path.join(sc, "tex.ppm"), "wb") as f:
        f.write(b"P6\n2 2\n255\n" + bytes(range(12)))
    with open(os.path.join(sc, "s.scene"), "w") as f:
        f.write("cameraSetting\n{\n eye 0 0 -3\n lookat 0 0 0\n up 0 1 0\n fov 40\n}\nmaterial m\n{\n color 1 0.5 0.25\n roughness 0.5\n albedoTex tex.ppm\n}\n"
                "mesh\n{\n file t.obj\n material m\n}\nlight\n{\n type quad\n position 0 2 0\n u 1 0 0\n v 0 0 1\n emission 5 5 5\n}\n")
    subprocess.run([tool, "convert", os.path.join(sc, "s.scene"), os.path.join(sc, "s.spcscene"), "--data-root", sc], check=True, capture_output=True)
    jobs = [(os.path.join(GOLD, f), "decode", False) for f in sorted(os.listdir(GOLD)) if f.endswith((".jpg", ".png"))]
    jobs += [(os.path.join(GOLD, "quirks.obj"), "obj", True), (os.path.join(sc, "s.scene"), "scene", True), (os.path.join(sc, "s.scene"), "convert", True),
             (os.path.join(sc, "s.spcscene"), "info", False)]
    # trained-state checkpoint files (host/train_state.cpp): a two-node state, one file mutated at a time
    st = os.path.join(tmp, "state")
    os.makedirs(os.path.join(st, "out"))
    for name, text_ in (("tree_eye.txt", "0 2 1 0.5 0.25 0.125 1 2 3 4 5 6 7 8\n" + "".join("1 %d\n" % i for i in range(8))), ("tree_light.txt", "1 0\n"),
                        ("Q.txt", "0.5\n0.25\n0.125\n1\n"), ("E.txt", "0.25 0.25 0.25 0.25\n" * 4)):
        with open(os.path.join(st, name), "w") as f:
            f.write(text_)
    jobs += [(os.path.join(st, n), "state", True) for n in ("tree_eye.txt", "Q.txt", "E.txt")]
    rng = np.random.default_rng(args.seed)
    runs = findings = 0
    for path, cmd, text in jobs:
        raw = open(path, "rb").read()
        victim = path if cmd == "state" else os.path.join(sc if cmd in ("scene", "convert") else tmp, "fz" + os.path.splitext(path)[1])
        argv = [tool, "state", st + "/", os.path.join(st, "out") + "/", "4"] if cmd == "state" else [tool, cmd, victim, os.path.join(tmp, "out")]
        for _ in range(args.runs):
            with open(victim, "wb") as f:
                f.write(mutate(raw, rng, text))
            try:
                r = subprocess.run(argv, capture_output=True, timeout=120)
                err = r.stderr.decode(errors="replace")
                bad = r.returncode not in (0, 1, 2) or "Sanitizer" in err or "runtime error" in err
            except subprocess.TimeoutExpired:
                err, bad = "timeout", True
            runs += 1
            if bad:
                findings += 1
                keep = os.path.join(tmp, "finding_%d%s" % (findings, os.path.splitext(path)[1]))
                with open(keep, "wb") as f:
                    f.write(open(victim, "rb").read())
                print("FINDING", cmd, os.path.basename(path), "->", keep, "\n", err[-1500:])
        if cmd == "state":
            with open(path, "wb") as f:
                f.write(raw)
    print("runs %d findings %d (work dir %s)" % (runs, findings, tmp))
    return 1 if findings else 0


if __name__ == "__main__":
    sys.exit(main())
