"""In-tree build of libspcbpt_b200.so (nvcc, sm_100a only) -- no JIT cache, the .so travels with the tree.

Usage:  python spcbpt-optix7_b200/build.py [--force] [--verbose]
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libspcbpt_b200.so")
STAMP = os.path.join(HERE, ".build_stamp")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "--shared",
    "-Xcompiler", "-fPIC",
    # parity: IEEE div/sqrt, denormals kept, no implicit a*b+c contraction (-fmad=false) so that
    # plain fp32 expressions round exactly like the host oracle built with -ffp-contract=off; fused
    # ops appear only where written (explicit __fmaf_rn in the traversal and the contract
    # arithmetic).  Never --use_fast_math.
    "-prec-div=true", "-prec-sqrt=true", "-ftz=false", "-fmad=false",
    "-Xptxas", "-v",
    "-diag-suppress", "177,550",
]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cu", ".cuh", ".h")):
                h.update(f.encode())
                with open(os.path.join(root, f), "rb") as fh:
                    h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    """compile every csrc/*.cu to an object (in parallel, cached per source+headers+flags digest), then link"""
    from concurrent.futures import ThreadPoolExecutor
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    hdr = hashlib.sha256()
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cuh", ".h")):
                hdr.update(f.encode())
                with open(os.path.join(root, f), "rb") as fh:
                    hdr.update(fh.read())
    hdr.update(" ".join(NVCC_FLAGS).encode())
    cflags = [f for f in NVCC_FLAGS if f != "--shared"]
    jobs, objs, logs = [], [], {}
    for src in sources():
        h = hashlib.sha256(hdr.digest())
        with open(src, "rb") as fh:
            h.update(fh.read())
        base = os.path.splitext(os.path.basename(src))[0]
        obj = os.path.join(objdir, base + ".o")
        stamp = os.path.join(objdir, base + ".stamp")
        log = os.path.join(objdir, base + ".log")
        objs.append(obj)
        if force or not (os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read().strip() == h.hexdigest()):
            jobs.append((src, obj, stamp, log, h.hexdigest()))

    def compile_one(job):
        src, obj, stamp, log, dig = job
        cmd = [nvcc] + cflags + ["-c", src, "-o", obj]
        if verbose:
            print(" ".join(cmd))
        r = subprocess.run(cmd, capture_output=True, text=True)
        with open(log, "w") as fh:
            fh.write(r.stdout + r.stderr)
        if r.returncode == 0:
            with open(stamp, "w") as fh:
                fh.write(dig)
        return src, r

    failed = False
    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for src, r in ex.map(compile_one, jobs):
                if verbose or r.returncode != 0:
                    sys.stderr.write(r.stdout + r.stderr)
                failed |= r.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed building libspcbpt_b200.so")
    if jobs or not os.path.exists(OUT):
        cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-Xcompiler", "-fPIC"] + objs + ["-o", OUT, "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link of libspcbpt_b200.so failed")
    with open(os.path.join(HERE, "ptxas_info.txt"), "w") as fh:
        for obj in objs:
            fh.write(open(os.path.splitext(obj)[0] + ".log").read())
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
