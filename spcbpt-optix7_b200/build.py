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


# The fast flavour (libspcbpt_b200_fast.so): the shading kernels as the reference's own build compiles its programs
# (src/CMakeLists.txt:214-215 --use_fast_math): FMA contraction, approximate division / square root, hardware special-function
# intrinsics (shade.cuh SPC_FAST_MATH).  -ftz stays off (explicit-intrinsic contract arithmetic of the traversal must not flush) and
# only the three shading translation units change: traversal batches, primary rays, BVH build, LVC binning and training are the
# exact objects in both libraries.  The exact library stays the default and the one all parity tests run.
OUT_FAST = os.path.join(HERE, "libspcbpt_b200_fast.so")
FAST_SOURCES = ("render.cu", "pt.cu", "pretrace.cu")
FAST_SWAP = {"-prec-div=true": "-prec-div=false", "-prec-sqrt=true": "-prec-sqrt=false", "-fmad=false": "-fmad=true"}
NVCC_FLAGS_FAST = [FAST_SWAP.get(f, f) for f in NVCC_FLAGS] + ["-DSPC_FAST_MATH=1"]


LINK_LIBS = ["-ldl"]   # comm.cu binds libnccl.so.2 at the first spc_comm_* call (types from /usr/include/nccl.h)


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
    cflags_fast = [f for f in NVCC_FLAGS_FAST if f != "--shared"]
    jobs, objs, objs_fast, logs = [], [], [], {}
    for src in sources():
        base = os.path.splitext(os.path.basename(src))[0]
        variants = [("", cflags)] + ([("_fast", cflags_fast)] if os.path.basename(src) in FAST_SOURCES else [])
        for suffix, flags in variants:
            h = hashlib.sha256(hdr.digest())
            h.update(" ".join(flags).encode())
            with open(src, "rb") as fh:
                h.update(fh.read())
            obj = os.path.join(objdir, base + suffix + ".o")
            stamp = os.path.join(objdir, base + suffix + ".stamp")
            log = os.path.join(objdir, base + suffix + ".log")
            if suffix:
                objs_fast.append(obj)
            else:
                objs.append(obj)
                if os.path.basename(src) not in FAST_SOURCES:
                    objs_fast.append(obj)
            if force or not (os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read().strip() == h.hexdigest()):
                jobs.append((src, obj, stamp, log, h.hexdigest(), flags))

    def compile_one(job):
        src, obj, stamp, log, dig, flags = job
        cmd = [nvcc] + flags + ["-c", src, "-o", obj]
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
    for out, olist in ((OUT, objs), (OUT_FAST, objs_fast)):
        if jobs or not os.path.exists(out):
            cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-Xcompiler", "-fPIC"] + olist + ["-o", out, "-lcudart"] + LINK_LIBS
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError("link of %s failed" % os.path.basename(out))
    with open(os.path.join(HERE, "ptxas_info.txt"), "w") as fh:
        for obj in sorted(set(objs + objs_fast)):
            fh.write(open(os.path.splitext(obj)[0] + ".log").read())
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
