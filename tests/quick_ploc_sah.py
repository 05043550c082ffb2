"""CPU study (numpy): binary-BVH SAH cost of the PLOC topology (csrc/bvh_build.cu step 3b) as a function of the search radius, next to
the Morton radix tree's, on the shipped house scene (data/_ref/house.spcscene) or a synthetic scene.  A planning aid for the builder
(which radius is worth its build time), not a test: python tests/quick_ploc_sah.py [radii...]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spcbpt_loader  # noqa: E402


def half_area(lo, hi):
    d = hi - lo
    return d[..., 0] * d[..., 1] + d[..., 1] * d[..., 2] + d[..., 2] * d[..., 0]


def morton_order(lo, hi):
    c = 0.5 * (lo + hi)
    mn, mx = c.min(0), c.max(0)
    q = np.clip(((c - mn) / np.maximum(mx - mn, 1e-30) * 2097152.0), 0, 2097151).astype(np.uint64)

    def spread(v):
        x = v & np.uint64(0x1fffff)
        x = (x | x << np.uint64(32)) & np.uint64(0x1f00000000ffff)
        x = (x | x << np.uint64(16)) & np.uint64(0x1f0000ff0000ff)
        x = (x | x << np.uint64(8)) & np.uint64(0x100f00f00f00f00f)
        x = (x | x << np.uint64(4)) & np.uint64(0x10c30c30c30c30c3)
        x = (x | x << np.uint64(2)) & np.uint64(0x1249249249249249)
        return x
    key = (spread(q[:, 0]) << np.uint64(2)) | (spread(q[:, 1]) << np.uint64(1)) | spread(q[:, 2])
    return np.argsort(key, kind="stable"), key


def ploc_sah(lo, hi, radius):
    """sum of the half areas of all internal nodes of the PLOC tree (the node-visit term of the SAH), rounds"""
    lo, hi = lo.copy(), hi.copy()
    total, rounds = 0.0, 0
    while len(lo) > 1:
        m = len(lo)
        best = np.full(m, np.inf)
        nn = np.full(m, -1, np.int64)
        for d in range(1, min(radius, m - 1) + 1):
            a = half_area(np.minimum(lo[:-d], lo[d:]), np.maximum(hi[:-d], hi[d:]))
            # candidate j = i + d for i, and j = i - d for i + d; ties keep the lower index (ascending-j scan of the kernel)
            fwd = a < best[:-d]
            idx = np.nonzero(fwd)[0]
            best[idx] = a[idx]
            nn[idx] = idx + d
            bwd = a <= best[d:]          # j = i - d is lower than anything seen so far for i + d at this d? scan order is ascending j
            # emulate ascending-j order: backward neighbours (lower j) must win ties against forward ones found at smaller d
            idx = np.nonzero((a < best[d:]) | ((a == best[d:]) & (nn[d:] > np.arange(m - d))))[0]
            best[idx + d] = a[idx]
            nn[idx + d] = idx
        i = np.arange(m)
        mutual = nn[nn] == i
        lower = mutual & (i < nn)
        gone = mutual & (i > nn)
        li = np.nonzero(lower)[0]
        lo[li] = np.minimum(lo[li], lo[nn[li]])
        hi[li] = np.maximum(hi[li], hi[nn[li]])
        total += float(half_area(lo[li], hi[li]).sum())
        keep = ~gone
        lo, hi = lo[keep], hi[keep]
        rounds += 1
    return total, rounds


def lbvh_sah(lo, hi, key):
    """same term for the Morton radix tree (recursive split at the highest differing bit)"""
    total = 0.0
    stack = [(0, len(lo) - 1)]
    # prefix boxes are not enough for arbitrary ranges: accumulate bottom-up with an explicit post-order
    order = []
    while stack:
        a, b = stack.pop()
        if a == b:
            continue
        x = int(key[a]) ^ int(key[b])
        if x == 0:
            s = (a + b) >> 1
        else:
            bit = x.bit_length() - 1
            # first index in (a, b] whose key has that bit set
            lo_i, hi_i = a, b
            while lo_i + 1 < hi_i:
                mid = (lo_i + hi_i) >> 1
                if (int(key[mid]) >> bit) & 1 == (int(key[a]) >> bit) & 1:
                    lo_i = mid
                else:
                    hi_i = mid
            s = lo_i
        order.append((a, b))
        stack.append((a, s))
        stack.append((s + 1, b))
    for a, b in order:
        total += float(half_area(lo[a:b + 1].min(0), hi[a:b + 1].max(0)))
    return total


def main():
    radii = [int(x) for x in sys.argv[1:]] or [2, 4, 8, 16, 32, 64]
    pkg = spcbpt_loader.load()
    cache = os.path.join(ROOT, "data", "_ref", "house.spcscene")
    if os.path.exists(cache):
        sc, name = pkg.scenes.load_spcscene(cache), "house"
    else:
        sc, name = pkg.scenes.cornell_scene(wall_cells=72, box_cells=60), "cornell"
    tri = np.concatenate([m["positions"][m["indices"].reshape(-1)].reshape(-1, 3, 3) for m in sc.meshes]).astype(np.float64)
    lo, hi = tri.min(1), tri.max(1)
    order, key = morton_order(lo, hi)
    lo, hi, key = lo[order], hi[order], key[order]
    root = float(half_area(lo.min(0), hi.max(0)))
    print("%s: %d triangles (quad lights not included)" % (name, len(lo)))
    if len(lo) <= 300000:
        t0 = time.time()
        print("  radix tree      inner-node area / root area = %8.2f   (%.1f s)" % (lbvh_sah(lo, hi, key) / root, time.time() - t0))
    for r in radii:
        t0 = time.time()
        tot, rounds = ploc_sah(lo, hi, r)
        print("  PLOC radius %3d  inner-node area / root area = %8.2f   %3d rounds (%.1f s)" % (r, tot / root, rounds, time.time() - t0))


if __name__ == "__main__":
    main()
