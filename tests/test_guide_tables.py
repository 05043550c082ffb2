"""CPU property tests of two host-checkable claims behind k_eye_sample (csrc/render.cu, csrc/shade.cuh):

1. guide tables ("cutpoint method"): a search bracketed by the guide table returns the SAME index as the reference's own bisect
   (binary_sample, cuProg.h:245-264) for every u, on the kinds of tables the library builds them for -- non-decreasing running
   sums divided by their total (with plateaus from zero weights), CDF rows whose last entry is forced to 1, and all-NaN tables of
   subspaces whose weights sum to zero;
2. candidate positions: resolving the first-stage draw of connection j among the 2C-1 candidate stream positions reproduces the
   serial loop of raygen.cu:390-408 (an empty subspace consumes one draw, a non-empty one two) for every pattern of empties.

The numpy models below restate the device code operation by operation in float32 (IEEE division for the cuts, the float32 product
u*n for the first cell guess); the device code itself is covered bit-for-bit by the GPU frame-parity tests against the oracle.
"""
import numpy as np

f32 = np.float32


def ref_bisect(cmf, u):
    """binary_sample (cuProg.h:245-264): mid = size/2-1 ... (l+r)/2-1, returns l"""
    size = len(cmf)
    mid, l, r = size // 2 - 1, 0, size
    while r - l > 1:
        if u < cmf[mid]:
            r = mid + 1
        else:
            l = mid + 1
        mid = (l + r) // 2 - 1
    return l


def guide_build(cmf):
    """guide_build_table (shade.cuh): G[j] = first i with cmf[i] > j/n, else n; j = 0..n"""
    n = len(cmf)
    fn = f32(n)
    G = np.empty(n + 1, np.int64)
    for j in range(n + 1):
        c = f32(j) / fn
        lo, hi = 0, n
        while lo < hi:
            mid = (lo + hi) >> 1
            if cmf[mid] > c:
                hi = mid
            else:
                lo = mid + 1
        G[j] = lo
    return G


def guide_cell(u, n):
    fn = f32(n)
    j = min(int(f32(u) * fn), n - 1)
    while j > 0 and u < f32(j) / fn:
        j -= 1
    while j < n - 1 and not (u < f32(j + 1) / fn):
        j += 1
    return j


def guided_search(cmf, G, u):
    """guided_sample (shade.cuh)"""
    n = len(cmf)
    cell = guide_cell(u, n)
    hi = min(int(G[cell + 1]), n - 1)
    lo = min(int(G[cell]), hi)
    probes = 0
    while lo < hi:
        mid = (lo + hi) >> 1
        probes += 1
        if u < cmf[mid]:
            hi = mid
        else:
            lo = mid + 1
    return lo, probes


def lcg_u(rng, k):
    """the values rnd() can return: multiples of 2^-24 in [0, 1)"""
    return (rng.integers(0, 1 << 24, k).astype(np.float32) / f32(16777216.0)).astype(np.float32)


def running_cdf(w):
    """k_lvc_cmf: running fp32 sum in order, then division by the total"""
    run = np.empty(len(w), np.float32)
    acc = f32(0)
    for i, x in enumerate(w):
        acc = f32(x) if i == 0 else f32(x + acc)
        run[i] = acc
    with np.errstate(invalid="ignore", divide="ignore"):
        return (run / run[-1]).astype(np.float32)


def check_table(cmf, rng, n_u=400):
    G = guide_build(cmf)
    us = np.concatenate([lcg_u(rng, n_u), np.array([0.0, 1.0 - 2.0 ** -24], np.float32)])
    # also every value adjacent to a table entry or a cut (the places where an off-by-one would show)
    near = np.concatenate([cmf, (np.arange(len(cmf) + 1, dtype=np.float32) / f32(len(cmf))).astype(np.float32)])
    near = near[np.isfinite(near)]
    near = np.concatenate([near, np.nextafter(near, f32(0)), np.nextafter(near, f32(2))])
    near = near[(near >= 0) & (near < 1)]
    total_probes = 0
    for u in np.concatenate([us, near.astype(np.float32)]):
        want = ref_bisect(cmf, u)
        got, probes = guided_search(cmf, G, u)
        assert got == want, (len(cmf), float(u), got, want)
        total_probes += probes
    return total_probes / (len(us) + len(near))


def test_guided_search_equals_reference_bisect_on_running_sum_tables():
    rng = np.random.default_rng(7)
    avg = []
    for n in [1, 2, 3, 5, 8, 31, 64, 257, 1000, 4096]:
        for kind in range(4):
            if kind == 0:
                w = rng.random(n)
            elif kind == 1:   # heavy tail: a few dominant entries, many tiny ones
                w = rng.random(n) ** 12
            elif kind == 2:   # plateaus: most weights exactly zero
                w = np.where(rng.random(n) < 0.7, 0.0, rng.random(n))
                if not w.any():
                    w[rng.integers(n)] = 1.0
            else:             # first and last entries zero
                w = rng.random(n)
                w[0] = 0.0
                w[-1] = 0.0
                if not w.any():
                    w[n // 2] = 1.0
            avg.append(check_table(running_cdf(w.astype(np.float32)), rng, n_u=150 if n > 500 else 300))
    # the point of the tables: a search is a couple of probes, not log2(n)
    assert np.mean(avg) < 2.0


def test_guided_search_on_cdf_rows_with_forced_last_entry():
    """Gamma2CMFGamma rows: running sum of 0.8*E + 0.2/K (positive), last entry forced to 1 -- the entry before it may exceed 1"""
    rng = np.random.default_rng(11)
    for K in [12, 64, 1000]:
        for _ in range(3):
            E = rng.random(K).astype(np.float32) ** 4
            E = (E / E.sum()).astype(np.float32)
            row = np.cumsum((f32(0.8) * E + f32(0.2) / f32(K)).astype(np.float32), dtype=np.float32)
            row[-1] = f32(1.0)
            check_table(row, rng)
        row = np.cumsum(np.full(K, 1.0 / K, np.float32), dtype=np.float32)
        row[-2:] = [f32(1.0000001), f32(1.0)] if K > 1 else row[-2:]
        check_table(row.astype(np.float32), rng)


def test_guided_search_on_all_nan_table():
    """a subspace whose weights sum to zero: 0/0 everywhere; the reference's bisect ends at the last entry"""
    rng = np.random.default_rng(3)
    for n in [1, 2, 7, 100]:
        cmf = np.full(n, np.nan, np.float32)
        G = guide_build(cmf)
        assert (G == n).all()
        for u in lcg_u(rng, 50):
            assert guided_search(cmf, G, u)[0] == ref_bisect(cmf, u) == n - 1


def serial_connections(empty_of_draw, C):
    """raygen.cu:390-408 reduced to its draw bookkeeping: returns [(stage-1 position, stage-2 position or None)] and the draws used"""
    pos, out = 0, []
    for _ in range(C):
        if empty_of_draw[pos]:
            out.append((pos, None))
            pos += 1
        else:
            out.append((pos, pos + 1))
            pos += 2
    return out, pos


def lockstep_connections(empty_of_draw, C):
    """eye_sample_lockstep (render.cu): candidates 0..2C-2, resolution in registers"""
    pos, out = 0, []
    for j in range(C):
        hit = None
        for c in range(j, 2 * j + 1):
            if c == pos:
                hit = c
        assert hit is not None
        if empty_of_draw[hit]:
            out.append((hit, None))
            pos += 1
        else:
            out.append((hit, hit + 1))
            pos += 2
    return out, pos


def test_candidate_positions_reproduce_the_serial_draw_order():
    for C in [1, 2, 3, 4]:
        n_draws = 2 * C
        for pattern in range(1 << n_draws):
            empty = [(pattern >> k) & 1 == 1 for k in range(n_draws)]
            assert lockstep_connections(empty, C) == serial_connections(empty, C)
            assert C <= serial_connections(empty, C)[1] <= 2 * C
