"""ORACLE (test infrastructure, never on the product path) — expected values (``mpmean``)
and the sampled background distribution (``nulldist``) restated.

Follows reference src/stripenn/getStripe.py:178-235 (mpmean) and :238-499 (nulldist:
sample budget :242-247,278-283; available_cols :249-273; main_null_calc :285-479 including
the top-up block :416-477 that re-uses the stale loop variable ``it``).

RNG convention: a fresh ``random.Random(seed)`` per chromosome task — the
behaviour of the reference with ``-n >= 2`` (every loky task unpickles the unadvanced
generator).
"""
import math
import random

import numpy as np

from .matrix import nan0

NDIST = 400


def expected_by_distance(mat, chroms):
    """mpmean: ex[d] = sum_i nan0(M[i, i+d]) / (N - d), d = 0..399 (NaN counted as 0 but
    still counted, getStripe.py:204-207)."""
    out = {}
    for c in chroms:
        n = mat.nbins(c)
        sums = np.zeros(NDIST)
        # 400-row frames exactly like calc(n) (getStripe.py:184-193)
        for f in range(math.ceil(n / NDIST)):
            r0 = f * NDIST
            r1 = min(r0 + NDIST, n)
            c1 = min(r0 + 2 * NDIST, n)
            w = nan0(mat.window(c, r0, r1, r0, c1))
            rows, cols = w.shape
            for d in range(min(NDIST, cols)):
                k = min(rows, cols - d)
                if k > 0:
                    sums[d] += w[np.arange(k), np.arange(k) + d].sum()
        with np.errstate(divide="ignore", invalid="ignore"):
            out[c] = (sums / (n - np.arange(NDIST))).tolist()
    return out


def _budget(values):
    v = (np.asarray(values, dtype=np.float64) / np.sum(values)) * 1000
    v = v.astype(np.uint64)
    dif = 1000 - int(np.sum(v))
    nz = np.nonzero(v != 0)[0]
    v[0] += np.uint64(dif)
    return v, nz


def _unit_geometry(mat, chrom, it, unit, r):
    """Row/column bin ranges of the window main_null_calc fetches for unit ``it``
    (getStripe.py:313-326 / :419-432)."""
    csize = mat.chromsize(chrom)
    start1 = int(unit * r * it + 1)
    start0 = start1 - 400 * r
    end1 = int(unit * r * (it + 1))
    end2 = int(unit * r * (it + 1) + 400 * r)
    if end2 > csize:
        end2 = csize - 1
    if end1 > csize - 400 * r:
        end1 = csize - 400 * r
    if start0 <= 1:
        start0 = 1
    i0, i1 = mat.bins_of(chrom, start1, end1)
    j0, j1 = mat.bins_of(chrom, start0, end2)
    return i0, i1, j0, j1


def available_rows(mat, chrom):
    """available_cols (getStripe.py:249-273): rows with a non-zero sum inside each unit's own
    square block."""
    r = mat.binsize
    csize = mat.chromsize(chrom)
    itera = int(min(csize / r / 500, 25))
    if itera == 0:
        # fewer than 500 bins: the reference divides by np.uint64(0) (-> inf, silenced by errstate)
        # and loops over range(0): poolsum = 0, the chromosome drops out of the sample budget
        return 0
    unit = int(np.floor(csize / r / itera))
    tot = 0
    for it in range(itera):
        i0, i1 = mat.bins_of(chrom, unit * r * it + 1, unit * r * (it + 1))
        s = nan0(mat.window(chrom, i0, i1, i0, i1)).sum(axis=1)
        tot += int((s != 0).sum())
    return tot


def plan_samples(mat, all_names, all_sizes, seed):
    """Everything of nulldist that decides WHICH rows are sampled.  Returns
    (chromnames2, per-chromosome list of sample specs).  A spec is
    ``(i0, i1, j0, j1, x, shift)``: fetched window extent in bins, local row, and the +400
    column offset actually applied."""
    r = mat.binsize
    v1, nz1 = _budget(all_sizes)
    names2 = [all_names[i] for i in nz1]
    avail = [available_rows(mat, c) for c in names2]
    budget, nz2 = _budget(avail)
    names2 = [names2[i] for i in nz2]
    plans = []
    for chrom in names2:
        prng = random.Random(seed)
        ss = int(budget[names2.index(chrom)])  # index into the UNfiltered budget (quirk)
        csize = mat.chromsize(chrom)
        itera = int(min(csize / r / 500, 25))
        unit = int(np.floor(csize / r / itera))
        specs, n_pool = [], []
        it = 0
        for it in range(itera):
            sss = int(ss / itera)
            i0, i1, j0, j1 = _unit_geometry(mat, chrom, it, unit, r)
            w = nan0(mat.window(chrom, i0, i1, j0, j1))
            live = w.sum(axis=1) != 0
            pool = [x for x in range(w.shape[0]) if live[x] and 20 < x < unit - 20]
            if it == 0:
                pool = [x for x in pool if 410 < x < w.shape[1]]
            n_pool.append(len(pool))
            if len(pool) == 0:
                continue
            k = len(pool) if len(pool) < sss else sss
            for x in prng.choices(pool, k=k):
                specs.append((i0, i1, j0, j1, x, 400 if it > 0 else 0))
        depl = ss - len(specs)
        if depl > 0:
            rich = int(np.argmax(n_pool))
            i0, i1, j0, j1 = _unit_geometry(mat, chrom, rich, unit, r)
            w = nan0(mat.window(chrom, i0, i1, j0, j1))
            live = w.sum(axis=1) != 0
            pool = [x for x in range(w.shape[0]) if live[x] and 20 < x < unit - 20]
            for x in prng.choices(pool, k=depl):
                specs.append((i0, i1, j0, j1, x, 400 if it > 0 else 0))  # stale `it`
        plans.append(specs)
    return names2, plans


def _pyslice(a, b, n):
    """Python slice(a, b) on a length-n axis -> (lo, hi), hi <= lo means empty."""
    if a < 0:
        a = max(a + n, 0)
    else:
        a = min(a, n)
    if b < 0:
        b = max(b + n, 0)
    else:
        b = min(b, n)
    return a, b


def background_columns(mat, chrom, specs):
    """The four (400 x nsamples) tables of one chromosome: centre-block mean minus the block
    above (left) / below (right), towards the diagonal (up) and away from it (down), with
    numpy slice semantics on the fetched sub-matrix (getStripe.py:347-368)."""
    r = mat.binsize
    bsf = 50000 / r
    bu = int(np.floor(bsf / 2))
    bd = int(bsf - np.floor(bsf / 2))
    bs = int(bsf)
    n = len(specs)
    lu = np.zeros((NDIST, n)); ru = np.zeros((NDIST, n))
    ld = np.zeros((NDIST, n)); rd = np.zeros((NDIST, n))
    cache = {}
    for k, (i0, i1, j0, j1, x, shift) in enumerate(specs):
        key = (i0, i1, j0, j1)
        if key not in cache:
            cache.clear()
            cache[key] = nan0(mat.window(chrom, i0, i1, j0, j1))
        w = cache[key]
        nr, nc = w.shape
        rows = [_pyslice(x - bu - bs, x - bu, nr), _pyslice(x - bu, x + bd, nr),
                _pyslice(x + bd, x + bd + bs, nr)]
        strips = [w[a:b].sum(axis=0) if b > a else None for a, b in rows]
        nrows = [max(b - a, 0) for a, b in rows]
        for j in range(NDIST):
            for sign, (tl, tr) in ((-1, (lu, ru)), (1, (ld, rd))):
                y = x + sign * j + shift
                ca, cb = _pyslice(y - bu, y + bd, nc)
                means = []
                for s, nrw in zip(strips, nrows):
                    cnt = nrw * max(cb - ca, 0)
                    if cnt == 0:
                        means.append(np.nan)
                    else:
                        means.append(s[ca:cb].sum() / cnt)
                tl[j, k] = means[1] - means[0]
                tr[j, k] = means[1] - means[2]
    return lu, ru, ld, rd


def null_distribution(mat, all_names, all_sizes, seed):
    """nulldist() -> bgleft_up, bgright_up, bgleft_down, bgright_down."""
    names2, plans = plan_samples(mat, all_names, all_sizes, seed)
    parts = [background_columns(mat, c, s) for c, s in zip(names2, plans)]
    return tuple(np.column_stack([p[i] for p in parts]) if parts else np.zeros((NDIST, 0))
                 for i in range(4))
