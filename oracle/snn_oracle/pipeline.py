"""ORACLE (test infrastructure, never on the product path) — quantiles, search windows,
StripeSearch, RemoveRedundant and ``extract`` restated.

Follows reference src/stripenn/getStripe.py:160-176 (getQuantile_original), :790-862
(extract / search_frame), :864-1114 (StripeSearch), :1116-1196 (RemoveRedundant).
Tables are plain dicts of numpy arrays (no pandas) with the reference's column names.
"""
import math

import numpy as np

from . import canny as cn
from . import image as im
from . import segments as sg
from .matrix import nan0

COLS = ("chr", "pos1", "pos2", "chr2", "pos3", "pos4", "length", "width", "total", "Mean",
        "maxpixel", "num", "start", "end", "x", "y", "h", "w")


def empty_table():
    t = {k: [] for k in COLS}
    return t


def table_len(t):
    return len(t["pos1"])


def concat_tables(ts):
    out = empty_table()
    keys = set(COLS)
    for t in ts:
        keys |= set(t.keys())
    out = {k: [] for k in keys}
    for t in ts:
        n = table_len(t)
        for k in keys:
            out[k].extend(list(t[k]) if k in t else [None] * n)
    return out


def take_rows(t, idx):
    return {k: [v[i] for i in idx] for k, v in t.items()}


def quantiles(mat, chroms, qs):
    """getQuantile_original: np.quantile(mat[mat>0], qs) over the full symmetric matrix."""
    out = {}
    for c in chroms:
        out[c] = np.quantile(mat.positive_population(c), qs)
    return out


def quantiles_slow(mat, chroms, qs):
    """getQuantile_slow (getStripe.py:107-158, the -s flag), literally: genome-wide row blocks of
    w = floor(25e6 / L) rows whose slice end (k+1)*w - 1 drops each block's last row, r_start =
    ceil(cumsize_before / r) + 1, r_end = ceil(cumsize / r).  Its values differ from
    quantiles() by construction."""
    names = list(mat.cm.chromnames)
    sizes = np.asarray(mat.cm.chromsizes)
    nbin = mat.binsize
    cum = np.nancumsum(sizes)
    out = {}
    for ci in sorted(c for c in range(len(names)) if names[c] in chroms):
        CHROMSIZE = sizes[ci]
        L = int(np.ceil(CHROMSIZE / nbin))
        r_start = 0 if ci == 0 else cum[ci - 1]
        r_end = cum[ci]
        r_start = np.ceil(r_start / nbin) + 1
        r_end = np.ceil(r_end / nbin)
        hs = np.empty(0)
        w = int(np.floor(25000000 / L))
        for k in range(int(np.ceil(L / w))):
            w_start = k * w + r_start
            w_end = (k + 1) * w - 1
            if w_end >= np.floor(CHROMSIZE / nbin):
                w_end = int(np.floor(CHROMSIZE / nbin))
            w_end += r_start
            blk = mat.genome_window(int(w_start), int(w_end), int(r_start), int(r_end))
            with np.errstate(invalid="ignore"):
                hs = np.concatenate((hs, blk[blk > 0]))
        out[names[ci]] = np.quantile(hs, qs)
    return out


def maxpixel_label(q):
    return str(q * 100) + "%"   # getStripe.py:1096


def frame_geometry(idx, nbins):
    """getStripe.py:794-801."""
    start = idx * 200 - 100
    end = (idx + 1) * 200 + 99
    if end >= nbins:
        end = nbins - 1
    if idx == 0:
        start = 0
    return start, end


def load_frame(mat, chrom, idx):
    """search_frame (getStripe.py:792-825) up to the call of StripeSearch.
    Returns None when <= 10 non-empty columns remain."""
    r = mat.binsize
    nbins = mat.nbins(chrom)
    csize = mat.chromsize(chrom)
    start, end = frame_geometry(idx, nbins)
    n = end - start + 1
    sa = (start + np.arange(n, dtype=np.int64)) * r + 1
    ea = sa + r - 1
    if ea[-1] >= csize:
        ea[-1] = csize
    raw = mat.fetch_bp(chrom, int(sa[0]), int(ea[-1]), int(sa[0]), int(ea[-1]))
    is_int = np.issubdtype(raw.dtype, np.integer)
    D = raw if is_int else nan0(raw)
    nz = np.nonzero(D.sum(axis=0) != 0)[0]
    if len(nz) <= 10:
        return None
    return dict(idx=idx, start=start, end=end, D=D[np.ix_(nz, nz)], sa=sa[nz], ea=ea[nz],
                S=len(nz))


def stripe_search(fr, M, q, chrom, minH, maxW, sigma, bf, variant=im.GRAY_IPP,
                  stage_hook=None):
    """StripeSearch (getStripe.py:864-1114) for one window and one maxpixel value,
    including the per-window RemoveRedundant('size')."""
    D, sa, ea = fr["D"], fr["sa"], fr["ea"]
    S = fr["S"]
    G = im.saturate(D, M)
    weights = cn.gaussian_weights(sigma)
    bleed = cn.bleed_over((S, S), *weights)
    t = empty_table()
    Df = D.astype(np.float64)
    for bi, b in enumerate(im.BRIGHTNESS):
        gray = im.gray_image(G, b, bf, variant)
        edges = cn.canny(gray, sigma, weights, bleed)
        vert = sg.vertical_line(edges)
        found = sg.search_pass(edges, vert, minH, maxW)
        if stage_hook is not None:
            stage_hook(fr, q, bi, gray, edges, vert, found)
        for (x, y, w, h) in found:
            tot = D[y:y + h, x:x + w].sum()
            with np.errstate(divide="ignore", invalid="ignore"):
                mean = np.float64(tot) / h / w
            p1, p2 = int(sa[x]), int(ea[x + w - 1])
            p3, p4 = int(sa[y]), int(ea[y + h - 1])
            row = dict(chr=chrom, pos1=p1, pos2=p2, chr2=chrom, pos3=p3, pos4=p4,
                       length=p4 - p3 + 1, width=p2 - p1 + 1, total=tot, Mean=mean,
                       maxpixel=maxpixel_label(q), num=fr["idx"], start=fr["start"],
                       end=fr["end"], x=x, y=y, h=h, w=w)
            for k in COLS:
                t[k].append(row[k])
    return remove_redundant(t, "size")


def overlap_marks(t, by):
    """Boolean keep-mask of RemoveRedundant (getStripe.py:1116-1196): every pair of rows of
    one chromosome whose window ids differ by at most 1 is tested; marks accumulate
    (non-greedy)."""
    n = table_len(t)
    keep = np.ones(n, dtype=bool)
    if n == 0:
        return keep
    chrs = np.array(t["chr"], dtype=object)
    num = np.array(t["num"], dtype=np.int64)
    p1 = np.array(t["pos1"], dtype=np.int64)
    p2 = np.array(t["pos2"], dtype=np.int64)
    p3 = np.array(t["pos3"], dtype=np.int64)
    p4 = np.array(t["pos4"], dtype=np.int64)
    hh = np.array(t["h"], dtype=np.float64)
    ww = np.array(t["w"], dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        ratio = hh / ww
    if by == "pvalue":
        key = np.array(t["pvalue"], dtype=np.float64)
    elif by == "score":
        key = np.array(t["Stripiness"], dtype=np.float64)
    elif by != "size":
        raise ValueError('"by" should be one of "size", "pvalue" and "score"')
    for c in set(chrs.tolist()):
        ci = np.nonzero(chrs == c)[0]
        order = ci[np.argsort(num[ci], kind="stable")]
        nums = num[order]
        lo = np.searchsorted(nums, nums - 1, side="left")
        hi = np.searchsorted(nums, nums + 1, side="right")
        for a in range(len(order)):
            i = order[a]
            others = order[lo[a]:hi[a]]
            others = others[others > i]          # each unordered pair once, i earlier row
            if len(others) == 0:
                continue
            ix = np.minimum(p2[i], p2[others]) - np.maximum(p1[i], p1[others]) + 1
            iy = np.minimum(p4[i], p4[others]) - np.maximum(p3[i], p3[others]) + 1
            ix = np.maximum(ix, 0)
            iy = np.maximum(iy, 0)
            with np.errstate(divide="ignore", invalid="ignore"):
                sx = ix / np.minimum(p2[i] - p1[i], p2[others] - p1[others])
                sy = iy / np.minimum(p4[i] - p3[i], p4[others] - p3[others])
            hit = (sx > 0.2) & (sy > 0.2)
            if not hit.any():
                continue
            js = others[hit]
            if by == "size":
                first = ratio[i] <= ratio[js]
            elif by == "score":
                first = key[i] <= key[js]
            else:
                first = key[i] > key[js]
            if first.any():
                keep[i] = False
            keep[js[~first]] = False
    return keep


def remove_redundant(t, by):
    keep = overlap_marks(t, by)
    return take_rows(t, np.nonzero(keep)[0])


def n_frames(nbins):
    return math.ceil(nbins / 200)


def extract_candidates(mat, chroms, MP, qi, q, minH, maxW, sigma, bf, variant=im.GRAY_IPP,
                       frame_filter=None, stage_hook=None):
    """extract() (getStripe.py:790-853) up to and including the cross-window
    RemoveRedundant('size'); the p-value column is added by the caller."""
    parts = []
    for chrom in chroms:
        M = MP[chrom][qi]
        for idx in range(n_frames(mat.nbins(chrom))):
            if frame_filter is not None and not frame_filter(chrom, idx):
                continue
            fr = load_frame(mat, chrom, idx)
            if fr is None:
                continue
            parts.append(stripe_search(fr, M, q, chrom, minH, maxW, sigma, bf, variant,
                                       stage_hook))
    return remove_redundant(concat_tables(parts) if parts else empty_table(), "size")
