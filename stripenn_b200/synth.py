"""Synthetic Hi-C contact maps with planted stripes.

There is no network and no HDF5 in the build image, so every test, golden fixture and bench
line uses these generators.  A map is produced directly in diagonal-band form
(offsets d = 0..600), then flattened to the upper-triangle COO pixel table a ``.cool`` file
would hold:

* ``count[i, i+d] ~ Poisson(60 * (1+d)**-0.9 * e_ij)``, zero beyond d = 600;
* ``K = max(4, N // 200)`` planted stripes per chromosome with enrichment ``e = 3``
  (down: rows a..a+L of columns a..a+w, i.e. hanging below the diagonal; up: rows a-L..a of
  columns a-w..a), anchor ``a ~ U{100..N-100}``, ``L ~ U{40..200}``, ``w ~ U{2..4}``;
* one 300-bin all-zero gap at N // 3 plus 0.5 % isolated all-zero bins (exercise the
  zero-column compaction of getStripe.py:812-821 and ``available_cols``);
* optional balancing: ``weight_i = 1/sqrt(cov_i/mean(cov))`` with 2 % of bins NaN, stored as a
  multiplicative ``weight`` column and as its reciprocal in a divisive ``KR`` column.

``rng = np.random.default_rng([seed, chrom_index])`` so chromosomes are independent of each
other and of the order they are generated in.
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from .contactmap import ChromPixels, ContactMap

DEFAULT_SEED = 20240501
BAND = 601  # offsets 0..600 carry signal

MM10 = {"1": 195471971, "2": 182113224, "3": 160039680, "4": 156508116, "5": 151834684,
        "6": 149736546, "7": 145441459, "8": 129401213, "9": 124595110, "10": 130694993,
        "11": 122082543, "12": 120129022, "13": 120421639, "14": 124902244, "15": 104043685,
        "16": 98207768, "17": 94987271, "18": 90702639, "19": 61431566, "X": 171031299}
HG38 = {"chr1": 248956422, "chr2": 242193529, "chr3": 198295559, "chr4": 190214555,
        "chr5": 181538259, "chr6": 170805979, "chr7": 159345973, "chr8": 145138636,
        "chr9": 138394717, "chr10": 133797422, "chr11": 135086622, "chr12": 133275309,
        "chr13": 114364328, "chr14": 107043718, "chr15": 101991189, "chr16": 90338345,
        "chr17": 83257441, "chr18": 80373285, "chr19": 58617616, "chr20": 64444167,
        "chr21": 46709983, "chr22": 50818468, "chrX": 156040895, "chrY": 57227415}


def make_chrom(name: str, size_bp: int, binsize: int, seed: int, chrom_index: int,
               balanced: bool) -> Tuple[ChromPixels, Optional[np.ndarray], List[dict]]:
    """Generate one chromosome.  Returns (pixels, weight-or-None, planted stripe list)."""
    rng = np.random.default_rng([seed, chrom_index])
    n = int(math.ceil(size_bp / binsize))
    nd = min(BAND, n)
    lam = (60.0 * (1.0 + np.arange(nd, dtype=np.float64)) ** -0.9).astype(np.float32)
    enr = np.ones((n, nd), dtype=np.float32)

    planted: List[dict] = []
    k = max(4, n // 200)
    lo_a, hi_a = 100, n - 100
    for _ in range(k):
        if hi_a <= lo_a:
            break
        a = int(rng.integers(lo_a, hi_a + 1))
        length = int(rng.integers(40, 201))
        w = int(rng.integers(2, 5))
        down = bool(rng.integers(0, 2))
        if down:
            # image rows a..a+L, image cols a..a+w  ==  upper-triangle rows c, cols c..a+L
            for c in range(a, min(a + w, n - 1) + 1):
                d1 = min(a + length, n - 1) - c
                if d1 >= 0:
                    enr[c, 0:min(d1, nd - 1) + 1] = 3.0
        else:
            # image rows a-L..a, cols a-w..a (already upper triangle where row <= col)
            for c in range(max(a - w, 0), a + 1):
                r0 = max(a - length, 0)
                rows = np.arange(r0, min(a, c) + 1)
                d = c - rows
                ok = d < nd
                enr[rows[ok], d[ok]] = 3.0
        planted.append({"anchor": a, "length": length, "width": w, "down": down})

    mean = enr * lam[None, :]
    counts = rng.poisson(mean).astype(np.int32)
    del mean, enr
    # entries past the chromosome end
    dd = np.arange(nd)[None, :]
    ii = np.arange(n)[:, None]
    counts[(ii + dd) >= n] = 0

    # gaps: all-zero rows/columns
    dead = np.zeros(n, dtype=bool)
    g0 = n // 3
    dead[g0:min(g0 + 300, n)] = True
    dead |= rng.random(n) < 0.005
    counts[dead, :] = 0
    jj = np.minimum(ii + dd, n - 1)
    counts[dead[jj]] = 0
    del jj

    r, d = np.nonzero(counts)
    pix = ChromPixels(name, int(size_bp), n, r.astype(np.int32), (r + d).astype(np.int32),
                      counts[r, d].astype(np.int32))
    weight = None
    if balanced:
        cov = np.zeros(n, dtype=np.float64)
        np.add.at(cov, pix.bin1, pix.count)
        off = pix.bin1 != pix.bin2
        np.add.at(cov, pix.bin2[off], pix.count[off])
        live = cov > 0
        weight = np.full(n, np.nan, dtype=np.float64)
        weight[live] = 1.0 / np.sqrt(cov[live] / cov[live].mean())
        weight[rng.random(n) < 0.02] = np.nan
    return pix, weight, planted


def make_map(chroms: Sequence[Tuple[str, int]], binsize: int, seed: int = DEFAULT_SEED,
             balanced: bool = False, only: Optional[Iterable[str]] = None) -> ContactMap:
    """Build a :class:`ContactMap`.  ``only`` limits pixel generation to some chromosomes
    (sizes of the others are still listed, as in a real file)."""
    names = [c[0] for c in chroms]
    sizes = np.array([c[1] for c in chroms], dtype=np.int64)
    cm = ContactMap(binsize=int(binsize), chromnames=names, chromsizes=sizes)
    if balanced:
        cm.weights["weight"] = {}
        cm.weights["KR"] = {}
    keep = set(only) if only is not None else None
    cm.planted = {}  # type: ignore[attr-defined]
    for idx, (name, size) in enumerate(chroms):
        if keep is not None and name not in keep:
            continue
        pix, w, planted = make_chrom(name, int(size), binsize, seed, idx, balanced)
        cm.chroms[name] = pix
        cm.planted[name] = planted  # type: ignore[attr-defined]
        if balanced:
            cm.weights["weight"][name] = w
            with np.errstate(divide="ignore", invalid="ignore"):
                cm.weights["KR"][name] = 1.0 / w
    return cm


# ---------------------------------------------------------------------- named configs
def config_chroms(cfg: str) -> Tuple[List[Tuple[str, int]], int, bool]:
    """(chromosomes, binsize, balanced) of BASELINE.json configs."""
    if cfg == "cfg1":  # mm10 chr16 @ 5 kb
        return [("16", MM10["16"])], 5000, False
    if cfg == "cfg2":  # mouse whole genome @ 10 kb
        return [(k, v) for k, v in MM10.items()], 10000, False
    if cfg in ("cfg3", "cfg5"):  # human micro-C-like @ 5 kb, KR balanced
        return [(k, v) for k, v in HG38.items()], 5000, True
    if cfg == "cfg4":  # hg38 chr1 @ 1 kb
        return [("chr1", HG38["chr1"])], 1000, False
    if cfg == "tiny":  # 2,400-bin toy chromosome used by the golden fixtures
        return [("T1", 2400 * 5000 - 1234)], 5000, False
    if cfg == "tiny1k":  # 1 kb bins: bs = 50, exercises the wide flanks / negative-slice wraps
        return [("U1", 2600 * 1000 - 321)], 1000, False
    if cfg == "tiny10k":  # 10 kb bins: bs = 5, odd block split (bu = 2, bd = 3), balanced
        return [("V1", 2100 * 10000 - 4321), ("V2", 1300 * 10000 - 99)], 10000, True
    if cfg == "mid":  # config-size parity fixture: a 12,050-bin KR chromosome + one shorter than 500 bins
        return [("chrL", 12050 * 5000 - 777), ("chrS", 450 * 5000 - 13)], 5000, True
    if cfg == "tiny2":  # two small chromosomes, balanced
        return [("chrA", 2400 * 5000 - 1234), ("chrB", 1500 * 5000 - 77)], 5000, True
    raise ValueError("unknown config %r" % cfg)


def make_config(cfg: str, seed: int = DEFAULT_SEED, only: Optional[Iterable[str]] = None,
                balanced: Optional[bool] = None) -> ContactMap:
    chroms, binsize, bal = config_chroms(cfg)
    if balanced is not None:
        bal = balanced
    return make_map(chroms, binsize, seed, bal, only)


def frame_pixels(nbins: int) -> int:
    """Sum over search windows of (end-start+1)^2 (getStripe.py:794-801) — the unit of the
    Mpixels/s metric."""
    nframes = math.ceil(nbins / 200)
    tot = 0
    for idx in range(nframes):
        start = 0 if idx == 0 else idx * 200 - 100
        end = min((idx + 1) * 200 + 99, nbins - 1)
        tot += (end - start + 1) ** 2
    return tot


def frame_pixels_range(nbins: int, lo: int, hi: int) -> int:
    """frame_pixels restricted to the windows [lo, hi)."""
    tot = 0
    for idx in range(lo, min(hi, math.ceil(nbins / 200))):
        start = 0 if idx == 0 else idx * 200 - 100
        end = min((idx + 1) * 200 + 99, nbins - 1)
        tot += (end - start + 1) ** 2
    return tot


def make_stripe_table(cm: ContactMap, n: int, seed: int = DEFAULT_SEED,
                      chroms: Optional[Sequence[str]] = None) -> Dict[str, np.ndarray]:
    """Candidate table for ``score`` mode (config 5): chromosome ∝ size, anchor uniform,
    h = round(lognormal(ln 55, 0.6)) clipped to [12, 300], w ~ U{3..12}, direction coin flip;
    constrained so the flanks stay inside the chromosome (a - bs >= 1, a+w+bs < N-1,
    a+w-h >= 0)."""
    rng = np.random.default_rng([seed, 555])
    r = cm.binsize
    bs = int(50000 / r)
    names = list(chroms) if chroms is not None else list(cm.chroms.keys())
    sizes = np.array([cm.chromsizes[cm.chromnames.index(c)] for c in names], dtype=np.float64)
    which = rng.choice(len(names), size=n, p=sizes / sizes.sum())
    out = {"chr": [], "pos1": [], "pos2": [], "pos3": [], "pos4": []}
    for ci in which:
        name = names[ci]
        nb = cm.nbins(name)
        while True:
            h = int(np.clip(np.round(rng.lognormal(math.log(55), 0.6)), 12, 300))
            w = int(rng.integers(3, 13))
            a = int(rng.integers(bs + 1, nb - w - bs - 2))
            down = bool(rng.integers(0, 2))
            if down and a + h < nb - 1:
                p1, p2 = a * r + 1, (a + w) * r
                p3, p4 = p1, (a + h) * r
                break
            if (not down) and a + w - h >= 0:
                p1, p2 = a * r + 1, (a + w) * r
                p3, p4 = (a + w - h) * r + 1, p2
                break
        out["chr"].append(name)
        out["pos1"].append(p1)
        out["pos2"].append(p2)
        out["pos3"].append(p3)
        out["pos4"].append(p4)
    return {"chr": np.array(out["chr"], dtype=object),
            "pos1": np.array(out["pos1"], dtype=np.int64),
            "pos2": np.array(out["pos2"], dtype=np.int64),
            "pos3": np.array(out["pos3"], dtype=np.int64),
            "pos4": np.array(out["pos4"], dtype=np.int64)}
