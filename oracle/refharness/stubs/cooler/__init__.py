"""Fake ``cooler`` package: just enough of ``cooler.Cooler`` for the UNMODIFIED reference
drivers ``stripenn.compute`` / ``score.getScore`` to run on the ``.npz`` contact maps of
``stripenn_b200.contactmap`` (cooler, h5py and libhdf5 are not installed; no network).

This file DEFINES the ``matrix(balance).fetch`` semantics the whole project is tested against
(cooler/api.py ``matrix``/``RangeSelector2D`` restated from its documented
behaviour):

* region ``"chr:start-end"`` is a half-open bp interval -> bins ``[start // r, ceil(end / r))``;
  a bare ``"chr"`` is the whole chromosome; ``end < start`` or ``end > chromsize`` raise
  ``ValueError`` (that is why the reference clamps, getStripe.py:187-190,317-322,805-806);
* ``balance=False`` -> raw int32 counts; ``True`` -> column ``weight``; a name -> that bins
  column, inverted first when it is one of cooler's divisive names
  (`_4DN_DIVISIVE_WEIGHTS` = KR, VC, SQRT_VC);
* dense balanced window = ``counts * np.outer(w_rows, w_cols)`` (outer product first), so a
  bin with NaN weight yields a whole NaN row and column — zeros included;
* the matrix is symmetric: the stored upper triangle is mirrored.

TEST INFRASTRUCTURE ONLY (oracle/refharness) — never imported by the product.
"""
import math
import re

import numpy as np
import scipy.sparse as sp

__version__ = "0.0-fake"

_DIVISIVE = ("KR", "VC", "SQRT_VC")
_CACHE = {}


class _SizeArray(np.ndarray):
    """chromsizes: a numpy array (positional indexing, fancy indexing, arithmetic — what the
    reference does at stripenn.py:95-116 and getStripe.py:33,242,805) that also answers
    ``.keys()`` like the pandas Series real cooler returns (getStripe.py:164)."""

    def __new__(cls, values, names):
        obj = np.asarray(values, dtype=np.int64).view(cls)
        obj._names = list(names)
        return obj

    def __array_finalize__(self, obj):
        self._names = getattr(obj, "_names", None)

    def keys(self):
        return list(self._names)

    def __getitem__(self, key):
        # a pandas Series is also indexed by chromosome NAME (seeimage.py:63)
        if isinstance(key, str):
            return np.int64(np.asarray(self)[self._names.index(key)])
        return super().__getitem__(key)


class _Columns(list):
    pass


class _Bins:
    def __init__(self, cols):
        self.columns = _Columns(cols)


def _load(path):
    if path in _CACHE:
        return _CACHE[path]
    z = np.load(path, allow_pickle=False)
    names = [str(x) for x in z["chromnames"]]
    sizes = z["chromsizes"].astype(np.int64)
    binsize = int(z["binsize"])
    wcols = [str(x) for x in z["weight_columns"]]
    mats, weights = {}, {c: {} for c in wcols}
    for i, name in enumerate(names):
        if "c%d_bin1" % i not in z:
            continue
        nb = int(math.ceil(int(sizes[i]) / binsize))
        b1 = z["c%d_bin1" % i].astype(np.int64)
        b2 = z["c%d_bin2" % i].astype(np.int64)
        cnt = z["c%d_count" % i].astype(np.int32)
        off = b1 != b2
        rows = np.concatenate([b1, b2[off]])
        cols = np.concatenate([b2, b1[off]])
        vals = np.concatenate([cnt, cnt[off]])
        mats[name] = sp.csr_matrix((vals, (rows, cols)), shape=(nb, nb), dtype=np.int32)
        for c in wcols:
            weights[c][name] = z["c%d_w_%s" % (i, c)].astype(np.float64)
    data = dict(names=names, sizes=sizes, binsize=binsize, wcols=wcols, mats=mats,
                weights=weights)
    _CACHE[path] = data
    return data


_REGION = re.compile(r"^(.+?)(?::([0-9,]+)-([0-9,]+))?$")


class _Selector:
    """``Cooler.matrix(balance=...)`` — only ``fetch`` is used by the reference."""

    def __init__(self, path, balance):
        self._path = path
        self._balance = balance
        self.nfetch = 0

    def __reduce__(self):  # pickle by reference: loky/cloudpickle copies stay cheap
        return (_Selector, (self._path, self._balance))

    def _parse(self, region):
        d = _load(self._path)
        region = str(region)
        chrom, start, end = None, None, None
        if ":" in region:
            chrom, rest = region.rsplit(":", 1)
            s, e = rest.split("-")
            start, end = int(s.replace(",", "")), int(e.replace(",", ""))
        else:
            chrom = region
        if chrom not in d["names"]:
            raise ValueError("Unknown sequence label: %s" % chrom)
        size = int(d["sizes"][d["names"].index(chrom)])
        if start is None:
            start, end = 0, size
        if end < start:
            raise ValueError("End cannot be less than start")
        if start < 0 or end > size:
            raise ValueError("Genomic region out of bounds: [%d, %d)" % (start, end))
        r = d["binsize"]
        return chrom, start // r, int(math.ceil(end / r))

    def fetch(self, region, region2=None):
        d = _load(self._path)
        if region2 is None:
            region2 = region
        c1, i0, i1 = self._parse(region)
        c2, j0, j1 = self._parse(region2)
        if c1 != c2:
            raise NotImplementedError("trans windows are never requested by stripenn")
        self.nfetch += 1
        arr = d["mats"][c1][i0:i1, j0:j1].toarray()
        bal = self._balance
        if bal is False or bal is None:
            return arr
        col = "weight" if bal is True else str(bal)
        w = d["weights"][col][c1]
        if col in _DIVISIVE:
            with np.errstate(divide="ignore", invalid="ignore"):
                w = 1.0 / w
        with np.errstate(invalid="ignore"):
            return arr * np.outer(w[i0:i1], w[j0:j1])


    def __getitem__(self, key):
        """``matrix(...)[a:b, c:d]`` — GENOME-WIDE bin coordinates (the bins of all chromosomes
        of the file concatenated, chromosome k starting at sum(ceil(size_j / binsize), j < k)),
        used only by getStripe.getQuantile_slow (getStripe.py:122).  The .npz contact maps are
        cis-only, so blocks between different chromosomes are zero; balancing multiplies by the
        outer product of the concatenated weight vector like ``fetch``.  Slices clip like numpy."""
        d = _load(self._path)
        rs, cs = key
        r = d["binsize"]
        nbs = [int(math.ceil(int(sz) / r)) for sz in d["sizes"]]
        offs = np.concatenate([[0], np.cumsum(nbs)]).astype(np.int64)
        total = int(offs[-1])

        def rng(sl):
            a, b, step = sl.indices(total)
            if step != 1:
                raise IndexError("slice step is not supported")
            return a, max(a, b)
        i0, i1 = rng(rs)
        j0, j1 = rng(cs)
        self.nfetch += 1
        arr = np.zeros((i1 - i0, j1 - j0), dtype=np.int32)
        bal = self._balance
        col = None if (bal is False or bal is None) else ("weight" if bal is True else str(bal))
        wg = np.full(total, np.nan) if col else None
        for k, name in enumerate(d["names"]):
            o, n = int(offs[k]), nbs[k]
            if col and name in d["weights"][col]:
                wg[o:o + n] = d["weights"][col][name]
            a0, a1 = max(i0, o), min(i1, o + n)
            b0, b1 = max(j0, o), min(j1, o + n)
            if a0 < a1 and b0 < b1 and name in d["mats"]:
                arr[a0 - i0:a1 - i0, b0 - j0:b1 - j0] = d["mats"][name][a0 - o:a1 - o, b0 - o:b1 - o].toarray()
        if not col:
            return arr
        if col in _DIVISIVE:
            with np.errstate(divide="ignore", invalid="ignore"):
                wg = 1.0 / wg
        with np.errstate(invalid="ignore"):
            return arr * np.outer(wg[i0:i1], wg[j0:j1])


class Cooler:
    def __init__(self, path):
        self._path = str(path)
        d = _load(self._path)
        self.chromnames = list(d["names"])
        self.chromsizes = _SizeArray(d["sizes"], d["names"])
        self.binsize = d["binsize"]
        self._info = {"bin-size": d["binsize"]}

    def bins(self):
        return _Bins(["chrom", "start", "end"] + list(_load(self._path)["wcols"]))

    def matrix(self, balance=True, **_kw):
        return _Selector(self._path, balance)
