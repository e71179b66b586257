"""ORACLE (test infrastructure, never on the product path) — the data boundary:
``cooler.Cooler.matrix(balance).fetch`` semantics restated on a ContactMap
(reference call sites: src/stripenn/getStripe.py:172,193,266,326,432,518,560,684,690,696,808;
semantics table: SURVEY.md §8 a1; cooler itself is an absent third-party dependency —
cooler/api.py ``matrix``: dense window = mirrored counts * np.outer(w_rows, w_cols)).
"""
import math

import numpy as np
import scipy.sparse as sp


class OracleMatrix:
    def __init__(self, cm, balance=False):
        self.cm = cm
        self.binsize = int(cm.binsize)
        self.balance = balance
        self._csr = {}
        self._w = {}

    def chromsize(self, chrom):
        return int(self.cm.chromsizes[self.cm.chromnames.index(chrom)])

    def nbins(self, chrom):
        return int(math.ceil(self.chromsize(chrom) / self.binsize))

    def _mat(self, chrom):
        if chrom not in self._csr:
            c = self.cm.chroms[chrom]
            b1 = c.bin1.astype(np.int64)
            b2 = c.bin2.astype(np.int64)
            off = b1 != b2
            rows = np.concatenate([b1, b2[off]])
            cols = np.concatenate([b2, b1[off]])
            vals = np.concatenate([c.count, c.count[off]]).astype(np.int32)
            self._csr[chrom] = sp.csr_matrix((vals, (rows, cols)), shape=(c.nbins, c.nbins))
            self._w[chrom] = self.cm.weight_vector(chrom, self.balance)
        return self._csr[chrom], self._w[chrom]

    def window(self, chrom, i0, i1, j0, j1):
        """Dense rows [i0,i1) x cols [j0,j1) in bin coordinates."""
        m, w = self._mat(chrom)
        arr = m[i0:i1, j0:j1].toarray()
        if w is None:
            return arr
        with np.errstate(invalid="ignore"):
            return arr * np.outer(w[i0:i1], w[j0:j1])

    def bins_of(self, chrom, start, end):
        """bp half-open [start,end) -> bin range, with cooler's range checks."""
        size = self.chromsize(chrom)
        if end < start:
            raise ValueError("End cannot be less than start")
        if start < 0 or end > size:
            raise ValueError("Genomic region out of bounds: [%d, %d)" % (start, end))
        return start // self.binsize, int(math.ceil(end / self.binsize))

    def fetch_bp(self, chrom, r_start, r_end, c_start, c_end):
        i0, i1 = self.bins_of(chrom, r_start, r_end)
        j0, j1 = self.bins_of(chrom, c_start, c_end)
        return self.window(chrom, i0, i1, j0, j1)

    def positive_population(self, chrom):
        """Multiset of positive pixels of the full symmetric matrix
        (what ``mat[mat > 0]`` selects at getStripe.py:172-173), unsorted."""
        c = self.cm.chroms[chrom]
        _, w = self._mat(chrom)
        if w is None:
            v = c.count.astype(np.float64)
        else:
            with np.errstate(invalid="ignore"):
                v = c.count * (w[c.bin1] * w[c.bin2])
        pos = v > 0
        off = (c.bin1 != c.bin2) & pos
        return np.concatenate([v[pos], v[off]])


    def genome_window(self, i0, i1, j0, j1):
        """``selector[i0:i1, j0:j1]`` in GENOME-WIDE bin coordinates (cooler RangeSelector2D
        indexing, used only at getStripe.py:122): chromosome k occupies bins
        [sum(ceil(size_j/r), j<k), ...); numpy-style clipping; blocks between different
        chromosomes are zero (a ContactMap is cis-only) times the weights' outer product."""
        r = self.binsize
        names = list(self.cm.chromnames)
        nbs = [int(math.ceil(int(sz) / r)) for sz in self.cm.chromsizes]
        offs = np.concatenate([[0], np.cumsum(nbs)]).astype(np.int64)
        total = int(offs[-1])
        i0, i1, _ = slice(i0, i1).indices(total)
        j0, j1, _ = slice(j0, j1).indices(total)
        i1, j1 = max(i0, i1), max(j0, j1)
        arr = np.zeros((i1 - i0, j1 - j0), dtype=np.float64)
        wg = np.ones(total)
        for k, name in enumerate(names):
            o, n = int(offs[k]), nbs[k]
            if name not in self.cm.chroms:
                if self.balance is not False and self.balance is not None:
                    wg[o:o + n] = np.nan
                continue
            m, w = self._mat(name)
            if w is not None:
                wg[o:o + n] = w
            a0, a1 = max(i0, o), min(i1, o + n)
            b0, b1 = max(j0, o), min(j1, o + n)
            if a0 < a1 and b0 < b1:
                arr[a0 - i0:a1 - i0, b0 - j0:b1 - j0] = m[a0 - o:a1 - o, b0 - o:b1 - o].toarray()
        with np.errstate(invalid="ignore"):
            return arr * np.outer(wg[i0:i1], wg[j0:j1])


def nan0(a):
    a = np.array(a, dtype=np.float64, copy=True)
    a[np.isnan(a)] = 0.0
    return a
