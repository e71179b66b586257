"""A stand-in for the `cooler` package with the part of its PUBLIC API that
stripenn_b200.contactmap.ContactMap.from_cooler uses (cooler, h5py and libhdf5 are not in the
image): Cooler(uri) with .chromnames, .chromsizes (pandas Series indexed by name), .binsize,
.extent(name) -> (lo, hi) in genome-wide bin ids, .bins() -> selector with .columns and
.fetch(name) -> DataFrame, .pixels() -> selector whose .fetch(name) returns every pixel whose
bin1 lies in the chromosome (cis AND trans, genome-wide bin ids), like cooler's pixel table.
Built from a ContactMap so the round trip can be checked.  TEST INFRASTRUCTURE ONLY."""
import math

import numpy as np
import pandas as pd

REGISTRY = {}      # uri -> (ContactMap, trans pixel DataFrame or None)
__version__ = "0.0-test"


class _BinSelector:
    def __init__(self, c):
        self._c = c
        self.columns = pd.Index(["chrom", "start", "end"] + list(c._cm.weights.keys()))

    def fetch(self, name):
        cm = self._c._cm
        n = cm.nbins(name)
        r = cm.binsize
        size = int(cm.chromsizes[cm.chromnames.index(name)])
        d = {"chrom": [name] * n, "start": np.arange(n) * r, "end": np.minimum((np.arange(n) + 1) * r, size)}
        for col, per in cm.weights.items():
            d[col] = per[name]
        lo, _ = self._c.extent(name)
        return pd.DataFrame(d, index=np.arange(lo, lo + n))


class _PixelSelector:
    def __init__(self, c):
        self._c = c

    def fetch(self, name):
        cm = self._c._cm
        lo, hi = self._c.extent(name)
        parts = []
        if name in cm.chroms:
            px = cm.chroms[name]
            parts.append(pd.DataFrame({"bin1_id": px.bin1.astype(np.int64) + lo,
                                       "bin2_id": px.bin2.astype(np.int64) + lo,
                                       "count": px.count.astype(np.int32)}))
        tr = self._c._trans
        if tr is not None:
            parts.append(tr[(tr["bin1_id"] >= lo) & (tr["bin1_id"] < hi)])
        df = pd.concat(parts) if parts else pd.DataFrame({"bin1_id": [], "bin2_id": [], "count": []})
        return df.sort_values(["bin1_id", "bin2_id"], kind="stable").reset_index(drop=True)


class Cooler:
    def __init__(self, uri):
        if uri not in REGISTRY:
            raise OSError("no such cooler: %s" % uri)
        self.uri = uri
        self._cm, self._trans = REGISTRY[uri]
        self.chromnames = list(self._cm.chromnames)
        self.chromsizes = pd.Series(np.asarray(self._cm.chromsizes, dtype=np.int64), index=self.chromnames,
                                    name="length")
        self.binsize = int(self._cm.binsize)
        self._info = {"bin-size": self.binsize}
        nb = [int(math.ceil(int(s) / self.binsize)) for s in self._cm.chromsizes]
        self._offs = np.concatenate([[0], np.cumsum(nb)]).astype(np.int64)

    def extent(self, name):
        i = self.chromnames.index(name)
        return int(self._offs[i]), int(self._offs[i + 1])

    def bins(self):
        return _BinSelector(self)

    def pixels(self, join=False):
        return _PixelSelector(self)
