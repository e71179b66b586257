"""In-memory contact-map container and the data sources that feed the CUDA band builder.

The reference touches the matrix only through ``cooler.Cooler.matrix(balance).fetch``
(reference: src/stripenn/stripenn.py:80,118; every call site listed in SURVEY.md §8 a1).
The B200 path does not fetch windows from the host at all: a chromosome's upper-triangle COO
pixels (``bin1 <= bin2``, int32 counts) and its per-bin balancing weights are uploaded once
and turned into a diagonal band in HBM (csrc/band.cu).  This module holds the host-side
representation of that upload plus (de)serialisation to ``.npz`` so that tests, the bench and
the reference harness all read byte-identical inputs.

``.cool/.mcool`` ingestion goes through :func:`ContactMap.from_cooler`, which imports
``cooler`` lazily (it is not installed in the build image).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

# cooler interprets these weight columns as divisive unless told otherwise: cooler/api.py
# `matrix(..., divisive_weights=None)` switches on `balance in _4DN_DIVISIVE_WEIGHTS`, the set
# {"KR", "VC", "SQRT_VC"} (4DN's hic2cool column names; restated from cooler's source from memory —
# cooler is not installed here).  hic2cool's older "VC_SQRT" spelling, which the reference's CLI
# help mentions, is NOT in that set: through cooler it is applied multiplicatively, so it is here.
DIVISIVE_WEIGHT_NAMES = ("KR", "VC", "SQRT_VC")


@dataclass
class ChromPixels:
    """Upper-triangle COO pixels of one chromosome, bin ids local to the chromosome."""

    name: str
    size_bp: int
    nbins: int
    bin1: np.ndarray  # int32, bin1 <= bin2
    bin2: np.ndarray  # int32
    count: np.ndarray  # int32

    @property
    def nnz(self) -> int:
        return int(self.bin1.shape[0])


@dataclass
class ContactMap:
    binsize: int
    chromnames: List[str]
    chromsizes: np.ndarray  # int64, same order as chromnames
    chroms: Dict[str, ChromPixels] = field(default_factory=dict)
    # weights[column][chrom] -> float64[nbins] (NaN = masked bin)
    weights: Dict[str, Dict[str, np.ndarray]] = field(default_factory=dict)

    # ------------------------------------------------------------------ helpers
    def nbins(self, chrom: str) -> int:
        return int(math.ceil(int(self.chromsizes[self.chromnames.index(chrom)]) / self.binsize))

    def bin_columns(self) -> List[str]:
        """Column names of ``Cooler.bins()`` (chrom,start,end + weight columns)."""
        return ["chrom", "start", "end"] + list(self.weights.keys())

    def weight_vector(self, chrom: str, balance) -> Optional[np.ndarray]:
        """Effective multiplicative weight per bin for ``matrix(balance=...)``.

        ``False`` -> None (raw counts); ``True`` -> column ``weight``; a string -> that column,
        inverted when the name is one of cooler's divisive columns.
        """
        if balance is False or balance is None:
            return None
        col = "weight" if balance is True else str(balance)
        if col not in self.weights:
            raise KeyError("no balancing column %r" % col)
        w = np.asarray(self.weights[col][chrom], dtype=np.float64)
        if col in DIVISIVE_WEIGHT_NAMES:
            with np.errstate(divide="ignore", invalid="ignore"):
                w = 1.0 / w
        return w

    # ------------------------------------------------------------------ .npz
    def save_npz(self, path: str) -> None:
        out = {
            "binsize": np.int64(self.binsize),
            "chromnames": np.array(self.chromnames, dtype=object).astype(str),
            "chromsizes": np.asarray(self.chromsizes, dtype=np.int64),
            "weight_columns": np.array(list(self.weights.keys()), dtype=object).astype(str)
            if self.weights else np.array([], dtype=str),
        }
        for i, name in enumerate(self.chromnames):
            if name not in self.chroms:
                continue
            c = self.chroms[name]
            out["c%d_bin1" % i] = c.bin1.astype(np.int32)
            out["c%d_bin2" % i] = c.bin2.astype(np.int32)
            out["c%d_count" % i] = c.count.astype(np.int32)
            for col, per in self.weights.items():
                out["c%d_w_%s" % (i, col)] = np.asarray(per[name], dtype=np.float64)
        np.savez_compressed(path, **out)

    @classmethod
    def load_npz(cls, path: str) -> "ContactMap":
        z = np.load(path, allow_pickle=False)
        names = [str(x) for x in z["chromnames"]]
        sizes = z["chromsizes"].astype(np.int64)
        binsize = int(z["binsize"])
        cols = [str(x) for x in z["weight_columns"]]
        cm = cls(binsize=binsize, chromnames=names, chromsizes=sizes)
        for col in cols:
            cm.weights[col] = {}
        for i, name in enumerate(names):
            if "c%d_bin1" % i not in z:
                continue
            nb = int(math.ceil(int(sizes[i]) / binsize))
            cm.chroms[name] = ChromPixels(name, int(sizes[i]), nb, z["c%d_bin1" % i],
                                          z["c%d_bin2" % i], z["c%d_count" % i])
            for col in cols:
                cm.weights[col][name] = z["c%d_w_%s" % (i, col)]
        return cm

    # ------------------------------------------------------------------ cooler
    @classmethod
    def from_cooler(cls, uri: str, chroms: Optional[List[str]] = None) -> "ContactMap":
        """Read a ``.cool`` / ``.mcool::resolutions/N`` file through the ``cooler`` package.

        Mirrors what the reference gets from ``cooler.Cooler(cool)`` (stripenn.py:80-118):
        chromosome names/sizes, bin size, the ``bins`` weight columns and the upper-triangle
        pixel table of every cis block (pixels whose bin2 lies in another chromosome are
        dropped: stripenn never fetches a trans window).  Lazy import: cooler/h5py are absent
        from the build image; tests/test_cooler_adaptor.py drives this code through a stand-in
        module with cooler's public API (tests/fake_cooler.py).
        """
        try:
            import cooler  # type: ignore
        except Exception as e:  # pragma: no cover - cannot be exercised offline
            raise RuntimeError(
                "reading %r needs the 'cooler' package (not installed); "
                "convert to the .npz layout of stripenn_b200.contactmap instead" % uri) from e
        lib = cooler.Cooler(uri)
        names = list(lib.chromnames)
        sizes = np.asarray(lib.chromsizes.values, dtype=np.int64)
        cm = cls(binsize=int(lib.binsize), chromnames=names, chromsizes=sizes)
        bins = lib.bins()
        wcols = [c for c in bins.columns if c not in ("chrom", "start", "end")]
        for col in wcols:
            cm.weights[col] = {}
        for name in names:
            if chroms is not None and name not in chroms:
                continue
            lo, hi = lib.extent(name)
            px = lib.pixels().fetch(name)
            px = px[(px["bin2_id"] >= lo) & (px["bin2_id"] < hi)]
            cm.chroms[name] = ChromPixels(
                name, int(lib.chromsizes[name]), hi - lo,
                (px["bin1_id"].values - lo).astype(np.int32),
                (px["bin2_id"].values - lo).astype(np.int32),
                px["count"].values.astype(np.int32))
            b = bins.fetch(name)
            for col in wcols:
                cm.weights[col][name] = b[col].values.astype(np.float64)
        return cm


def open_contact_map(path: str, chroms: Optional[List[str]] = None) -> ContactMap:
    """``.npz`` -> :meth:`ContactMap.load_npz`; anything else -> cooler."""
    if str(path).endswith(".npz"):
        return ContactMap.load_npz(path)
    return ContactMap.from_cooler(path, chroms)
