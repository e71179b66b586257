"""TEST DOUBLE (tests only): an object with the interface of stripenn_b200.engine.Device whose
answers come from the numpy oracle.  It lets the CPU suite exercise the product's host logic
(getstripe.py sampling plan, table assembly, dedup calls, compute.py pipeline and its
multi-rank exchanges) without a GPU.  Never imported by the product."""
import math

import numpy as np

from snn_oracle import background as bgm
from snn_oracle import canny as cn
from snn_oracle import image as im
from snn_oracle import pipeline as pl
from snn_oracle import scoring as sc
from snn_oracle import segments as sg
from snn_oracle.matrix import OracleMatrix, nan0
from stripenn_b200._lib import CANDIDATE_DTYPE, NDIST


class FakeDevice:
    def __init__(self, cm, balance):
        self.cm = cm
        self.mat = OracleMatrix(cm, balance)
        self.loaded = {}
        self.bg = None
        self.mask = "0"

    def set_mask(self, cid, start_bp=0, end_bp=0):
        self.mask = "0" if cid is None else "%s:%d-%d" % (self._name(cid), start_bp, end_bp)

    def close(self):
        pass

    def _name(self, cid):
        return self.cm.chromnames[cid]

    def load_chrom(self, cid, nbins, binsize, chromsize, b1, b2, cnt, w=None, band_width=None, async_=False):
        self.loaded[cid] = dict(nbins=nbins, binsize=binsize, chromsize=chromsize, bw=band_width or 10 ** 9,
                                nnz=len(b1))

    def quantiles(self, cid, qs):
        return np.quantile(self.mat.positive_population(self._name(cid)), qs)

    def quantiles_multi(self, cids, qs):
        return np.array([self.quantiles(c, qs) for c in cids]).reshape(len(cids), len(qs))

    def expected(self, cid):
        c = self._name(cid)
        return np.array(bgm.expected_by_distance(self.mat, [c])[c])

    def row_flags(self, cid, geom):
        c = self._name(cid)
        nb = self.mat.nbins(c)
        avail = np.zeros(nb, np.uint8)
        live = np.zeros(nb, np.uint8)
        for a0, a1, i0, i1, j0, j1 in np.asarray(geom).reshape(-1, 6):
            s = nan0(self.mat.window(c, a0, a1, a0, a1)).sum(axis=1)
            avail[a0:a1] |= (s != 0).astype(np.uint8)
            s = nan0(self.mat.window(c, i0, i1, j0, j1)).sum(axis=1)
            live[i0:i1] |= (s != 0).astype(np.uint8)
        return avail, live

    def background(self, cid, spec):
        c = self._name(cid)
        specs = [tuple(int(v) for v in row) for row in np.asarray(spec).reshape(-1, 6)]
        lu, ru, ld, rd = bgm.background_columns(self.mat, c, specs)
        return np.stack([lu, ru, ld, rd])

    def set_background(self, tables):
        self.bg = tuple(np.asarray(t) for t in tables)

    def extract(self, cid, mp, sigma, minH, maxW, bfilter, gray_variant=0, frame_lo=0, frame_hi=None,
                weights=None):
        c = self._name(cid)
        mat = self.mat
        nfr = pl.n_frames(mat.nbins(c))
        frame_hi = nfr if frame_hi is None else frame_hi
        rows = []
        w = cn.gaussian_weights(sigma)
        for qi, M in enumerate(mp):
            for idx in range(frame_lo, frame_hi):
                fr = pl.load_frame(mat, c, idx)
                if fr is None:
                    continue
                D, sa, ea, S = fr["D"], fr["sa"], fr["ea"], fr["S"]
                G = im.saturate(D, M)
                bleed = cn.bleed_over((S, S), *w)
                for bi, b in enumerate(im.BRIGHTNESS):
                    gray = im.gray_image(G, b, bfilter, gray_variant)
                    edges = cn.canny(gray, sigma, w, bleed)
                    vert = sg.vertical_line(edges)
                    cols, ends, ud = sg.candidate_columns(vert, minH)
                    for d in (1, 2):
                        sel = ud == d
                        T = sg.build_testmat(S, cols[sel], ends[sel], d, edges, vert)
                        cnt = (T == 1).sum(axis=0)
                        keep = np.nonzero(cnt >= 3)[0]
                        X = sg.group_columns([int(v) for v in keep], [np.int64(cnt[v]) for v in keep])
                        for k, (x, y, ww, hh) in enumerate(sg.pair_columns(T, X, d, maxW)):
                            tot = float(D[y:y + hh, x:x + ww].sum())
                            rows.append((qi, idx, bi, d, k, x, y, ww, hh, fr["start"], fr["end"], S,
                                         int(sa[x]), int(ea[x + ww - 1]), int(sa[y]), int(ea[y + hh - 1]), tot))
        return np.array(rows, dtype=CANDIDATE_DTYPE)

    def score(self, cid, pos, expected=None, want_pvalue=True, stale=None):
        c = self._name(cid)
        pos = np.asarray(pos).reshape(-1, 4)
        n = len(pos)
        table = {"chr": [c] * n, "pos1": pos[:, 0].tolist(), "pos2": pos[:, 1].tolist(),
                 "pos3": pos[:, 2].tolist(), "pos4": pos[:, 3].tolist()}
        res = {k: np.full(n, np.nan) for k in ("pvalue", "stripiness", "oe_mean", "oe_total", "o_mean", "o_sum")}
        if want_pvalue:
            res["pvalue"] = np.array(sc.pvalues(self.mat, self.bg, table), dtype=float)
        if expected is not None:
            s, m, t = sc.stripiness(self.mat, table, {c: expected}, self.mask, on_index_error="nan")
            res["stripiness"], res["oe_mean"], res["oe_total"] = (np.array(v, dtype=float) for v in (s, m, t))
            om, osum = sc.observed_mean(self.mat, table)
            res["o_mean"], res["o_sum"] = np.array(om, dtype=float), np.array(osum, dtype=float)
        return res

    def score_multi(self, chrom_ids, chrom_index, pos, expected=None, want_pvalue=True, stale=None):
        pos = np.asarray(pos).reshape(-1, 4)
        index = np.asarray(chrom_index)
        res = {k: np.full(len(pos), np.nan) for k in ("pvalue", "stripiness", "oe_mean", "oe_total", "o_mean", "o_sum")}
        for k, cid in enumerate(chrom_ids):
            idx = np.nonzero(index == k)[0]
            if not len(idx):
                continue
            part = self.score(int(cid), pos[idx], None if expected is None else np.asarray(expected)[k], want_pvalue)
            for name, v in part.items():
                res[name][idx] = v
        return res
