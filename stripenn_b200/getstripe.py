"""Host-side mirror of the reference class ``getStripe`` (src/stripenn/getStripe.py:17-33).

Same constructor arguments (the cooler matrix selector is replaced by a :class:`DeviceMatrix`,
the contact map resident on one B200), same method names, argument meaning and return types,
so ``compute``/``getScore`` read like the reference drivers.  Every method hands its arithmetic
to the CUDA library through :class:`stripenn_b200.engine.Device`; what stays on the host is
exactly what the north star assigns to it: the ``random.Random`` sample draws, the table
bookkeeping and file I/O.
"""
from __future__ import annotations

import math
import os
import random
from typing import Dict, List, Optional, Sequence

import numpy as np
import pandas as pd

from .contactmap import ContactMap
from ._lib import CANDIDATE_DTYPE  # noqa: F401 (re-exported)
from .engine import Device, gaussian_weights, remove_redundant_mask  # noqa: F401 (re-exported)

NDIST = 400
RESULT_COLUMNS = ['chr', 'pos1', 'pos2', 'chr2', 'pos3', 'pos4', 'length', 'width', 'total', 'Mean',
                  'maxpixel', 'num', 'start', 'end', 'x', 'y', 'h', 'w', 'medpixel']


class DeviceMatrix:
    """What ``Lib.matrix(balance=norm)`` is to the reference (stripenn.py:118): the balanced
    contact matrix, here resident in HBM.  Chromosomes are uploaded lazily, once."""

    def __init__(self, cm: ContactMap, balance, device: Optional[Device] = None, device_id: int = 0,
                 band_width: Optional[int] = None):
        self.cm = cm
        self.balance = balance
        self.dev = device if device is not None else Device(device_id)
        self.band_width = band_width
        self.binsize = int(cm.binsize)

    def chrom_id(self, chrom: str) -> int:
        return self.cm.chromnames.index(str(chrom))

    def chromsize(self, chrom: str) -> int:
        return int(self.cm.chromsizes[self.chrom_id(chrom)])

    def nbins(self, chrom: str) -> int:
        return int(math.ceil(self.chromsize(chrom) / self.binsize))

    def ensure(self, chrom: str, band_width: Optional[int] = None) -> int:
        cid = self.chrom_id(chrom)
        need = band_width or self.band_width
        have = self.dev.loaded.get(cid)
        if have is not None and (need is None or have["bw"] >= min(need, have["nbins"])):
            return cid
        self._upload(chrom, cid, need, False)
        return cid

    def prefetch(self, chroms) -> None:
        """Start the uploads of several chromosomes (snn_load_chrom_async): the copies overlap the
        kernels of whatever is used first; a chromosome's band is built when it is first used."""
        for chrom in chroms:
            cid = self.chrom_id(chrom)
            if cid not in self.dev.loaded:
                self._upload(chrom, cid, self.band_width, True)

    def _upload(self, chrom, cid, need, async_):
        if str(chrom) not in self.cm.chroms:
            raise KeyError("chromosome %r has no pixels in the contact map" % chrom)
        c = self.cm.chroms[str(chrom)]
        w = self.cm.weight_vector(str(chrom), self.balance)
        self.dev.load_chrom(cid, c.nbins, self.binsize, c.size_bp, c.bin1, c.bin2, c.count, w, need, async_=async_)


def _bins_of(start: int, end: int, size: int, r: int):
    """cooler region -> bin range, with cooler's range checks."""
    if end < start:
        raise ValueError("End cannot be less than start")
    if start < 0 or end > size:
        raise ValueError("Genomic region out of bounds: [%d, %d)" % (start, end))
    return start // r, int(math.ceil(end / r))


class getStripe:
    def __init__(self, unbalLib: DeviceMatrix, resol, minH, maxW, canny, all_chromnames, chromnames,
                 all_chromsizes, chromsizes, core, bfilter, seed, gray_variant: int = 0):
        self.unbalLib = unbalLib
        self.dev = unbalLib.dev
        self.resol = int(resol)
        self.minH = minH
        self.maxW = maxW
        self.canny = canny
        self.all_chromnames = list(all_chromnames)
        self.all_chromsizes = np.asarray(all_chromsizes)
        self.chromnames = list(chromnames)
        self.chromsizes = np.asarray(chromsizes)
        self.core = core          # kept for signature compatibility; the GPU does the work
        self.bfilter = bfilter
        self.seed = seed
        self.gray_variant = gray_variant
        self.chromnames2sizes = {self.all_chromnames[i]: int(self.all_chromsizes[i])
                                 for i in range(len(self.all_chromnames))}
        self._bg_installed = None

    def prefetch(self, chroms):
        """Begin uploading the chromosomes the following stages will touch (no-op for matrix
        objects without asynchronous uploads)."""
        fn = getattr(self.unbalLib, "prefetch", None)
        if fn is not None and not os.environ.get("SNN_NO_PREFETCH"):   # A/B switch: uploads at first use
            fn(list(chroms))

    # ------------------------------------------------------------------ a2
    def getQuantile_original(self, coolinfo, ChrList, quantile):
        """getStripe.py:160-176 -> {chrom: ndarray[len(quantile)]}."""
        ChrList = list(ChrList)
        ids = [self.unbalLib.ensure(c) for c in ChrList]
        vals = self.dev.quantiles_multi(ids, quantile)      # one device call for the whole list
        return {c: vals[i] for i, c in enumerate(ChrList)}

    def slow_selection(self, ChrList):
        """The populations getQuantile_slow (getStripe.py:107-158) builds, as (chromosomes, sources):
        sources = [(source chromosome, job, s, w, F, c0, c1)] in the terms of snn_quantiles_select.
        The reference slices the selector in GENOME-WIDE bins: rows r_start + [k*w, min((k+1)*w-1,
        floor(size/r))) for k < ceil(L/w), w = floor(25e6/L); columns [r_start, r_end) with r_start =
        ceil(cumsize_before/r) + 1, r_end = ceil(cumsize/r) — not the chromosome's own bin offset
        sum(ceil(size_j/r)), so the first bin is skipped and the window can begin in the previous
        chromosome(s).  Pixels between different chromosomes are not part of a ContactMap (cis
        only), like in the oracle's selector."""
        cm = self.unbalLib.cm
        names = list(cm.chromnames)
        sizes = np.asarray(cm.chromsizes)
        nbin = cm.binsize
        cum = np.nancumsum(sizes)
        nb = [int(np.ceil(sz / nbin)) for sz in sizes]
        offs = np.concatenate([[0], np.cumsum(nb)]).astype(np.int64)
        chridx = sorted(c for c in range(len(names)) if names[c] in ChrList)
        jobs, sources = [], []
        for ci in chridx:
            CHROMSIZE = sizes[ci]
            L = int(np.ceil(CHROMSIZE / nbin))
            r_start = 0 if ci == 0 else cum[ci - 1]
            r_end = cum[ci]
            r_start = int(np.ceil(r_start / nbin) + 1)
            r_end = int(np.ceil(r_end / nbin))
            w = int(np.floor(25000000 / L))
            F = int(np.floor(CHROMSIZE / nbin))
            job = len(jobs)
            jobs.append(names[ci])
            for p in range(ci + 1):
                if offs[p + 1] <= r_start and p != ci:
                    continue        # the window starts after this chromosome
                if r_end <= offs[p]:
                    continue
                if names[p] not in cm.chroms:
                    continue        # no pixels stored for it
                sources.append((names[p], job, r_start - int(offs[p]), w, F, r_start - int(offs[p]),
                                r_end - int(offs[p])))
        return jobs, sources

    def getQuantile_slow(self, coolinfo, ChrList, quantile):
        """getStripe.py:107-158 (-s) -> {chrom: ndarray[len(quantile)]}.  NOT the values of
        getQuantile_original: see slow_selection."""
        jobs, sources = self.slow_selection(list(ChrList))
        ids = [self.unbalLib.ensure(s[0]) for s in sources]
        sel = np.array([s[2:] for s in sources], dtype=np.int64).reshape(-1, 5)
        vals = self.dev.quantiles_select(ids, [s[1] for s in sources], sel, len(jobs), quantile)
        return {c: vals[i] for i, c in enumerate(jobs)}

    # ------------------------------------------------------------------ a3
    def mpmean(self):
        """getStripe.py:178-235 -> {chrom: list[400]} for the selected chromosomes."""
        return {c: self.expected_values(c) for c in self.chromnames}

    def expected_values(self, chrom):
        """``calc`` of mpmean (getStripe.py:180-209) for one chromosome -> list[400]."""
        return self.dev.expected(self.unbalLib.ensure(chrom)).tolist()

    # ------------------------------------------------------------------ a4
    @staticmethod
    def _budget(values):
        v = (np.asarray(values, dtype=np.float64) / np.sum(values)) * 1000
        v = v.astype(np.uint64)
        dif = 1000 - int(np.sum(v))
        nz = np.nonzero(v != 0)[0]
        v[0] += np.uint64(dif)
        return v, nz

    def _units(self, chrom):
        r = self.resol
        csize = self.chromnames2sizes[chrom]
        itera = int(min(csize / r / 500, 25))
        if itera == 0:
            # fewer than 500 bins: available_cols (getStripe.py:249-273) divides by np.uint64(0)
            # (inf under errstate) and loops over range(0) -> poolsum 0, so the chromosome drops
            # out of the sample budget and main_null_calc never sees it
            return csize, 0, 0
        unit = int(np.floor(csize / r / itera))
        return csize, itera, unit

    def _unit_geometry(self, chrom, it, unit):
        """Row/column bin ranges of the window main_null_calc fetches (getStripe.py:313-326)."""
        r = self.resol
        csize = self.chromnames2sizes[chrom]
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
        i0, i1 = _bins_of(start1, end1, csize, r)
        j0, j1 = _bins_of(start0, end2, csize, r)
        return i0, i1, j0, j1

    def _flags(self, chrom):
        """Per-bin flags from the device: avail (non-empty row inside the unit's own block,
        available_cols :249-273) and live (non-empty row of the unit's sampling window :328-331)."""
        r = self.resol
        csize, itera, unit = self._units(chrom)
        if itera == 0:
            nb = int(math.ceil(csize / r))
            return [], np.zeros(nb, dtype=np.uint8), np.zeros(nb, dtype=np.uint8)
        geom = []
        for it in range(itera):
            a0, a1 = _bins_of(unit * r * it + 1, unit * r * (it + 1), csize, r)
            i0, i1, j0, j1 = self._unit_geometry(chrom, it, unit)
            geom.append((a0, a1, i0, i1, j0, j1))
        cid = self.unbalLib.ensure(chrom)
        avail, live = self.dev.row_flags(cid, np.array(geom, dtype=np.int64))
        return geom, avail, live

    def background_chromosomes(self):
        """chromnames2 after the size budget (getStripe.py:242-247): chromosomes whose share of
        1000 samples by size is non-zero."""
        _, nz1 = self._budget(self.all_chromsizes)
        return [self.all_chromnames[i] for i in nz1]

    def available_counts(self, chroms):
        """available_cols (getStripe.py:249-273) for some chromosomes -> {chrom: count}.  The
        flags stay cached for plan_chromosome."""
        if not hasattr(self, "_flag_cache"):
            self._flag_cache = {}
        out = {}
        for c in chroms:
            if c not in self._flag_cache:
                self._flag_cache[c] = self._flags(c)
            out[c] = int(self._flag_cache[c][1].sum())
        return out

    def sample_budget(self, names2, avail: Dict[str, int]):
        """getStripe.py:278-283 -> (chromnames2 after the availability budget, samplesize)."""
        budget, nz2 = self._budget([avail[c] for c in names2])
        return [names2[i] for i in nz2], budget

    def plan_chromosome(self, chrom, ss):
        """Sample specs of one chromosome (getStripe.py:296-346, 416-440) as an (n, 6) int32
        array {i0, i1, j0, j1, x, shift}.  RNG convention: a fresh random.Random(seed) per
        chromosome — what every loky task of the reference sees with -n >= 2."""
        prng = random.Random(self.seed)
        csize, itera, unit = self._units(chrom)
        if itera == 0:
            return np.zeros((0, 6), dtype=np.int32)
        if not hasattr(self, "_flag_cache"):
            self._flag_cache = {}
        if chrom not in self._flag_cache:
            self._flag_cache[chrom] = self._flags(chrom)
        geom, _, live = self._flag_cache[chrom]
        specs, n_pool = [], []
        it = 0

        def pool_of(g, first):
            _, _, i0, i1, j0, j1 = g
            x = np.arange(i1 - i0)
            ok = (live[i0:i1] != 0) & (x > 20) & (x < unit - 20)
            if first:
                ok &= (x > 410) & (x < (j1 - j0))
            return x[ok].tolist()

        for it in range(itera):
            sss = int(ss / itera)
            g = geom[it]
            pool = pool_of(g, it == 0)
            n_pool.append(len(pool))
            if len(pool) == 0:
                continue
            k = len(pool) if len(pool) < sss else sss
            for x in prng.choices(pool, k=k):
                specs.append((g[2], g[3], g[4], g[5], x, 400 if it > 0 else 0))
        depl = ss - len(specs)
        if depl > 0:
            rich = int(np.argmax(n_pool))
            g = geom[rich]
            pool = pool_of(g, False)
            for x in prng.choices(pool, k=depl):
                specs.append((g[2], g[3], g[4], g[5], x, 400 if it > 0 else 0))  # stale `it`
        return np.array(specs, dtype=np.int32).reshape(-1, 6)

    def plan_samples(self):
        """Everything of nulldist that decides WHICH rows are sampled."""
        names2 = self.background_chromosomes()
        avail = self.available_counts(names2)
        names2b, budget = self.sample_budget(names2, avail)
        # samplesize[c] indexes the UNfiltered budget with the filtered position (reference quirk)
        plans = [self.plan_chromosome(c, int(budget[names2b.index(c)])) for c in names2b]
        return names2b, plans

    def background_columns(self, chrom, spec):
        """main_null_calc's four tables for one chromosome -> array (4, 400, n)."""
        cid = self.unbalLib.ensure(chrom)
        return self.dev.background(cid, spec)

    def background_columns_dev(self, chrom, spec, d_buf, slot0, expect):
        """The same tables written to slots slot0.. of a device buffer (snn_background_dev): the
        form the sharded pipeline exchanges between GPUs.  `expect` = the sample budget the slot
        layout was derived from."""
        if len(spec) != expect:
            raise RuntimeError("chromosome %s: sampling plan has %d samples, budget %d" % (chrom, len(spec), expect))
        cid = self.unbalLib.ensure(chrom)
        return self.dev.background_dev(cid, spec, d_buf, slot0)

    def install_background_dev(self, d_buf, ncol, slots):
        """Genome-wide tables from a device buffer (snn_set_background_dev)."""
        self.dev.set_background_dev(d_buf, ncol, slots)
        self._bg_installed = ("dev", int(d_buf), int(ncol))
        self._bg_keepalive = None

    def nulldist(self):
        """getStripe.py:238-499 -> bgleft_up, bgright_up, bgleft_down, bgright_down."""
        names2, plans = self.plan_samples()
        parts = [self.background_columns(chrom, spec) for chrom, spec in zip(names2, plans)]
        if parts:
            full = np.concatenate(parts, axis=2)
        else:
            full = np.zeros((4, NDIST, 0))
        return full[0], full[1], full[2], full[3]

    # ------------------------------------------------------------------ a5-a15
    def _candidate_frame(self, cand, chrom_names, perc):
        """Candidate records -> DataFrame with the reference's columns (getStripe.py:1099-1110).
        `chrom_names` is one name or an array with one name per record."""
        n = len(cand)
        h = cand["h"].astype(np.int64)
        w = cand["w"].astype(np.int64)
        with np.errstate(divide="ignore", invalid="ignore"):
            mean = cand["total"] / h / w
        chrs = np.full(n, chrom_names, dtype=object) if isinstance(chrom_names, str) else np.asarray(chrom_names, dtype=object)
        label = np.full(n, str(perc * 100) + '%', dtype=object)
        df = pd.DataFrame({
            'chr': chrs, 'pos1': cand["pos1"], 'pos2': cand["pos2"], 'chr2': chrs,
            'pos3': cand["pos3"], 'pos4': cand["pos4"],
            'length': cand["pos4"] - cand["pos3"] + 1, 'width': cand["pos2"] - cand["pos1"] + 1,
            'total': cand["total"], 'Mean': mean, 'maxpixel': label,
            'num': cand["frame"].astype(np.int64), 'start': cand["start"].astype(np.int64),
            'end': cand["end"].astype(np.int64), 'x': cand["x"].astype(np.int64),
            'y': cand["y"].astype(np.int64), 'h': h, 'w': w, 'medpixel': np.full(n, np.nan)}, copy=False)
        return df

    # extract_all is composed of three steps so that a pipelined driver (compute.run_pipeline) can
    # keep the device busy with the next chromosome while the host prunes the previous one:
    def search_candidates(self, chrom, mp_values, frame_range=None):
        """Device part of ``extract`` for one chromosome and all maxpixel values: raw candidate
        records of every window (StripeSearch, getStripe.py:864-1114), before any RemoveRedundant."""
        cid = self.unbalLib.ensure(chrom)
        lo, hi = (0, None) if frame_range is None else frame_range
        return self.dev.extract(cid, np.asarray(mp_values, dtype=np.float64), self.canny, self.minH, self.maxW,
                                self.bfilter, self.gray_variant, lo, hi, gaussian_weights(self.canny))

    @staticmethod
    def prune_candidates(cand):
        """Host part: the per-window (getStripe.py:1112) and the cross-window (:852)
        RemoveRedundant('size') of every maxpixel value in one call each (pairs never cross a
        maxpixel value or a chromosome).  Returns (records, n x 4 positions)."""
        if not len(cand):
            return cand, np.zeros((0, 4), np.int64)
        pos = np.stack([cand["pos1"], cand["pos2"], cand["pos3"], cand["pos4"]], axis=1)
        hh, ww = cand["h"].astype(np.float64), cand["w"].astype(np.float64)
        frame = cand["frame"].astype(np.int64)
        nfr = int(frame.max()) + 1
        grp = (cand["q"].astype(np.int64) * nfr + frame).astype(np.int32)       # one group per (q, window)
        keep = remove_redundant_mask(grp, np.zeros(len(cand), np.int64), pos, hh, ww, None, "size")
        cand, pos, hh, ww, frame = cand[keep], pos[keep], hh[keep], ww[keep], frame[keep]
        keep = remove_redundant_mask(cand["q"], frame, pos, hh, ww, None, "size")   # one group per q
        return cand[keep], pos[keep]

    def candidate_pvalues(self, kept, bg=None):
        """p-values of every chromosome's candidates in ONE device call (all chromosomes are
        resident).  kept = [(chrom, records, positions)] with non-empty records; bg = the four host
        tables, or None when the background is already installed."""
        if not kept:
            return np.zeros(0)
        if bg is not None:
            self._install_background(bg)
        ids = np.array([self.unbalLib.ensure(k[0]) for k in kept], dtype=np.int32)
        index = np.concatenate([np.full(len(k[1]), i, dtype=np.int32) for i, k in enumerate(kept)])
        allpos = np.concatenate([k[2] for k in kept])
        return self.dev.score_multi(ids, index, allpos, None, True, self._stale_rows(allpos))["pvalue"]

    def candidate_tables(self, kept, allp, quantiles):
        """One table per maxpixel value, chromosomes in order (what the reference's per-quantile
        loop, stripenn.py:134-138, collects)."""
        out = []
        if kept:
            allc = np.concatenate([k[1] for k in kept])
            alln = np.concatenate([np.full(len(k[1]), k[0], dtype=object) for k in kept])
        for qi, perc in enumerate(quantiles):
            sel = (allc["q"] == qi) if kept else None
            if kept and sel.any():
                res = self._candidate_frame(allc[sel], alln[sel], perc)
                res['pvalue'] = allp[sel]
            else:
                res = pd.DataFrame(columns=RESULT_COLUMNS + ['pvalue'])
            out.append(res)
        return out

    def extract_all(self, MP, quantiles: Sequence[float], bgleft_up, bgright_up, bgleft_down,
                    bgright_down, frame_range=None, chroms=None):
        """``extract`` for every maxpixel value at once: the device builds each search window
        once and runs all Q x 6 brightness passes on it.  Returns one DataFrame per quantile,
        identical to what the reference's per-quantile loop (stripenn.py:134-138) collects."""
        kept = []
        for chrom in (self.chromnames if chroms is None else chroms):
            cand, pos = self.prune_candidates(self.search_candidates(chrom, MP[chrom], frame_range))
            if len(cand):
                kept.append((chrom, cand, pos))
        allp = self.candidate_pvalues(kept, (bgleft_up, bgright_up, bgleft_down, bgright_down))
        return self.candidate_tables(kept, allp, quantiles)

    def extract(self, MP, index, perc, bgleft_up, bgright_up, bgleft_down, bgright_down):
        """getStripe.py:790-862 for one maxpixel value."""
        MP1 = {c: np.asarray(MP[c], dtype=np.float64)[[index]] for c in self.chromnames}
        return self.extract_all(MP1, [perc], bgleft_up, bgright_up, bgleft_down, bgright_down)[0]

    def RemoveRedundant(self, df, by):
        """getStripe.py:1116-1196."""
        if by not in ('size', 'pvalue', 'score'):
            raise ValueError('"by" should be one of "size", "pvalue" and "score"')
        if len(df) == 0:
            return df
        codes, _ = pd.factorize(df['chr'].astype(str))
        key = None
        if by == 'pvalue':
            key = df['pvalue'].to_numpy(dtype=np.float64)
        elif by == 'score':
            key = df['Stripiness'].to_numpy(dtype=np.float64)
        keep = remove_redundant_mask(codes, df['num'].to_numpy(dtype=np.int64),
                                     df[['pos1', 'pos2', 'pos3', 'pos4']].to_numpy(dtype=np.int64),
                                     df['h'].to_numpy(dtype=np.float64), df['w'].to_numpy(dtype=np.float64),
                                     key, by)
        return df[keep]

    # ------------------------------------------------------------------ a16-a18
    def _install_background(self, tables):
        ident = tuple(id(t) for t in tables)
        if self._bg_installed != ident:
            self.dev.set_background(tables)
            self._bg_installed = ident
            self._bg_keepalive = tables

    def _by_chrom(self, df):
        chrs = df['chr'].astype(str).to_numpy()
        pos = df[['pos1', 'pos2', 'pos3', 'pos4']].to_numpy(dtype=np.int64)
        codes, uniq = pd.factorize(chrs)
        order = np.argsort(codes, kind="stable")
        bounds = np.searchsorted(codes[order], np.arange(len(uniq) + 1))
        groups = {str(uniq[k]): order[bounds[k]:bounds[k + 1]] for k in range(len(uniq))}
        return chrs, pos, groups

    def _slots(self, chrs, groups):
        """(chromosome ids, per-row index into them) for Device.score_multi; uploads what is missing."""
        ids = np.array([self.unbalLib.ensure(c) for c in groups], dtype=np.int32)
        index = np.empty(len(chrs), dtype=np.int32)
        for k, idx in enumerate(groups.values()):
            index[idx] = k
        return ids, index

    def _stale_rows(self, pos):
        """Background row inherited by stripes that are neither on nor under the diagonal
        (getStripe.py:583-597: the loop variables persist across stripes).  None when every
        stripe is on or under the diagonal (always the case for StripeSearch output)."""
        r = self.resol
        pos = np.asarray(pos, dtype=np.int64).reshape(-1, 4)
        x1, x2 = (pos[:, 0] - 1) // r, pos[:, 1] // r
        y1, y2 = (pos[:, 2] - 1) // r, pos[:, 3] // r
        down, up = x1 == y1, (x1 != y1) & (x2 == y2)
        if bool(np.all(down | up)):
            return None
        h = -(-pos[:, 3] // r) - pos[:, 2] // r
        n = len(pos)
        stale = np.full((n, 2), -1, dtype=np.int32)
        last = (-1, -1)
        for i in range(n):
            if h[i] > 0 and down[i]:
                last = (0, int(min(h[i] - 1, 399)))
            elif h[i] > 0 and up[i]:
                last = (1, int(min(max(y2[i] - y1[i] - (h[i] - 1) - 1, -400), 399) % 400))
            elif not (down[i] or up[i]):
                stale[i] = last
        return stale

    def pvalue(self, bgleft_up, bgright_up, bgleft_down, bgright_down, df):
        """getStripe.py:536-606 -> list of p-values in table order."""
        if len(df) == 0:
            return []
        self._install_background((bgleft_up, bgright_up, bgleft_down, bgright_down))
        chrs, pos, groups = self._by_chrom(df)
        stale = self._stale_rows(pos)
        ids, index = self._slots(chrs, groups)
        return self.dev.score_multi(ids, index, pos, None, True, stale)["pvalue"].tolist()

    def _apply_mask(self, mask):
        """--mask 'chr:start-end' (getStripe.py:612-618) -> device state; '0' clears it.  A mask on a
        chromosome that is not in the contact map never matches (`mask_chr == c` is False)."""
        if mask == '0':
            self.dev.set_mask(None)
            return
        parts = mask.split(':')
        mask_chr = parts[0]
        mask_start = int(parts[1].split('-')[0])
        mask_end = int(parts[1].split('-')[1])
        if mask_chr in self.unbalLib.cm.chromnames:
            self.dev.set_mask(self.unbalLib.chrom_id(mask_chr), mask_start, mask_end)
        else:
            self.dev.set_mask(None)

    def scoringstripes(self, df, expecVal, mask='0'):
        """getStripe.py:608-788 -> (Stripiness, O/E mean, O/E total) lists.  With a mask, stripes
        on which the reference's masking() raises IndexError get NaN."""
        n = len(df)
        g, m, t = np.zeros(n), np.zeros(n), np.zeros(n)
        if n == 0:
            return g.tolist(), m.tolist(), t.tolist()
        chrs, pos, groups = self._by_chrom(df)
        ids, index = self._slots(chrs, groups)
        ex = np.stack([np.asarray(expecVal[c], dtype=np.float64) for c in groups])
        self._apply_mask(mask)
        try:
            res = self.dev.score_multi(ids, index, pos, ex, False)
        finally:
            if mask != '0':
                self.dev.set_mask(None)
        return res["stripiness"].tolist(), res["oe_mean"].tolist(), res["oe_total"].tolist()

    def stripiness_of(self, parts, expecVal, mask='0'):
        """scoringstripes for [(chrom, n x 4 positions)] in one device call -> Stripiness array."""
        parts = [(c, pos) for c, pos in parts if len(pos)]
        if not parts:
            return np.zeros(0)
        ids = np.array([self.unbalLib.ensure(c) for c, _ in parts], dtype=np.int32)
        index = np.concatenate([np.full(len(pos), k, dtype=np.int32) for k, (_, pos) in enumerate(parts)])
        ex = np.stack([np.asarray(expecVal[c], dtype=np.float64) for c, _ in parts])
        self._apply_mask(mask)
        try:
            return self.dev.score_multi(ids, index, np.concatenate([pos for _, pos in parts]), ex, False)["stripiness"]
        finally:
            if mask != '0':
                self.dev.set_mask(None)

    def getMean(self, df):
        """getStripe.py:501-534 -> (nanmean, nansum) lists of the stripe bodies."""
        n = len(df)
        ms, ss = np.zeros(n), np.zeros(n)
        if n == 0:
            return ms.tolist(), ss.tolist()
        chrs, pos, groups = self._by_chrom(df)
        ids, index = self._slots(chrs, groups)
        res = self.dev.score_multi(ids, index, pos, np.zeros((len(ids), NDIST)), False)
        return res["o_mean"].tolist(), res["o_sum"].tolist()
