"""Thin object wrapper over the C ABI: one :class:`Device` per GPU.

Replaces the reference's only data-access object, the cooler matrix selector
``Lib.matrix(balance=norm)`` (stripenn.py:118): chromosomes are uploaded once (COO pixels +
weights) and live as diagonal bands in HBM; every numerical stage is a call into
libstripenn_b200.so.  Nothing here computes on the CPU.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Dict, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import CANDIDATE_DTYPE, NDIST, check, ptr


def default_band_width(binsize: int) -> int:
    """Offsets the band must hold: 400-bin windows, the +-bs flanks of the scoring windows and
    the 3*bs-row strips of the background sampler.  Farther pixels stay reachable through the
    CSR fallback inside the kernels."""
    bs = int(50000 / binsize)
    return (400 + 3 * bs + 16 + 7) // 8 * 8


def gaussian_weights(sigma: float, truncate: float = 4.0):
    """scipy.ndimage._filters._gaussian_kernel1d(sigma, 0, int(truncate*sigma + 0.5)), evaluated
    with numpy exactly like scipy does, so the device convolves with the very same bits
    (skimage.feature.canny -> gaussian_filter, called at getStripe.py:917)."""
    sd = float(sigma)
    radius = int(truncate * sd + 0.5)
    sigma2 = sd * sd
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / sigma2 * x ** 2)
    phi = phi / phi.sum()
    return np.ascontiguousarray(phi, dtype=np.float64), radius


def unique_maxpixels(mp):
    """Distinct maxpixel values in first-appearance order and, per input value, the index of its
    representative (bitwise equality: NaN and -0.0 are not merged with anything else)."""
    bits = np.ascontiguousarray(mp, dtype=np.float64).view(np.uint64)
    first = {}
    inverse = np.empty(len(bits), dtype=np.int64)
    for i, b in enumerate(bits.tolist()):
        inverse[i] = first.setdefault(b, len(first))
    uniq = np.empty(len(first), dtype=np.float64)
    for b, k in first.items():
        uniq.view(np.uint64)[k] = b
    return uniq, inverse


def expand_duplicate_maxpixels(cand, inverse):
    """Candidates of the distinct maxpixel values (sorted by q, then emission order) -> candidates
    of every requested value: block of representative inverse[i] copied with q = i, blocks in q
    order — the order snn_extract itself emits."""
    blocks = []
    q = cand["q"]
    for i, u in enumerate(inverse.tolist()):
        b = cand[q == u].copy()
        b["q"] = i
        blocks.append(b)
    return np.concatenate(blocks) if blocks else cand


class Device:
    def __init__(self, device: int = 0, timing: bool = False):
        self.lib = _lib.load()
        h = C.c_void_p()
        check(self.lib.snn_create(int(device), C.byref(h)))
        self.h = h
        self.device = int(device)
        self.loaded: Dict[int, dict] = {}
        self._keep: Dict[int, tuple] = {}   # host arrays of asynchronous uploads still in flight
        self.h2d_bytes = 0   # host<->device payload bytes moved by this object's calls
        self.d2h_bytes = 0
        if timing:
            self.enable_timing(True)

    # ------------------------------------------------------------------ life cycle
    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.lib.snn_destroy(self.h)    # synchronises both streams
            self.h = None
            self._keep = {}

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def set_stream(self, cuda_stream: Optional[int]):
        check(self.lib.snn_set_stream(self.h, C.c_void_p(cuda_stream or 0)), self.h)

    def synchronize(self):
        check(self.lib.snn_synchronize(self.h), self.h)

    # ------------------------------------------------------------------ a1
    def load_chrom(self, chrom_id: int, nbins: int, binsize: int, chromsize: int, bin1, bin2, count,
                   weight=None, band_width: Optional[int] = None, async_: bool = False):
        """async_=True: snn_load_chrom_async — the copies run on the context's copy stream while
        this call returns; the arrays are kept alive here until the chromosome is unloaded."""
        bw = int(band_width or default_band_width(binsize))
        b1 = np.ascontiguousarray(bin1, dtype=np.int32)
        b2 = np.ascontiguousarray(bin2, dtype=np.int32)
        ct = np.ascontiguousarray(count, dtype=np.int32)
        w = None if weight is None else np.ascontiguousarray(weight, dtype=np.float64)
        if w is not None and w.shape[0] != nbins:
            raise ValueError("weight vector has %d entries, chromosome has %d bins" % (w.shape[0], nbins))
        fn = self.lib.snn_load_chrom_async if async_ else self.lib.snn_load_chrom
        check(fn(self.h, int(chrom_id), int(nbins), int(binsize), int(chromsize),
                 int(b1.shape[0]), ptr(b1), ptr(b2), ptr(ct), ptr(w), bw), self.h)
        if async_:
            self._keep[int(chrom_id)] = (b1, b2, ct, w)
        self.h2d_bytes += b1.nbytes * 3 + (0 if w is None else w.nbytes)
        self.loaded[int(chrom_id)] = dict(nbins=int(nbins), binsize=int(binsize),
                                          chromsize=int(chromsize), bw=min(bw, int(nbins)),
                                          nnz=int(b1.shape[0]))

    def load_chrom_device(self, chrom_id, nbins, binsize, chromsize, nnz, d_bin1, d_bin2, d_count,
                          d_weight=None, band_width=None):
        """Pixels already resident in HBM (raw device pointers as ints)."""
        bw = int(band_width or default_band_width(binsize))
        check(self.lib.snn_load_chrom_device(self.h, int(chrom_id), int(nbins), int(binsize),
                                             int(chromsize), int(nnz), C.c_void_p(d_bin1),
                                             C.c_void_p(d_bin2), C.c_void_p(d_count),
                                             C.c_void_p(d_weight or 0), bw), self.h)
        self.loaded[int(chrom_id)] = dict(nbins=int(nbins), binsize=int(binsize),
                                          chromsize=int(chromsize), bw=min(bw, int(nbins)), nnz=int(nnz))

    def unload_chrom(self, chrom_id: int):
        check(self.lib.snn_unload_chrom(self.h, int(chrom_id)), self.h)
        self.loaded.pop(int(chrom_id), None)
        if int(chrom_id) in self._keep:
            self.synchronize()               # the copy may still be reading the host arrays
            self._keep.pop(int(chrom_id), None)

    def fetch(self, chrom_id, i0, i1, j0, j1):
        out = np.empty((int(i1 - i0), int(j1 - j0)), dtype=np.float64)
        check(self.lib.snn_fetch(self.h, int(chrom_id), int(i0), int(i1), int(j0), int(j1), ptr(out)), self.h)
        return out

    def heatmap(self, chrom_id, i0, i1, maxpixel_value, want_channel=True):
        """seeimage's image (seeimage.py:77-90) of bins [i0,i1): (float64 green/blue channel or
        None, uint8 RGB)."""
        n = max(int(i1 - i0), 0)        # a reversed window is the library's error to report
        chan = np.empty((n, n), dtype=np.float64) if want_channel else None
        rgb = np.empty((n, n, 3), dtype=np.uint8)
        check(self.lib.snn_heatmap(self.h, int(chrom_id), int(i0), int(i1), float(maxpixel_value),
                                   ptr(chan) if want_channel else None, ptr(rgb)), self.h)
        return chan, rgb

    # ------------------------------------------------------------------ a2 / a3
    def quantiles(self, chrom_id, qs: Sequence[float]):
        q = np.ascontiguousarray(qs, dtype=np.float64)
        out = np.empty(q.shape[0], dtype=np.float64)
        check(self.lib.snn_quantiles(self.h, int(chrom_id), int(q.shape[0]), ptr(q), ptr(out)), self.h)
        return out

    def quantiles_multi(self, chrom_ids, qs: Sequence[float]):
        """snn_quantiles for several resident chromosomes in one call -> (len(chrom_ids), len(qs))."""
        ids = np.ascontiguousarray(chrom_ids, dtype=np.int32)
        q = np.ascontiguousarray(qs, dtype=np.float64)
        out = np.empty((ids.shape[0], q.shape[0]), dtype=np.float64)
        if ids.shape[0]:
            check(self.lib.snn_quantiles_multi(self.h, int(ids.shape[0]), ptr(ids), int(q.shape[0]), ptr(q), ptr(out)),
                  self.h)
        return out

    def quantiles_select(self, chrom_ids, jobs, sel, njob, qs: Sequence[float]):
        """snn_quantiles_select (the -s populations of getQuantile_slow) -> (njob, len(qs))."""
        ids = np.ascontiguousarray(chrom_ids, dtype=np.int32)
        jb = np.ascontiguousarray(jobs, dtype=np.int32)
        sl = np.ascontiguousarray(sel, dtype=np.int64).reshape(-1, 5)
        q = np.ascontiguousarray(qs, dtype=np.float64)
        out = np.empty((int(njob), q.shape[0]), dtype=np.float64)
        check(self.lib.snn_quantiles_select(self.h, int(ids.shape[0]), ptr(ids), ptr(jb), ptr(sl), int(njob),
                                            int(q.shape[0]), ptr(q), ptr(out)), self.h)
        return out

    def expected(self, chrom_id):
        out = np.empty(NDIST, dtype=np.float64)
        check(self.lib.snn_expected(self.h, int(chrom_id), ptr(out)), self.h)
        return out

    # ------------------------------------------------------------------ a4
    def row_flags(self, chrom_id, geom):
        g = np.ascontiguousarray(geom, dtype=np.int64).reshape(-1, 6)
        nb = self.loaded[int(chrom_id)]["nbins"]
        avail = np.zeros(nb, dtype=np.uint8)
        live = np.zeros(nb, dtype=np.uint8)
        check(self.lib.snn_row_flags(self.h, int(chrom_id), int(g.shape[0]), ptr(g), ptr(avail), ptr(live)),
              self.h)
        self.d2h_bytes += 2 * nb
        return avail, live

    def background(self, chrom_id, spec):
        s = np.ascontiguousarray(spec, dtype=np.int32).reshape(-1, 6)
        n = int(s.shape[0])
        out = np.empty((4, NDIST, n), dtype=np.float64)
        if n:
            check(self.lib.snn_background(self.h, int(chrom_id), n, ptr(s), ptr(out)), self.h)
            self.h2d_bytes += s.nbytes
            self.d2h_bytes += out.nbytes
        return out

    def set_background(self, tables):
        """tables: 4 arrays (400 x ncol): bgleft_up, bgright_up, bgleft_down, bgright_down."""
        t = np.ascontiguousarray(np.stack([np.asarray(x, dtype=np.float64) for x in tables]))
        assert t.shape[0] == 4 and t.shape[1] == NDIST
        check(self.lib.snn_set_background(self.h, int(t.shape[2]), ptr(t)), self.h)
        self.h2d_bytes += t.nbytes

    # device-resident background (sharded pipeline: the columns never visit the host)
    BG_SLOT = 4 * NDIST          # doubles per sample slot

    def device_alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        check(self.lib.snn_device_alloc(self.h, int(nbytes), C.byref(p)), self.h)
        return int(p.value or 0)

    def device_free(self, d_ptr: int):
        check(self.lib.snn_device_free(self.h, C.c_void_p(int(d_ptr))), self.h)

    def device_read(self, d_ptr: int, shape, dtype=np.float64):
        out = np.empty(shape, dtype=dtype)
        check(self.lib.snn_device_read(self.h, ptr(out), C.c_void_p(int(d_ptr)), out.nbytes), self.h)
        self.d2h_bytes += out.nbytes
        return out

    def background_dev(self, chrom_id, spec, d_buf: int, slot0: int) -> int:
        """snn_background_dev: the samples go to slots slot0.. of the device buffer (sample-major);
        returns the number of samples written."""
        s = np.ascontiguousarray(spec, dtype=np.int32).reshape(-1, 6)
        n = int(s.shape[0])
        if n:
            check(self.lib.snn_background_dev(self.h, int(chrom_id), n, ptr(s), C.c_void_p(int(d_buf)), int(slot0)),
                  self.h)
            self.h2d_bytes += s.nbytes
        return n

    def set_background_dev(self, d_buf: int, ncol: int, slots=None):
        sl = None if slots is None else np.ascontiguousarray(slots, dtype=np.int32)
        if sl is not None and sl.shape[0] != int(ncol):
            raise ValueError("set_background_dev: %d slots for %d columns" % (sl.shape[0], ncol))
        check(self.lib.snn_set_background_dev(self.h, int(ncol), ptr(sl), C.c_void_p(int(d_buf))), self.h)
        self.h2d_bytes += 0 if sl is None else sl.nbytes

    # ------------------------------------------------------------------ a5-a14
    def extract(self, chrom_id, maxpixel_values, sigma, minH, maxW, bfilter, gray_variant=0,
                frame_lo=0, frame_hi=None, weights=None):
        """snn_extract for several maxpixel values.  Equal values (common with raw counts: the 0.95
        and 0.96 quantiles of small integers often coincide) give identical images, so the image
        chain runs once per DISTINCT value and the candidates are replicated (same rows, same
        order as separate passes would emit; STRIPENN_B200_NO_MP_DEDUP=1 disables it)."""
        mp_all = np.ascontiguousarray(maxpixel_values, dtype=np.float64)
        uniq, inverse = unique_maxpixels(mp_all)
        if len(uniq) < len(mp_all) and not os.environ.get("STRIPENN_B200_NO_MP_DEDUP"):
            out = self._extract(chrom_id, uniq, sigma, minH, maxW, bfilter, gray_variant, frame_lo, frame_hi, weights)
            return expand_duplicate_maxpixels(out, inverse)
        return self._extract(chrom_id, mp_all, sigma, minH, maxW, bfilter, gray_variant, frame_lo, frame_hi, weights)

    def _extract(self, chrom_id, maxpixel_values, sigma, minH, maxW, bfilter, gray_variant=0,
                 frame_lo=0, frame_hi=None, weights=None):
        mp = np.ascontiguousarray(maxpixel_values, dtype=np.float64)
        gw, radius = weights if weights is not None else gaussian_weights(sigma)
        nb = self.loaded[int(chrom_id)]["nbins"]
        nframes = math.ceil(nb / 200)
        if frame_hi is None:
            frame_hi = nframes
        cap = max(4096, (frame_hi - frame_lo) * mp.shape[0] * 64)
        while True:
            out = np.empty(cap, dtype=CANDIDATE_DTYPE)
            n = C.c_int64(0)
            rc = self.lib.snn_extract(self.h, int(chrom_id), int(mp.shape[0]), ptr(mp), ptr(gw), int(radius),
                                      int(minH), int(maxW), int(bfilter), int(gray_variant), int(frame_lo),
                                      int(frame_hi), ptr(out), cap, C.byref(n))
            if rc == _lib.ECAP:
                cap = int(n.value)
                continue
            check(rc, self.h)
            self.d2h_bytes += int(n.value) * CANDIDATE_DTYPE.itemsize
            return out[:n.value]

    def debug_stages(self, chrom_id, maxpixel_value, sigma, bfilter, frame, bright, gray_variant=0):
        gw, radius = gaussian_weights(sigma)
        S = C.c_int32(0)
        nz = np.empty(400, dtype=np.int32)
        gray = np.zeros(160000, dtype=np.float32)
        smoothed = np.zeros(160000, dtype=np.float32)
        mag = np.zeros(160000, dtype=np.float32)
        low = np.zeros(160000, dtype=np.uint8)
        edges = np.zeros(160000, dtype=np.uint8)
        vert = np.zeros(160000, dtype=np.uint8)
        check(self.lib.snn_debug_stages(self.h, int(chrom_id), float(maxpixel_value), ptr(gw), int(radius),
                                        int(bfilter), int(gray_variant), int(frame), int(bright),
                                        C.byref(S), ptr(nz), ptr(gray), ptr(smoothed), ptr(mag), ptr(low),
                                        ptr(edges), ptr(vert)), self.h)
        s = int(S.value)
        cut = lambda a: a[:s * s].reshape(s, s).copy()
        return dict(S=s, nz=nz[:s].copy(), gray=cut(gray), smoothed=cut(smoothed), mag=cut(mag),
                    low=cut(low), edges=cut(edges), vert=cut(vert))

    # ------------------------------------------------------------------ a16-a18
    def score(self, chrom_id, pos, expected=None, want_pvalue=True, stale=None):
        p = np.ascontiguousarray(pos, dtype=np.int64).reshape(-1, 4)
        n = int(p.shape[0])
        res = {k: np.full(n, np.nan) for k in ("pvalue", "stripiness", "oe_mean", "oe_total", "o_mean", "o_sum")}
        if n == 0:
            return res
        ex = None if expected is None else np.ascontiguousarray(expected, dtype=np.float64)
        st = None if stale is None else np.ascontiguousarray(stale, dtype=np.int32).reshape(-1, 2)
        check(self.lib.snn_score(self.h, int(chrom_id), n, ptr(p), ptr(st), ptr(ex),
                                 ptr(res["pvalue"]) if want_pvalue else None,
                                 ptr(res["stripiness"]) if ex is not None else None,
                                 ptr(res["oe_mean"]) if ex is not None else None,
                                 ptr(res["oe_total"]) if ex is not None else None,
                                 ptr(res["o_mean"]) if ex is not None else None,
                                 ptr(res["o_sum"]) if ex is not None else None), self.h)
        self.h2d_bytes += p.nbytes + (0 if ex is None else ex.nbytes)
        self.d2h_bytes += n * 48
        return res

    def score_multi(self, chrom_ids, chrom_index, pos, expected=None, want_pvalue=True, stale=None):
        """snn_score for stripes of several resident chromosomes in one call: chrom_ids = the
        chromosomes involved, chrom_index[i] = index into chrom_ids of stripe i, expected =
        (len(chrom_ids), 400) mpmean rows or None."""
        p = np.ascontiguousarray(pos, dtype=np.int64).reshape(-1, 4)
        n = int(p.shape[0])
        res = {k: np.full(n, np.nan) for k in ("pvalue", "stripiness", "oe_mean", "oe_total", "o_mean", "o_sum")}
        if n == 0:
            return res
        ids = np.ascontiguousarray(chrom_ids, dtype=np.int32)
        ci = np.ascontiguousarray(chrom_index, dtype=np.int32)
        assert ci.shape[0] == n
        ex = None if expected is None else np.ascontiguousarray(expected, dtype=np.float64).reshape(len(ids), NDIST)
        st = None if stale is None else np.ascontiguousarray(stale, dtype=np.int32).reshape(-1, 2)
        check(self.lib.snn_score_multi(self.h, int(len(ids)), ptr(ids), ptr(ex), n, ptr(ci), ptr(p), ptr(st),
                                       ptr(res["pvalue"]) if want_pvalue else None,
                                       ptr(res["stripiness"]) if ex is not None else None,
                                       ptr(res["oe_mean"]) if ex is not None else None,
                                       ptr(res["oe_total"]) if ex is not None else None,
                                       ptr(res["o_mean"]) if ex is not None else None,
                                       ptr(res["o_sum"]) if ex is not None else None), self.h)
        self.h2d_bytes += p.nbytes + ci.nbytes + (0 if ex is None else ex.nbytes)
        self.d2h_bytes += n * 48
        return res

    def set_mask(self, chrom_id: Optional[int], start_bp: int = 0, end_bp: int = 0):
        """--mask of scoringstripes (getStripe.py:612-639) for the following score calls; None clears it."""
        check(self.lib.snn_set_mask(self.h, -1 if chrom_id is None else int(chrom_id), int(start_bp), int(end_bp)),
              self.h)

    # ------------------------------------------------------------------ timers
    def enable_timing(self, on=True):
        check(self.lib.snn_enable_stage_timing(self.h, 1 if on else 0), self.h)

    def reset_stage_times(self):
        check(self.lib.snn_reset_stage_times(self.h), self.h)

    def stage_times(self):
        cap = 32
        names = (C.c_char_p * cap)()
        ms = (C.c_double * cap)()
        launches = (C.c_int64 * cap)()
        n = self.lib.snn_stage_times(self.h, cap, names, ms, launches)
        if n < 0:
            check(n, self.h)
        return {names[i].decode(): dict(ms=ms[i], launches=int(launches[i])) for i in range(n)}


def _score_routing(self, reset=True):
    """(fast, warp, CTA) stripe counts of the scoring calls since the last reset."""
    out = (C.c_int64 * 3)()
    check(self.lib.snn_score_routing(self.h, out, 1 if reset else 0), self.h)
    return tuple(int(v) for v in out)


Device.score_routing = _score_routing


def remove_redundant_mask(group, num, pos, h, w, key, by: str):
    """getStripe.RemoveRedundant keep-mask (host C++ in the same library)."""
    lib = _lib.load()
    code = {"size": 0, "pvalue": 1, "score": 2}.get(by)
    if code is None:
        raise ValueError('"by" should be one of "size", "pvalue" and "score"')
    p = np.ascontiguousarray(pos, dtype=np.int64).reshape(-1, 4)
    n = int(p.shape[0])
    keep = np.ones(n, dtype=np.uint8)
    if n == 0:
        return keep.astype(bool)
    g = np.ascontiguousarray(group, dtype=np.int32)
    nm = np.ascontiguousarray(num, dtype=np.int64)
    hh = None if h is None else np.ascontiguousarray(h, dtype=np.float64)
    ww = None if w is None else np.ascontiguousarray(w, dtype=np.float64)
    kk = None if key is None else np.ascontiguousarray(key, dtype=np.float64)
    rc = lib.snn_remove_redundant(n, ptr(g), ptr(nm), ptr(p), ptr(hh), ptr(ww), ptr(kk), code, ptr(keep))
    if rc != 0:
        raise _lib.StripennB200Error(rc, "snn_remove_redundant: bad arguments")
    return keep.astype(bool)
