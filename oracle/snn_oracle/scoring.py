"""ORACLE (test infrastructure, never on the product path) — empirical p-value, stripiness
and observed mean of candidate stripes restated.

Follows reference src/stripenn/getStripe.py:536-606 (pvalue), :608-788 (scoringstripes with
expecMatrix :641-659 and iterate_idx :661-759), :501-534 (getMean) and
src/stripenn/stats.py:184-199 (elementwise_product_sum).  The ``--mask`` branch
(getStripe.py:612-639,701-707) is restated too.
"""
import numpy as np

from .matrix import nan0


def pvalues(mat, bg, table):
    """pvalue(): bg = (bgleft_up, bgright_up, bgleft_down, bgright_down)."""
    blu, bru, bld, brd = bg
    r = mat.binsize
    bs = int(50000 / r)
    out = []
    bleft = bright = None  # stale across stripes, like the reference's loop variables
    for i in range(len(table["pos1"])):
        chrom = str(table["chr"][i])
        clen = mat.chromsize(chrom)
        p1, p2, p3, p4 = (table[k][i] for k in ("pos1", "pos2", "pos3", "pos4"))
        lm = max(p1 - bs * r, 1)
        rm = min(p2 + bs * r, clen)
        w = mat.fetch_bp(chrom, int(float(p3)), int(float(p4)), int(float(lm)), int(float(rm)))
        w = w.astype(np.float64)
        center = nan0(w[:, bs:-bs]).mean(axis=1) if w[:, bs:-bs].shape[1] else np.full(w.shape[0], np.nan)
        left = nan0(w[:, :bs]).mean(axis=1)
        right = nan0(w[:, -bs:]).mean(axis=1)
        dl = center - left
        dr = center - right
        x1 = int((p1 - 1) / r); x2 = int(p2 / r)
        y1 = int((p3 - 1) / r); y2 = int(p4 / r)
        pv = []
        for j in range(len(center)):
            if x1 == y1:
                d = min(j, 399)
                bleft, bright = bld[d, :], brd[d, :]
            elif x2 == y2:
                d = min(y2 - y1 - j - 1, 399)
                bleft, bright = blu[d, :], bru[d, :]
            if bleft is None:
                raise NameError("stripe is neither on nor under the diagonal and no previous "
                                "background row exists (reference raises UnboundLocalError)")
            with np.errstate(divide="ignore", invalid="ignore"):
                a = np.count_nonzero(bleft >= dl[j]) / np.count_nonzero(~np.isnan(bleft))
                b = np.count_nonzero(bright >= dr[j]) / np.count_nonzero(~np.isnan(bright))
            p = max(a, b)
            if p == 0:
                p = 1 / len(bleft)
            pv.append(p)
        out.append(np.median(pv))
    return out


def _expected(ex, x0, x1, y0, y1):
    xs = np.arange(x0, x1)
    ys = np.arange(y0, y1)
    d = np.minimum(np.abs(xs[None, :] - ys[:, None]), 399)
    return np.asarray(ex, dtype=np.float64)[d]


def _mask_cols(m, ms, me, start, end, direc):
    """masking() (getStripe.py:620-639): the index list runs over range(end - start + 1), one
    longer than the part, and the element loops raise IndexError exactly where the reference's do
    (only when the other axis is non-empty)."""
    m = m.astype(float)
    rel0 = ms - start
    rel1 = rel0 + (me - ms)
    L = end - start + 1
    idx = [x for x in range(rel0, rel1 + 1) if 0 <= x < L]
    if idx:
        if direc == 1:
            if m.shape[0] > 0 and max(idx) >= m.shape[1]:
                raise IndexError("masking(): column %d out of bounds for axis 1 with size %d" % (max(idx), m.shape[1]))
            m[:, idx] = np.nan
        else:
            if m.shape[1] > 0 and max(idx) >= m.shape[0]:
                raise IndexError("masking(): row %d out of bounds for axis 0 with size %d" % (max(idx), m.shape[0]))
            m[[x for x in idx if x < m.shape[0]], :] = np.nan
    return m


def _sobel_sums(k, *cols):
    """stats.elementwise_product_sum."""
    n = cols[0].size
    res = []
    for i in range(1, n - 1):
        t = 0
        for j in range(len(cols)):
            t += k[0][j] * cols[j][i - 1] + k[1][j] * cols[j][i] + k[2][j] * cols[j][i + 1]
        if len(cols) == 3 and t < 0:
            t *= -1
        res.append(t)
    return res


K_XL = np.array([[-1, 1], [-2, 2], [-1, 1]])
K_XR = np.array([[1, -1], [2, -2], [1, -1]])
K_Y = np.array([[1, 2, 1], [0, 0, 0], [-1, -2, -1]])


def stripiness(mat, table, expected, mask="0", on_index_error="raise"):
    """scoringstripes() -> (Stripiness, O/E mean, O/E total) lists.  on_index_error="nan": rows on
    which masking() raises IndexError (the reference run dies there) get NaN instead — the
    convention of the CUDA path."""
    r = mat.binsize
    bs = int(50000 / r)
    n = len(table["pos1"])
    g_out, m_out, t_out = [0] * n, [0] * n, [0] * n
    if mask != "0":
        mc, rest = mask.split(":")
        ms, me = int(rest.split("-")[0]), int(rest.split("-")[1])
        mxs, mxe = int(ms / r), int(me / r)
    for i in range(n):
        chrom = str(table["chr"][i])
        is_mask = mask != "0" and mc == chrom
        nb = mat.nbins(chrom)
        ex = expected[chrom]
        xs, xe, ys, ye = (table[k][i] for k in ("pos1", "pos2", "pos3", "pos4"))
        xsi, xei = int(xs / r), int(xe / r)
        ysi, yei = int(ys / r), int(ye / r)
        lm = max(xsi - bs, 1)
        rm = xei + bs
        if rm >= nb:
            rm = nb - 1
        with np.errstate(divide="ignore", invalid="ignore"):
            c_obs = mat.fetch_bp(chrom, int(ys), int(ye), int(xs), int(xe))
            center = c_obs / (_expected(ex, xsi, xei, ysi, yei) + .00000001)
            l_obs = mat.fetch_bp(chrom, int(ys), int(ye), int(lm * r), int(xsi * r))
            left = l_obs / (_expected(ex, lm, xsi, ysi, yei) + .00000001)
            r_obs = mat.fetch_bp(chrom, int(ys), int(ye), int(xei * r), int(rm * r))
            right = r_obs / (_expected(ex, xei + 1, rm + 1, ysi, yei) + .00000001)
        if is_mask and ms > np.min([xs - 50000, ys]) and ms < np.max([xe + 50000, ye]):
            try:
                center = _mask_cols(center, mxs, mxe, xsi, xei, 1)
                center = _mask_cols(center, mxs, mxe, ysi, yei, 2)
                left = _mask_cols(left, mxs, mxe, lm, xsi, 1)
                left = _mask_cols(left, mxs, mxe, ysi, yei, 2)
                right = _mask_cols(right, mxs, mxe, xei + 1, rm, 1)
                right = _mask_cols(right, mxs, mxe, ysi, yei, 2)
            except IndexError:
                if on_index_error != "nan":
                    raise
                g_out[i], m_out[i], t_out[i] = float("nan"), float("nan"), float("nan")
                continue
        parts = [center.astype(float), left.astype(float), right.astype(float)]
        rowdel = []
        for k, p in enumerate(parts):
            dead = [x for x in range(p.shape[1]) if np.isnan(p[:, x]).all()]
            if dead:
                p = np.delete(p, dead, axis=1)
                parts[k] = p
                rowdel += dead if xs == ys else [p.shape[0] - 1 - x for x in dead]
        rowdel = np.unique(rowdel)
        if len(rowdel) > 0:
            parts = [np.delete(p, rowdel.astype(int), axis=0) for p in parts]
        center, left, right = (nan0(p) for p in parts)
        with np.errstate(divide="ignore", invalid="ignore"):
            cm_, lm_, rm_ = center.mean(axis=1), left.mean(axis=1), right.mean(axis=1)
            rows_rep = np.nonzero(~np.isnan(center))[0]
            total = np.sum(center[rows_rep])
            mean = np.mean(center[rows_rep])
            gxl = _sobel_sums(K_XL, lm_, cm_)
            gxr = _sobel_sums(K_XR, cm_, rm_)
            gy = _sobel_sums(K_Y, lm_, cm_, rm_)
            gx = np.minimum(gxl, gxr)
            diff = [a - b for a, b in zip(gx, gy)]
            diff = [v for v in diff if v >= 0 or v < 0]
            g = float(np.nanmedian(cm_) * np.mean(diff))
        g_out[i], m_out[i], t_out[i] = g, mean, total
    return g_out, m_out, t_out


def observed_mean(mat, table):
    """getMean(): nanmean / nansum of the stripe body."""
    ms, ss = [], []
    for i in range(len(table["pos1"])):
        chrom = str(table["chr"][i])
        w = mat.fetch_bp(chrom, int(table["pos3"][i]), int(table["pos4"][i]),
                         int(table["pos1"][i]), int(table["pos2"][i]))
        with np.errstate(divide="ignore", invalid="ignore"):
            ms.append(np.nanmean(w))
            ss.append(np.nansum(w))
    return ms, ss
