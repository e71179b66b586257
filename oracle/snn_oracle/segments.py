"""ORACLE (test infrastructure, never on the product path) — vertical-segment and
candidate-stripe search of ``getStripe.StripeSearch`` restated over whole-column numpy vectors.

Follows reference src/stripenn/ImageProcessing.py:61-83 (verticalLine), :100-195 (block) and
src/stripenn/getStripe.py:923-1085 (candidate test, refinement, grouping state machine,
pairing).  All integer work; the only floating point is the count-weighted mean column of
the grouping step, evaluated in the reference's sequential order.
"""
import numpy as np


def vertical_line(edges):
    """verticalLine(M, 60, 120): true convolution with Gx, Gy (zero padding), orientation
    atan2(Fx, Fy) in (60°,120°)  <=>  Fx > 0 and 3*Fy^2 < Fx^2 for the integer responses;
    a hit at (x, y) sets (x, y-1) (ImageProcessing.py:78-81)."""
    E = np.asarray(edges).astype(np.int64)
    S0, S1 = E.shape
    P = np.pad(E, 1)
    col3 = P[0:S0, :] + 2 * P[1:S0 + 1, :] + P[2:S0 + 2, :]        # (S0, S1+2) vertical 1-2-1
    fx = col3[:, 0:S1] - col3[:, 2:S1 + 2]                           # left minus right
    row3 = P[:, 0:S1] + 2 * P[:, 1:S1 + 1] + P[:, 2:S1 + 2]         # (S0+2, S1) horizontal 1-2-1
    fy = row3[2:S0 + 2, :] - row3[0:S0, :]                           # below minus above
    hit = (fx > 0) & (3 * fy * fy < fx * fx)
    out = np.zeros((S0, S1), dtype=np.uint8)
    out[:, :-1] = hit[:, 1:]          # the y=0 -> -1 wrap is unreachable (fx[:,0] <= 0)
    return out


def block_scan(vert):
    """block(vert, c) for every column c at once: longest run of ones in max(vert[:,c-1:c+2])
    under the 5-zero gap buffer that ones do not reset (ImageProcessing.py:119-138,192-194)."""
    V = np.asarray(vert).astype(bool)
    S = V.shape[1]
    nrow = V.shape[1]  # the reference loops range(S) with S = shape[1] over rows (square)
    P = np.pad(V, ((0, 0), (1, 1)))
    val = P[:, 0:S] | P[:, 1:S + 1] | P[:, 2:S + 2]
    count = np.zeros(S, np.int64)
    buf = np.zeros(S, np.int64)
    best = np.zeros(S, np.int64)
    end = np.zeros(S, np.int64)
    last = np.zeros(S, np.int64)
    for i in range(nrow):
        one = val[i]
        count += one
        last = np.where(one, i, last)
        zero = ~one
        room = buf < 5
        close = zero & ~room
        buf += zero & room
        upd = close & (count > best)
        best = np.where(upd, count, best)
        end = np.where(upd, last, end)
        count[close] = 0
        buf[close] = 0
    upd = count > best
    best = np.where(upd, count, best)
    end = np.where(upd, last, end)
    cols = np.arange(S)
    end = np.where(end < cols, end - best + 1, end)
    return best, end


def candidate_columns(vert, minH):
    """getStripe.py:924-940 -> (columns, END, updown) with updown 1 = up, 2 = down."""
    V = np.asarray(vert).astype(np.int64)
    S = V.shape[0]
    best, end = block_scan(V)
    cols = np.arange(S)
    above = np.minimum(cols, end)
    below = np.maximum(cols, end)
    cs = np.vstack([np.zeros((1, S), np.int64), np.cumsum(V, axis=0)])
    has = (cs[below + 1, cols] - cs[above, cols]) != 0
    ok = (best > minH) & has
    c = cols[ok]
    e = end[ok]
    ud = np.where(e > c, 2, 1)
    return c, e, ud


def _runs(row):
    """Start/end (inclusive) of horizontal runs of ones (getStripe.py:958-966)."""
    v = np.concatenate([[0], row.astype(np.int8), [0]])
    d = np.diff(v)
    return np.nonzero(d == 1)[0], np.nonzero(d == -1)[0] - 1


def build_testmat(S, cols, ends, ud_sel, edges, vert):
    """getStripe.py:946-978 for one direction: paint candidate segments (half-open row
    ranges), then per row replace each horizontal run by vert (if the run touches an edge
    pixel) or by its half-even-rounded midpoint."""
    T = np.zeros((S, S), dtype=np.uint8)
    for c, e in zip(cols, ends):
        st, en = (e, c) if ud_sel == 1 else (c, e)
        if en > st:
            T[max(st, 0):en, c] = 1   # st >= 0 always
    E = np.asarray(edges).astype(bool)
    V = np.asarray(vert).astype(np.uint8)
    for r in np.nonzero(T.any(axis=1))[0]:
        row = T[r].copy()
        st, en = _runs(row)
        for s, e in zip(st, en):
            if E[r, s:e + 1].any():
                T[r, s:e] = V[r, s:e]
            else:
                med = int(np.round((s + e) / 2))
                T[r, s:e] = 0
                T[r, med] = 1
    return T


def group_columns(cols, lens):
    """Grouping state machine of getStripe.py:996-1030 -> sorted unique rounded means.
    ``cols`` ascending column ids with >= 3 ones, ``lens`` their counts."""
    n = len(cols)
    mean_x = []
    cont, ln, open_ = [], [], False

    def flush(cc, ll):
        tot = sum(ll)
        fr = [a / tot for a in ll]
        return np.round(sum([a * b for a, b in zip(cc, fr)]))

    for k in range(n - 1):
        cur, nxt = cols[k], cols[k + 1]
        adjacent = (nxt - cur) == 1
        if adjacent and open_:
            cont.append(nxt); ln.append(lens[k + 1])
        elif adjacent:
            cont = [cur, nxt]; ln = [lens[k], lens[k + 1]]; open_ = True
        elif not open_:
            cont = [cur]; ln = [np.float64(1.0)]   # already normalised: len/len
            mean_x.append(flush(cont, ln))
        else:
            mean_x.append(flush(cont, ln))
            cont = [cur]; ln = [lens[k]]; open_ = False
    if len(ln) == 0:
        mean_x.append(np.float64(0.0))
    else:
        mean_x.append(flush(cont, ln))
    return sorted(set(float(v) for v in mean_x))


def pair_columns(T, X, ud_sel, maxW):
    """getStripe.py:1034-1078 -> list of (x, y, w, h)."""
    S = T.shape[0]
    out = []
    for k in range(len(X) - 1):
        n, m = int(X[k]), int(X[k + 1])
        gap = abs(m - n)
        if not (gap > 1 and gap <= maxW):
            continue
        right = m - 2 if gap > 4 else m
        r1 = np.nonzero(T[:, max(0, n - 1):min(n + 2, S)] == 1)[0]
        r2 = np.nonzero(T[:, max(0, m - 1):min(m + 2, S)] == 1)[0]
        lo = min(r1.min(), r2.min())
        hi = max(r1.max(), r2.max())
        if ud_sel == 1:
            hi = right
        else:
            lo = n
        out.append((n, int(lo), int(right - n + 1), int(hi - lo + 1)))
    return out


def search_pass(edges, vert, minH, maxW):
    """Everything after verticalLine for one brightness pass -> [(x, y, w, h), ...] in the
    reference's emission order (up candidates first, then down)."""
    S = vert.shape[0]
    c, e, ud = candidate_columns(vert, minH)
    res = []
    for d in (1, 2):
        sel = ud == d
        T = build_testmat(S, c[sel], e[sel], d, edges, vert)
        cnt = (T == 1).sum(axis=0)
        keep = np.nonzero(cnt >= 3)[0]
        X = group_columns([int(v) for v in keep], [np.int64(cnt[v]) for v in keep])
        res.extend(pair_columns(T, X, d, maxW))
    return res
