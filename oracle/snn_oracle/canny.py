"""ORACLE (test infrastructure, never on the product path) — ``skimage.feature.canny`` as
called at reference src/stripenn/getStripe.py:917, restated in plain numpy with the exact
arithmetic of the scipy.ndimage C loops it is built on.

Third-party algorithm restated (scikit-image is `*`-pinned in the reference's
pyproject.toml:38,52 and absent here — PARITY UNPINNED against scikit-image itself):
 skimage/feature/_canny.py  (>= 0.19): _preprocess, canny
 skimage/feature/_canny_cy.pyx: _nonmaximum_suppression_bilinear
 scipy/ndimage/_filters.py: _gaussian_kernel1d, gaussian_filter1d, sobel, correlate1d
 scipy/ndimage/src/ni_filters.c: NI_Correlate1D (symmetric / antisymmetric fast paths:
   double line buffers, ``o = x[0]*w[0]; for jj=-r..-1: o += (x[jj] +/- x[-jj]) * w[jj]``,
   result cast to the float32 output after EACH axis).
Pinned against the installed scipy in tests/test_oracle_thirdparty.py.
"""
import numpy as np

F32 = np.float32
F64 = np.float64
LOW = F32(0.1)
HIGH = F32(0.2)


def gaussian_weights(sigma, truncate=4.0):
    """scipy.ndimage._filters._gaussian_kernel1d(sigma, 0, int(truncate*sigma+0.5))."""
    sd = float(sigma)
    radius = int(truncate * sd + 0.5)
    sigma2 = sd * sd
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / sigma2 * x ** 2)
    phi = phi / phi.sum()
    return phi, radius


def _corr_sym(x32, w, radius, axis, pad_mode):
    """One NI_Correlate1D pass with a symmetric kernel.  float32 in -> float32 out,
    float64 accumulation, farthest pair first."""
    x = np.moveaxis(x32, axis, 0).astype(F64)
    n = x.shape[0]
    if pad_mode == "constant":
        P = np.pad(x, ((radius, radius), (0, 0)), mode="constant")
    else:  # scipy 'reflect' == numpy 'symmetric'
        P = np.pad(x, ((radius, radius), (0, 0)), mode="symmetric")
    c = radius
    t = P[c:c + n] * w[c]
    for k in range(radius, 0, -1):
        t = t + (P[c - k:c - k + n] + P[c + k:c + k + n]) * w[c - k]
    return np.moveaxis(t.astype(F32), 0, axis)


def _corr_antisym3(x32, axis):
    """correlate1d(x, [-1, 0, 1]) with mode reflect: o = x[0]*0 ; o += (x[-1]-x[+1]) * -1."""
    x = np.moveaxis(x32, axis, 0).astype(F64)
    n = x.shape[0]
    P = np.pad(x, ((1, 1), (0, 0)), mode="symmetric")
    t = P[1:1 + n] * 0.0
    t = t + (P[0:n] - P[2:2 + n]) * (-1.0)
    return np.moveaxis(t.astype(F32), 0, axis)


def gaussian2d(x32, w, radius):
    """ndi.gaussian_filter(float32, sigma, mode='constant'): axis 0 then axis 1."""
    t = _corr_sym(x32, w, radius, 0, "constant")
    return _corr_sym(t, w, radius, 1, "constant")


def bleed_over(shape, w, radius):
    ones = np.ones(shape, dtype=F32)
    return gaussian2d(ones, w, radius) + np.finfo(F32).eps


def sobel(x32, axis):
    """ndi.sobel(x, axis): derivative along `axis`, then [1,2,1] along the other axis."""
    w121 = np.array([1.0, 2.0, 1.0])
    d = _corr_antisym3(x32, axis)
    return _corr_sym(d, w121, 1, 1 - axis, "symmetric")


def gradient(gray32, w, radius, bleed=None):
    if bleed is None:
        bleed = bleed_over(gray32.shape, w, radius)
    smoothed = gaussian2d(gray32, w, radius) / bleed
    jsobel = sobel(smoothed, 1)
    isobel = sobel(smoothed, 0)
    mag = isobel * isobel
    mag = mag + jsobel * jsobel
    mag = np.sqrt(mag)
    return smoothed, isobel, jsobel, mag


def nms(isobel, jsobel, mag):
    """_nonmaximum_suppression_bilinear: returns the boolean 'kept' mask (m >= low, interior,
    local maximum along the gradient).  float32 build of the fused Cython function: ``w`` and
    ``n2*w`` are float32, ``(1.0 - w)`` and the sum are double."""
    S0, S1 = mag.shape
    keep = np.zeros((S0, S1), dtype=bool)
    if S0 < 3 or S1 < 3:
        return keep
    sl = (slice(1, S0 - 1), slice(1, S1 - 1))
    i = isobel[sl]
    j = jsobel[sl]
    m = mag[sl]

    def nb(dx, dy):
        return mag[1 + dx:S0 - 1 + dx, 1 + dy:S1 - 1 + dy]

    up, down = i >= 0, i <= 0
    right, left = j >= 0, j <= 0
    cond1 = (up & right) | (down & left)
    cond2 = (down & right) | (up & left)
    ai, aj = np.abs(i), np.abs(j)
    with np.errstate(divide="ignore", invalid="ignore"):
        wj_i = aj / ai  # float32 / float32
        wi_j = ai / aj
    gt = ai > aj
    lt = ai < aj
    cases = (
        (cond1 & gt, wj_i, nb(1, 0), nb(1, 1), nb(-1, 0), nb(-1, -1)),
        (cond1 & ~gt, wi_j, nb(0, 1), nb(1, 1), nb(0, -1), nb(-1, -1)),
        (~cond1 & cond2 & lt, wi_j, nb(0, 1), nb(-1, 1), nb(0, -1), nb(1, -1)),
        (~cond1 & cond2 & ~lt, wj_i, nb(-1, 0), nb(-1, 1), nb(1, 0), nb(1, -1)),
    )
    ok = np.zeros(m.shape, dtype=bool)
    m64 = m.astype(F64)
    for sel, w, n11, n12, n21, n22 in cases:
        w = w.astype(F32)
        omw = 1.0 - w.astype(F64)
        with np.errstate(invalid="ignore"):
            plus = (n12 * w).astype(F32).astype(F64) + n11.astype(F64) * omw
            minus = (n22 * w).astype(F32).astype(F64) + n21.astype(F64) * omw
            ok |= sel & (plus <= m64) & (minus <= m64)
    keep[sl] = ok & (m >= LOW)
    return keep


def hysteresis(low_mask, high_mask):
    """Keep the 8-connected components of low_mask that contain a high_mask pixel
    (ndi.label + good_label lookup in skimage == geodesic reconstruction)."""
    try:
        import scipy.ndimage as ndi
        labels, count = ndi.label(low_mask, np.ones((3, 3), bool))
        if count == 0:
            return low_mask.copy()
        good = np.zeros(count + 1, dtype=bool)
        good[np.unique(labels[high_mask])] = True
        good[0] = False
        return good[labels]
    except ImportError:  # pure numpy fallback: iterate dilation to the fixed point
        cur = high_mask & low_mask
        while True:
            p = np.pad(cur, 1)
            grow = np.zeros_like(cur)
            for dx in (0, 1, 2):
                for dy in (0, 1, 2):
                    grow |= p[dx:dx + cur.shape[0], dy:dy + cur.shape[1]]
            nxt = grow & low_mask
            if (nxt == cur).all():
                return cur
            cur = nxt


def canny(gray32, sigma, weights=None, bleed=None, return_stages=False):
    if weights is None:
        weights = gaussian_weights(sigma)
    w, radius = weights
    smoothed, isobel, jsobel, mag = gradient(gray32.astype(F32), w, radius, bleed)
    low_mask = nms(isobel, jsobel, mag)
    high_mask = low_mask & (mag >= HIGH)
    edges = hysteresis(low_mask, high_mask)
    if return_stages:
        return dict(smoothed=smoothed, isobel=isobel, jsobel=jsobel, mag=mag,
                    low=low_mask, high=high_mask, edges=edges)
    return edges
