"""Restatement of ``skimage.feature.canny`` (scikit-image >= 0.19: skimage/feature/_canny.py
and _canny_cy.pyx) for the call ``feature.canny(gray_float32, sigma=s)`` made at
reference src/stripenn/getStripe.py:917.

scikit-image is a `*`-pinned dependency of the reference (pyproject.toml:38,52) whose source
is not under /root/reference and which is not installed here, so its published algorithm is
restated with the SAME building blocks it uses (scipy.ndimage gaussian_filter / sobel /
label, which ARE installed) plus a vectorised copy of the Cython bilinear NMS:

* float32 input stays float32; thresholds default to low=0.1, high=0.2 (dtype_max = 1);
* mode='constant', cval=0  =>  bleed-over normalisation:
  ``smoothed = gaussian(image) / (gaussian(ones) + finfo(float32).eps)``;
* eroded mask = everything except the 1-pixel border;
* jsobel = ndi.sobel(axis=1), isobel = ndi.sobel(axis=0) (mode reflect);
  magnitude = sqrt(isobel*isobel + jsobel*jsobel) in float32;
* NMS: float32 ``w`` and ``n2*w``; ``(1.0 - w)`` and the sum promoted to double;
* hysteresis: 8-connected components of (low_masked > 0) containing a pixel >= high.

PARITY UNPINNED against scikit-image itself (absent); this file DEFINES the behaviour the
oracle and the CUDA path are tested against.
TEST INFRASTRUCTURE ONLY (oracle/refharness) — never imported by the product.
"""
import numpy as np
import scipy.ndimage as ndi


def _nms_bilinear(isobel, jsobel, magnitude, eroded_mask, low_threshold):
    f32 = np.float32
    rows, cols = magnitude.shape
    out = np.zeros((rows, cols), dtype=magnitude.dtype)
    low = f32(low_threshold)
    if low == 0:
        low = f32(1e-14)
    m = magnitude
    cand = eroded_mask & (m >= low)
    is_down = isobel <= 0
    is_up = isobel >= 0
    is_left = jsobel <= 0
    is_right = jsobel >= 0
    cond1 = (is_up & is_right) | (is_down & is_left)
    cond2 = (is_down & is_right) | (is_up & is_left)
    ai = np.abs(isobel)
    aj = np.abs(jsobel)

    pad = np.zeros((rows + 2, cols + 2), dtype=m.dtype)
    pad[1:-1, 1:-1] = m

    def sh(dx, dy):  # magnitude[x+dx, y+dy]
        return pad[1 + dx:1 + dx + rows, 1 + dy:1 + dy + cols]

    with np.errstate(divide="ignore", invalid="ignore"):
        w_ji = (aj / ai).astype(f32)
        w_ij = (ai / aj).astype(f32)

    c1a = cond1 & (ai > aj)
    c1b = cond1 & ~(ai > aj)
    c2a = ~cond1 & cond2 & (ai < aj)
    c2b = ~cond1 & cond2 & ~(ai < aj)

    w = np.zeros_like(m)
    n11 = np.zeros_like(m); n12 = np.zeros_like(m)
    n21 = np.zeros_like(m); n22 = np.zeros_like(m)
    for sel, ww, a, b, c, d in (
        (c1a, w_ji, sh(1, 0), sh(1, 1), sh(-1, 0), sh(-1, -1)),
        (c1b, w_ij, sh(0, 1), sh(1, 1), sh(0, -1), sh(-1, -1)),
        (c2a, w_ij, sh(0, 1), sh(-1, 1), sh(0, -1), sh(1, -1)),
        (c2b, w_ji, sh(-1, 0), sh(-1, 1), sh(1, 0), sh(1, -1)),
    ):
        w[sel] = ww[sel]
        n11[sel] = a[sel]; n12[sel] = b[sel]; n21[sel] = c[sel]; n22[sel] = d[sel]

    one_minus_w = 1.0 - w.astype(np.float64)
    m64 = m.astype(np.float64)
    with np.errstate(invalid="ignore"):
        c_plus = ((n12 * w).astype(f32).astype(np.float64) + n11.astype(np.float64) * one_minus_w) <= m64
        c_minus = ((n22 * w).astype(f32).astype(np.float64) + n21.astype(np.float64) * one_minus_w) <= m64
    keep = cand & (cond1 | cond2) & c_plus & c_minus
    out[keep] = m[keep]
    return out


def canny(image, sigma=1.0, low_threshold=None, high_threshold=None, mask=None,
          use_quantiles=False, *, mode="constant", cval=0.0):
    if image.ndim != 2:
        raise ValueError("canny expects a 2-D image")
    if mask is not None or use_quantiles or mode != "constant" or cval != 0.0:
        raise NotImplementedError("only the call form used by stripenn is restated")
    if image.dtype != np.float32:
        image = image.astype(np.float64)
    ftype = image.dtype.type
    low_threshold = 0.1 if low_threshold is None else low_threshold
    high_threshold = 0.2 if high_threshold is None else high_threshold

    ones = np.ones(image.shape, dtype=image.dtype)
    bleed = ndi.gaussian_filter(ones, sigma, mode="constant", cval=0.0, truncate=4.0) \
        + np.finfo(ftype).eps
    smoothed = ndi.gaussian_filter(image, sigma, mode="constant", cval=0.0, truncate=4.0)
    smoothed /= bleed
    eroded = np.ones(image.shape, dtype=bool)
    eroded[:1, :] = False
    eroded[-1:, :] = False
    eroded[:, :1] = False
    eroded[:, -1:] = False

    jsobel = ndi.sobel(smoothed, axis=1)
    isobel = ndi.sobel(smoothed, axis=0)
    magnitude = isobel * isobel
    magnitude += jsobel * jsobel
    np.sqrt(magnitude, out=magnitude)

    low_masked = _nms_bilinear(isobel, jsobel, magnitude, eroded, low_threshold)
    low_mask = low_masked > 0
    labels, count = ndi.label(low_mask, np.ones((3, 3), bool))
    if count == 0:
        return low_mask
    high_mask = low_mask & (low_masked >= ftype(high_threshold))
    good = np.zeros((count + 1,), bool)
    good[np.unique(labels[high_mask])] = True
    return good[labels]
