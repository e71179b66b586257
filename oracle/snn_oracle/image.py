"""ORACLE (test infrastructure, never on the product path) — image construction chain of
``getStripe.StripeSearch`` restated with explicit IEEE arithmetic order in plain numpy.

Follows reference src/stripenn/getStripe.py:889-913 and ImageProcessing.py:8-41; the two
OpenCV calls (cv.filter2D 64F, cv.cvtColor float32 RGB2GRAY) are third-party code whose
arithmetic is restated here and pinned against the installed cv2 in
tests/test_oracle_thirdparty.py.
"""
import numpy as np

F32 = np.float32
F64 = np.float64

# np.arange(0.5, 1.01, 0.1)  (getStripe.py:898) — NOT 0.5,0.6,...,1.0 exactly
BRIGHTNESS = np.arange(0.5, 1.01, 0.1)

GRAY_IPP, GRAY_FMA2, GRAY_PLAIN = 0, 1, 2
CR, CG, CB = F32(0.299), F32(0.587), F32(0.114)


def saturate(D, M):
    """getStripe.py:889-895: blue = 255*(M-D)/M ; <0 -> 0 ; /255 ; clip [0,1].
    Returns the green(=blue) channel as float64; the red channel is the constant 1.0."""
    D = np.asarray(D)
    with np.errstate(divide="ignore", invalid="ignore"):
        blue = (255 * (M - D)) / M
        blue = np.where(blue < 0, 0.0, blue)
        g = blue / 255
    return np.clip(g, 0.0, 1.0)


def brightness(g, b):
    """ImageProcessing.py:15-31 for one channel with In=[0,b], Out=[0,1]:
    v<=0 -> 0 ; v>b -> 1 ; else k*v with k = (1-0)/(b-0)."""
    k = (1.0 - 0.0) / (b - 0.0)
    out = np.zeros_like(g, dtype=F64)
    mid = (g > 0.0) & (g <= b)
    out[g > b] = 1.0
    out[mid] = k * (g[mid] - 0.0) + 0.0
    return out


def mean_blur(ch, bf):
    """cv.filter2D(img, -1, ones((bf,bf))/(bf*bf)) on float64 (getStripe.py:907-910):
    BORDER_REFLECT_101, accumulator starts at 0, taps visited row-major, one multiply and one
    add per tap (no FMA)."""
    bf = int(bf)
    kv = (np.ones((bf, bf)) / (bf * bf))[0, 0]
    a = bf // 2
    S0, S1 = ch.shape
    P = np.pad(ch.astype(F64), ((a, bf - 1 - a), (a, bf - 1 - a)), mode="reflect")
    acc = np.zeros((S0, S1), dtype=F64)
    for ky in range(bf):
        for kx in range(bf):
            acc = acc + kv * P[ky:ky + S0, kx:kx + S1]
    return acc


def blurred_red(bf):
    """What the constant red channel (1.0 everywhere) becomes after blur+clip+float32."""
    bf = int(bf)
    kv = (np.ones((bf, bf)) / (bf * bf))[0, 0]
    s = 0.0
    for _ in range(bf * bf):
        s = s + kv * 1.0
    return F32(min(max(s, 0.0), 1.0))


def _fma_f32(a, b, c):
    """Correctly rounded float32 fma(a,b,c) for float32 arrays, via exact float64 product,
    TwoSum and round-to-odd before the final narrowing (no double rounding)."""
    p = a.astype(F64) * b.astype(F64)  # exact: 24+24 bits
    c = c.astype(F64)
    s = p + c
    bb = s - p
    err = (p - (s - bb)) + (c - bb)  # exact error of the sum
    bits = s.view(np.int64).copy()
    inexact = err != 0
    even = (bits & 1) == 0
    # neighbour of s in the direction of err
    toward_up = (err > 0) == (s > 0)  # magnitude grows if err has the sign of s
    adj = np.where(toward_up, 1, -1)
    bits = np.where(inexact & even, bits + adj, bits)
    return bits.view(F64).astype(F32)


def to_gray(r32, g32, variant=GRAY_IPP):
    """cv.cvtColor(float32 RGB, COLOR_RGB2GRAY) (getStripe.py:913) with G == B.
    ``r32`` scalar float32 (blurred red), ``g32`` float32 array."""
    g32 = g32.astype(F32)
    r = np.full(g32.shape, r32, dtype=F32)
    if variant == GRAY_IPP:      # fmaf(B,cb, fl(fl(R*cr) + fl(G*cg)))
        t = (r * CR) + (g32 * CG)
        return _fma_f32(g32, np.full(g32.shape, CB, F32), t.astype(F32))
    if variant == GRAY_FMA2:     # fmaf(B,cb, fmaf(G,cg, fl(R*cr)))
        t = _fma_f32(g32, np.full(g32.shape, CG, F32), (r * CR).astype(F32))
        return _fma_f32(g32, np.full(g32.shape, CB, F32), t)
    if variant == GRAY_PLAIN:    # fl(fl(fl(R*cr)+fl(G*cg)) + fl(B*cb))
        return ((r * CR) + (g32 * CG)) + (g32 * CB)
    raise ValueError("unknown gray variant")


def gray_image(G, b, bf, variant=GRAY_IPP):
    """saturated channel -> brightness -> mean blur -> clip -> float32 -> gray
    (one pass of the loop at getStripe.py:898-913)."""
    adj = brightness(G, b)
    blur = np.clip(mean_blur(adj, bf), 0.0, 1.0)
    return to_gray(blurred_red(bf), blur.astype(F32), variant)
