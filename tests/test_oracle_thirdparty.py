"""CPU: pin the oracle's restatements of third-party arithmetic against the libraries that are
installed here (scipy.ndimage, OpenCV): the Gaussian/Sobel passes, the 64F mean filter and the
float32 RGB->GRAY rounding variant this cv2 build uses."""
import numpy as np
import pytest
import scipy.ndimage as ndi

from snn_oracle import canny as cn
from snn_oracle import image as im

cv = pytest.importorskip("cv2")


def _img(seed, s=97):
    rng = np.random.default_rng(seed)
    return rng.random((s, s + 5)).astype(np.float32)


@pytest.mark.parametrize("sigma", [2.0, 2.5])
def test_gaussian_matches_scipy(sigma):
    x = _img(1)
    w, r = cn.gaussian_weights(sigma)
    assert np.array_equal(cn.gaussian2d(x, w, r), ndi.gaussian_filter(x, sigma, mode="constant"))


def test_sobel_matches_scipy():
    x = _img(2)
    for axis in (0, 1):
        assert np.array_equal(cn.sobel(x, axis), ndi.sobel(x, axis=axis))


def test_hysteresis_matches_label():
    rng = np.random.default_rng(3)
    low = rng.random((80, 80)) < 0.35
    high = low & (rng.random((80, 80)) < 0.05)
    labels, n = ndi.label(low, np.ones((3, 3), bool))
    good = np.zeros(n + 1, bool)
    good[np.unique(labels[high])] = True
    good[0] = False
    assert np.array_equal(cn.hysteresis(low, high), good[labels])


@pytest.mark.parametrize("bf", [1, 3, 5])
def test_mean_blur_matches_opencv(bf):
    rng = np.random.default_rng(4)
    ch = rng.random((61, 61))
    img = cv.merge((np.ones_like(ch), ch, ch))
    out = cv.filter2D(img, -1, np.ones((bf, bf)) / (bf * bf))
    assert np.array_equal(out[:, :, 1], im.mean_blur(ch, bf))
    assert np.float32(np.clip(out[0, 0, 0], 0, 1)) == im.blurred_red(bf)


def test_gray_variant_of_this_opencv_build():
    rng = np.random.default_rng(5)
    g = rng.random((128, 128)).astype(np.float32)
    r = np.full_like(g, im.blurred_red(3))
    got = cv.cvtColor(cv.merge((r, g, g)), cv.COLOR_RGB2GRAY)
    matches = [v for v in (im.GRAY_IPP, im.GRAY_FMA2, im.GRAY_PLAIN)
               if np.array_equal(got, im.to_gray(im.blurred_red(3), g, v))]
    assert im.GRAY_IPP in matches, "this cv2 build uses gray variant(s) %r" % matches
