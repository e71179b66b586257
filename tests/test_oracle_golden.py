"""CPU: pin the numpy oracle (oracle/snn_oracle) against (a) the golden fixtures produced by
the UNMODIFIED reference (tests/golden/make_golden.py) and (b) the known-answer vectors
derived by executing reference code verbatim."""
import os

import numpy as np
import pytest

from snn_oracle import background as bgm
from snn_oracle import canny as cn
from snn_oracle import driver
from snn_oracle import image as im
from snn_oracle import pipeline as pl
from snn_oracle import segments as sg
from snn_oracle.matrix import OracleMatrix
from stripenn_b200 import synth

BG = ("blu", "bru", "bld", "brd")


def _cmp_tables(got, g, prefix, rtol=1e-9):
    n = len(g[prefix + "pos1"])
    assert len(got["pos1"]) == n
    for k in ("pos1", "pos2", "pos3", "pos4", "length", "width"):
        assert list(got[k]) == g[prefix + k].tolist(), k
    assert [str(x) for x in got["chr"]] == g[prefix + "chr"].tolist()
    assert list(got["maxpixel"]) == g[prefix + "maxpixel"].tolist()
    for k in ("Mean", "pvalue", "Stripiness"):
        np.testing.assert_allclose(np.array(got[k], float), g[prefix + k], rtol=rtol, equal_nan=True)


def test_tiny_compute_matches_reference(tiny_cm, golden_dir):
    g = np.load(os.path.join(golden_dir, "tiny_compute.npz"))
    unf, fil, st = driver.compute(tiny_cm, norm="None", maxpixel="0.95,0.96,0.97,0.98,0.99",
                                  return_stages=True)
    _cmp_tables(unf, g, "unf_")
    _cmp_tables(fil, g, "fil_")
    assert np.array_equal(st["MP"]["T1"], g["MP_T1"])
    np.testing.assert_allclose(st["EV"]["T1"], g["EV_T1"], rtol=1e-12)
    for name, t in zip(BG, st["bg"]):
        assert list(t.shape) == g["bg_%s_shape" % name].tolist()
        np.testing.assert_allclose(t[:, ::25], g["bg_" + name], rtol=1e-12, equal_nan=True)
        np.testing.assert_allclose(np.nansum(t, axis=1), g["bg_%s_nansum" % name], rtol=1e-9, atol=1e-9)
        assert np.array_equal(np.isnan(t).sum(axis=1), g["bg_%s_nnan" % name])
    for i, t in enumerate(st["per_q"]):
        for k in ("pos1", "pos2", "pos3", "pos4", "num", "x", "y", "h", "w"):
            assert list(t[k]) == g["q%d_%s" % (i, k)].tolist(), (i, k)
        np.testing.assert_allclose(np.array(t["pvalue"], float), g["q%d_pvalue" % i], rtol=1e-12)


def test_tiny2_kr_compute_matches_reference(tiny2_cm, golden_dir):
    g = np.load(os.path.join(golden_dir, "tiny2_compute.npz"))
    unf, fil, st = driver.compute(tiny2_cm, norm="KR", maxpixel="0.95,0.98", return_stages=True)
    _cmp_tables(unf, g, "unf_")
    _cmp_tables(fil, g, "fil_")
    for c in ("chrA", "chrB"):
        np.testing.assert_allclose(st["MP"][c], g["MP_" + c], rtol=1e-15)
        np.testing.assert_allclose(st["EV"][c], g["EV_" + c], rtol=1e-12)


def test_mid_kr_compute_matches_reference(golden_dir):
    """Config-size pin (tests/golden/make_golden_mid.py): a 12,050-bin KR chromosome, five maxpixel
    quantiles, plus a chromosome shorter than 500 bins that nulldist must drop from the sample
    budget (itera = 0, getStripe.py:249-283) — end tables and every stage product of the
    UNMODIFIED reference."""
    g = np.load(os.path.join(golden_dir, "mid_compute.npz"))
    cm = synth.make_config("mid")
    unf, fil, st = driver.compute(cm, norm="KR", maxpixel="0.95,0.96,0.97,0.98,0.99", return_stages=True)
    _cmp_tables(unf, g, "unf_")
    _cmp_tables(fil, g, "fil_")
    for c in ("chrL", "chrS"):
        np.testing.assert_allclose(st["MP"][c], g["MP_" + c], rtol=1e-15)
        np.testing.assert_allclose(st["EV"][c], g["EV_" + c], rtol=1e-12)
    for name, t in zip(BG, st["bg"]):
        assert list(t.shape) == g["bg_%s_shape" % name].tolist()
        # differences of block means: near-zero entries carry the summation-order noise of the means
        np.testing.assert_allclose(t[:, ::25], g["bg_" + name], rtol=1e-12, atol=1e-12, equal_nan=True)
        assert np.array_equal(np.isnan(t).sum(axis=1), g["bg_%s_nnan" % name])
    for i, t in enumerate(st["per_q"]):
        for k in ("pos1", "pos2", "pos3", "pos4", "num", "x", "y", "h", "w"):
            assert list(t[k]) == g["q%d_%s" % (i, k)].tolist(), (i, k)
        np.testing.assert_allclose(np.array(t["pvalue"], float), g["q%d_pvalue" % i], rtol=1e-12)


def test_tiny2_score_matches_reference(tiny2_cm, golden_dir):
    g = np.load(os.path.join(golden_dir, "tiny2_score.npz"))
    table = {"chr": g["chr"].tolist(), "pos1": g["pos1"].tolist(), "pos2": g["pos2"].tolist(),
             "pos3": g["pos3"].tolist(), "pos4": g["pos4"].tolist()}
    res = driver.score(tiny2_cm, table, norm="KR")
    for k, v in res.items():
        np.testing.assert_allclose(np.array(v, float), g[k.replace("/", "_")], rtol=1e-9, equal_nan=True)


def test_stage_intermediates_match_reference(tiny_cm, golden_dir):
    g = np.load(os.path.join(golden_dir, "tiny_stages.npz"))
    mat = OracleMatrix(tiny_cm, False)
    M = float(g["M"])
    w = cn.gaussian_weights(2.0)
    for idx in (0, 5):
        fr = pl.load_frame(mat, "T1", idx)
        S = int(g["f%d_S" % idx])
        assert fr["S"] == S
        G = im.saturate(fr["D"], M)
        found = []
        for b in range(6):
            gray = im.gray_image(G, im.BRIGHTNESS[b], 3)
            if "f%d_gray%d" % (idx, b) in g:
                assert np.array_equal(gray, g["f%d_gray%d" % (idx, b)])
            assert float(gray.astype(np.float64).sum()) == float(g["f%d_graysum%d" % (idx, b)])
            edges = cn.canny(gray, 2.0, w)
            vert = sg.vertical_line(edges)
            assert np.array_equal(np.packbits(edges), g["f%d_edges%d" % (idx, b)])
            assert np.array_equal(np.packbits(vert.astype(bool)), g["f%d_vert%d" % (idx, b)])
        res = pl.stripe_search(fr, M, 0.97, "T1", 10, 8, 2.0, 3)
        for k in ("x", "y", "w", "h"):
            assert list(res[k]) == g["f%d_res_%s" % (idx, k)].tolist()


# ---- known-answer vectors from executing reference code verbatim -----------------
def _col10(rows):
    m = np.zeros((40, 40), dtype=np.uint8)
    m[:len(rows), 10] = rows
    return m


@pytest.mark.parametrize("rows,want", [
    ([1] * 12 + [0] * 5 + [1] * 12, (24, 28)),
    ([0] * 6 + [1] * 12 + [0] * 5 + [1] * 12, (24, 34)),
    ([0] * 3 + [1] * 12 + [0] * 3 + [1] * 12, (12, 14)),
    ([0] * 40, (0, 1)),
])
def test_block_known_answers(rows, want):
    best, end = sg.block_scan(_col10(rows))
    assert (int(best[10]), int(end[10])) == want


@pytest.mark.parametrize("cols,lens,want", [
    ([5, 9], [4, 4], [5]), ([5, 6, 9], [4, 4, 4], [6]), ([5, 9, 10], [4, 4, 6], [5, 10]),
    ([5, 6, 7, 20, 21, 40], [3, 9, 3, 5, 5, 7], [6, 20, 21]), ([], [], [0]), ([7], [5], [0]),
    ([4, 5], [3, 3], [4]), ([5, 6], [3, 3], [6]),
])
def test_grouping_known_answers(cols, lens, want):
    assert sg.group_columns(cols, [np.int64(v) for v in lens]) == [float(v) for v in want]


def test_vertical_line_known_answers():
    for col, expect in ((5, 5), (0, 0), (11, None)):
        e = np.zeros((12, 12), dtype=bool)
        e[2:10, col] = True
        v = sg.vertical_line(e)
        want = np.zeros((12, 12), dtype=np.uint8)
        if expect is not None:
            want[2:10, expect] = 1
        assert np.array_equal(v, want)


def test_brightness_values_and_labels():
    assert im.BRIGHTNESS.tolist() == [0.5, 0.6, 0.7, 0.7999999999999999, 0.8999999999999999, 0.9999999999999999]
    assert [pl.maxpixel_label(q) for q in (0.95, 0.96, 0.97, 0.975, 0.98, 0.985, 0.99)] == \
        ["95.0%", "96.0%", "97.0%", "97.5%", "98.0%", "98.5%", "99.0%"]
