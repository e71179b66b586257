"""GPU (-m gpu): parity at BASELINE.json config sizes (VERDICT r01 "parity at config size").

* `mid`: the UNMODIFIED reference's end tables and stage products for a 12,050-bin KR-balanced
  chromosome, five maxpixel quantiles, plus a chromosome shorter than 500 bins
  (tests/golden/mid_compute.npz, made by tests/golden/make_golden_mid.py) — end to end.
* raw-candidate EQUALITY with the oracle (every brightness pass, before any dedup) on sampled
  windows of config 1 (chr16-sized, 5 kb, raw counts), a config-3 chromosome (hg38 chr21 at 5 kb,
  KR weights with NaN bins) and config 4 (hg38 chr1 at 1 kb, bs = 50).
* config-5 shape: 5,000 of the bench's synthetic stripes on config-3 chromosomes, all six scoring
  outputs against snn_oracle.scoring.
Bar: coordinates / maxpixel bit-exact; Mean, pvalue, Stripiness within 1e-5 relative."""
import os

import numpy as np
import pandas as pd
import pytest

from snn_oracle import canny as cn
from snn_oracle import driver
from snn_oracle import image as im
from snn_oracle import pipeline as pl
from snn_oracle import segments as sg
from snn_oracle.matrix import OracleMatrix
from stripenn_b200 import compute as cp
from stripenn_b200 import getstripe as gs
from stripenn_b200 import synth
from stripenn_b200.score import score_table

pytestmark = pytest.mark.gpu
RTOL = 1e-5
QS = [0.95, 0.96, 0.97, 0.98, 0.99]
BG = ("blu", "bru", "bld", "brd")


def _unload_all(dev):
    for cid in list(dev.loaded):
        dev.unload_chrom(cid)


@pytest.fixture()
def dev(device):
    _unload_all(device)
    yield device
    _unload_all(device)


def test_mid_compute_matches_reference(dev, golden_dir):
    g = np.load(os.path.join(golden_dir, "mid_compute.npz"))
    cm = synth.make_config("mid")
    unf, fil, st = cp.run_stages(cm, norm="KR", maxpixel="0.95,0.96,0.97,0.98,0.99", device=dev,
                                 return_stages=True)
    for df, pre in ((unf, "unf_"), (fil, "fil_")):
        assert len(df) == len(g[pre + "pos1"])
        for k in ("pos1", "pos2", "pos3", "pos4", "length", "width"):
            assert df[k].astype(np.int64).tolist() == g[pre + k].tolist(), (pre, k)
        assert df["chr"].astype(str).tolist() == g[pre + "chr"].tolist()
        assert df["maxpixel"].tolist() == g[pre + "maxpixel"].tolist()
        for k in ("Mean", "pvalue", "Stripiness"):
            np.testing.assert_allclose(df[k].to_numpy(dtype=float), g[pre + k], rtol=RTOL, atol=0, equal_nan=True)
    for c in ("chrL", "chrS"):
        np.testing.assert_allclose(st["MP"][c], g["MP_" + c], rtol=1e-15)
        np.testing.assert_allclose(st["EV"][c], g["EV_" + c], rtol=1e-12)
    for name, t in zip(BG, st["bg"]):
        assert list(t.shape) == g["bg_%s_shape" % name].tolist()     # chrS contributes no column
        np.testing.assert_allclose(t[:, ::25], g["bg_" + name], rtol=1e-12, atol=1e-12, equal_nan=True)
        assert np.array_equal(np.isnan(t).sum(axis=1), g["bg_%s_nnan" % name])
    for i, t in enumerate(st["per_q"]):
        for k in ("pos1", "pos2", "pos3", "pos4", "num", "x", "y", "h", "w"):
            assert t[k].astype(np.int64).tolist() == g["q%d_%s" % (i, k)].tolist(), (i, k)
        np.testing.assert_allclose(t["pvalue"].to_numpy(dtype=float), g["q%d_pvalue" % i], rtol=RTOL)


def _raw_oracle_window(mat, c, idx, M, minH=10, maxW=8, sigma=2.0, bf=3):
    """Every candidate the reference's StripeSearch emits for one window and maxpixel value, in
    emission order (brightness pass, up/down, pair index): snn_oracle restatement."""
    fr = pl.load_frame(mat, c, idx)
    if fr is None:
        return []
    w = cn.gaussian_weights(sigma)
    G = im.saturate(fr["D"], M)
    bleed = cn.bleed_over((fr["S"], fr["S"]), *w)
    rows = []
    for bi, b in enumerate(im.BRIGHTNESS):
        gray = im.gray_image(G, b, bf)
        edges = cn.canny(gray, sigma, w, bleed)
        vert = sg.vertical_line(edges)
        for (x, y, ww, hh) in sg.search_pass(edges, vert, minH, maxW):
            rows.append((bi, x, y, ww, hh, int(fr["sa"][x]), int(fr["ea"][x + ww - 1]), int(fr["sa"][y]),
                         int(fr["ea"][y + hh - 1])))
    return rows


def _check_sampled_windows(dev, cm, chrom, norm, nsample, qsel, seed):
    bal = cp.resolve_norm(cm, norm, "compute", False)
    mat = OracleMatrix(cm, bal)
    lib = gs.DeviceMatrix(cm, bal, dev)
    cid = lib.ensure(chrom)
    mp = dev.quantiles(cid, QS)
    assert np.array_equal(mp, np.quantile(mat.positive_population(chrom), QS))
    nfr = pl.n_frames(mat.nbins(chrom))
    rng = np.random.default_rng(seed)
    frames = sorted(set([0, nfr - 1] + rng.choice(np.arange(1, nfr - 1), size=nsample - 2, replace=False).tolist()))
    ncand = 0
    for idx in frames:
        # one window per call: exercises the window-range entry the multi-GPU shards use
        cand = dev.extract(cid, mp, 2.0, 10, 8, 3, frame_lo=idx, frame_hi=idx + 1)
        for qi in qsel:
            want = _raw_oracle_window(mat, chrom, idx, mp[qi])
            sel = cand[(cand["q"] == qi) & (cand["frame"] == idx)]
            got = [(int(r["bright"]), int(r["x"]), int(r["y"]), int(r["w"]), int(r["h"]), int(r["pos1"]),
                    int(r["pos2"]), int(r["pos3"]), int(r["pos4"])) for r in sel]
            assert got == want, (chrom, idx, qi)
            ncand += len(got)
    assert ncand > 0
    dev.unload_chrom(cid)


def test_raw_candidates_cfg1_sampled(dev):
    _check_sampled_windows(dev, synth.make_config("cfg1"), "16", "None", 8, (0, 2, 4), 1)


def test_raw_candidates_cfg3_chromosome_sampled(dev):
    """hg38 chr21 at 5 kb with KR weights (2 % NaN-weight bins) — a chromosome of config 3."""
    chroms, binsize, _ = synth.config_chroms("cfg3")
    cm = synth.make_map(chroms, binsize, balanced=True, only=["chr21"])
    _check_sampled_windows(dev, cm, "chr21", "KR", 8, (0, 4), 3)


def test_raw_candidates_cfg4_sampled(dev):
    """hg38 chr1 at 1 kb: bs = 50, 1,245 windows (config 4)."""
    _check_sampled_windows(dev, synth.make_config("cfg4"), "chr1", "None", 8, (1, 4), 4)


@pytest.mark.parametrize("fast", ["0", "1"])
def test_cfg5_shape_scores_match_oracle(dev, fast, monkeypatch):
    """`stripenn score` on 5,000 stripes drawn like the bench's config-5 table (h ~ lognormal(55),
    w 3..12, both directions) over three KR-balanced config-3 chromosomes: p-value, Stripiness,
    O/E mean and total, observed mean and sum against the oracle."""
    monkeypatch.setenv("SNN_SCORE_FAST", fast)      # k_score_warp / k_score_fast, both against the oracle
    dev.score_routing()
    chroms, binsize, _ = synth.config_chroms("cfg3")
    keep = ("chr20", "chr21", "chr22")
    cm = synth.make_map([c for c in chroms if c[0] in keep], binsize, balanced=True)
    tab = synth.make_stripe_table(cm, 5000)
    df = pd.DataFrame({"chr": tab["chr"], "pos1": tab["pos1"], "pos2": tab["pos2"], "chr2": tab["chr"],
                       "pos3": tab["pos3"], "pos4": tab["pos4"]})
    got = score_table(cm, df, norm="KR", device=dev)
    want = driver.score(cm, {k: v.tolist() for k, v in tab.items()}, norm="KR")
    for k in ("pvalue_added", "Stripiness_added", "O_Mean_added", "O_Sum_added", "O/E_Mean_added", "O/E_Total_added"):
        np.testing.assert_allclose(got[k].to_numpy(dtype=float), np.array(want[k], dtype=float), rtol=RTOL, atol=0,
                                   equal_nan=True, err_msg=k)
    assert np.isfinite(got["pvalue_added"].to_numpy(dtype=float)).all()
    routed = dev.score_routing()
    # score_table makes one call per output group: every stripe of the table is regular
    assert (routed[0] > 0 and routed[1] + routed[2] == 0) if fast == "1" else routed[0] == 0, routed
