"""GPU (-m gpu): parity of the CUDA path, called through the C ABI, against the oracle on the
same seeded inputs and against the golden fixtures produced by the unmodified reference.
Bar: coordinates / maxpixel / masks / float32 images bit-exact; Mean, pvalue, Stripiness
within 1e-5 relative (the north star's tolerance; observed ~1e-15)."""
import os

import numpy as np
import pandas as pd
import pytest

from snn_oracle import background as bgm
from snn_oracle import canny as cn
from snn_oracle import driver
from snn_oracle import image as im
from snn_oracle import pipeline as pl
from snn_oracle import segments as sg
from snn_oracle.matrix import OracleMatrix
from stripenn_b200 import _lib
from stripenn_b200 import compute as cp
from stripenn_b200 import getstripe as gs
from stripenn_b200 import synth
from stripenn_b200.score import score_table

pytestmark = pytest.mark.gpu
RTOL = 1e-5
BG = ("blu", "bru", "bld", "brd")


def _obj(cm, norm, dev, **kw):
    bal = cp.resolve_norm(cm, norm, "compute", False)
    all_names, all_sizes, names, sizes = cp.compute_chromosomes(cm, ["all"])
    o = cp.make_object(cm, bal, dev, all_names, all_sizes, names, sizes, kw.get("minL", 10), kw.get("maxW", 8),
                       kw.get("canny", 2.0), 1, kw.get("bfilter", 3), 123456789, kw.get("gray_variant", 0))
    return o, OracleMatrix(cm, bal), names, all_names, all_sizes


def _unload_all(dev):
    for cid in list(dev.loaded):
        dev.unload_chrom(cid)


@pytest.fixture()
def dev(device):
    _unload_all(device)
    yield device
    _unload_all(device)


@pytest.mark.parametrize("cfg,norm", [("tiny", "None"), ("tiny2", "KR"), ("tiny2", "weight")])
def test_band_quantiles_expected(dev, cfg, norm):
    cm = synth.make_config(cfg)
    obj, mat, names, _, _ = _obj(cm, norm, dev)
    qs = [0.0, 0.5, 0.95, 0.96, 0.97, 0.98, 0.99, 0.999, 1.0]
    MP = obj.getQuantile_original(cm, names, qs)
    EV = obj.mpmean()
    EVo = bgm.expected_by_distance(mat, names)
    for c in names:
        cid = obj.unbalLib.chrom_id(c)
        a = dev.fetch(cid, 300, 700, 280, 690)
        assert np.array_equal(a, mat.window(c, 300, 700, 280, 690).astype(float), equal_nan=True)
        assert np.array_equal(MP[c], np.quantile(mat.positive_population(c), qs)), c
        np.testing.assert_allclose(EV[c], EVo[c], rtol=1e-12)
    with pytest.raises(ValueError):
        dev.quantiles(obj.unbalLib.chrom_id(names[0]), [1.5])
    with pytest.raises(_lib.StripennB200Error):
        dev.fetch(obj.unbalLib.chrom_id(names[0]), 0, 10, 1000, 1010)   # outside the band


def test_raw_count_quantiles_histogram_and_radix_paths(dev, monkeypatch):
    """--norm None: quantiles come from a one-pass histogram of the integer values (k_int_hist,
    exact bins below 4096); a rank beyond the exact bins, or SNN_NO_INT_HIST=1, takes the radix
    select.  All three must equal numpy on the same population."""
    qs = [0.0, 0.31, 0.5, 0.95, 0.97, 0.99, 0.9999, 1.0]
    cm = synth.make_config("tiny2")
    for step, mult in enumerate((1, 70, 12)):   # max count ~1e2 (all exact) / ~1e4 (top ranks overflow) / ~1e5 (most do)
        for c in cm.chroms.values():
            c.count = (c.count.astype(np.int64) * mult).astype(np.int32)
        obj, mat, names, _, _ = _obj(cm, "None", dev)
        want = {c: np.quantile(mat.positive_population(c), qs) for c in names}
        got = obj.getQuantile_original(cm, names, qs)
        monkeypatch.setenv("SNN_NO_INT_HIST", "1")
        radix = obj.getQuantile_original(cm, names, qs)
        monkeypatch.delenv("SNN_NO_INT_HIST")
        for c in names:
            assert np.array_equal(got[c], want[c]), (step, c)
            assert np.array_equal(radix[c], want[c]), (step, c)
        assert max(int(x.count.max()) for x in cm.chroms.values()) > (0 if step == 0 else 4096)
        _unload_all(dev)
    # a chromosome without positive pixels next to a normal one
    cm = synth.make_config("tiny2")
    cm.chroms["chrB"].count[:] = 0
    obj, mat, names, _, _ = _obj(cm, "None", dev)
    got = obj.getQuantile_original(cm, names, [0.5, 0.99])
    assert np.all(np.isnan(got["chrB"]))
    assert np.array_equal(got["chrA"], np.quantile(mat.positive_population("chrA"), [0.5, 0.99]))


@pytest.mark.parametrize("cfg,norm", [("tiny", "None"), ("tiny2", "KR")])
def test_background_exact(dev, cfg, norm):
    cm = synth.make_config(cfg)
    obj, mat, names, all_names, all_sizes = _obj(cm, norm, dev)
    n2, plans = obj.plan_samples()
    n2o, planso = bgm.plan_samples(mat, all_names, all_sizes, 123456789)
    assert n2 == n2o
    for a, b in zip(plans, planso):
        assert np.array_equal(a, np.array(b, dtype=np.int32).reshape(-1, 6))
    bg = obj.nulldist()
    bgo = bgm.null_distribution(mat, all_names, all_sizes, 123456789)
    for a, b in zip(bg, bgo):
        assert a.shape == b.shape
        assert np.array_equal(np.isnan(a), np.isnan(b))
        np.testing.assert_allclose(a, b, rtol=1e-12, atol=1e-12, equal_nan=True)


def test_background_far_pixels_use_csr(dev):
    """A sample whose column offset is stale (+400 on a window that starts at bin 0) reaches
    offsets ~800 > band width: the kernel must answer from the CSR pixels, not with zeros."""
    cm = synth.make_config("tiny")
    obj, mat, names, _, _ = _obj(cm, "None", dev)
    cid = obj.unbalLib.ensure("T1")
    assert dev.loaded[cid]["bw"] < 600
    spec = [(0, 599, 0, 999, x, 400) for x in (30, 200, 450)]
    got = dev.background(cid, np.array(spec, dtype=np.int32))
    want = bgm.background_columns(mat, "T1", spec)
    for t in range(4):
        np.testing.assert_allclose(got[t], want[t], rtol=1e-12, atol=1e-12, equal_nan=True)
    assert np.nanmax(np.abs(want[2])) > 0


def test_stage_images_match_reference_golden(dev, tiny_cm, golden_dir):
    g = np.load(os.path.join(golden_dir, "tiny_stages.npz"))
    obj, mat, names, _, _ = _obj(tiny_cm, "None", dev)
    cid = obj.unbalLib.ensure("T1")
    M = float(g["M"])
    for idx in (0, 5):
        for b in range(6):
            d = dev.debug_stages(cid, M, 2.0, 3, idx, b)
            assert d["S"] == int(g["f%d_S" % idx])
            assert np.array_equal(d["nz"], g["f%d_nz" % idx])
            if "f%d_gray%d" % (idx, b) in g:
                assert np.array_equal(d["gray"], g["f%d_gray%d" % (idx, b)])
            assert float(d["gray"].astype(np.float64).sum()) == float(g["f%d_graysum%d" % (idx, b)])
            assert np.array_equal(np.packbits(d["edges"].astype(bool)), g["f%d_edges%d" % (idx, b)])
            assert np.array_equal(np.packbits(d["vert"].astype(bool)), g["f%d_vert%d" % (idx, b)])


@pytest.mark.parametrize("sigma,bf,variant", [(2.0, 3, 0), (2.5, 1, 0), (2.0, 5, 1), (1.5, 3, 2), (3.0, 3, 0)])
def test_stage_images_match_oracle(dev, tiny2_cm, sigma, bf, variant):
    obj, mat, names, _, _ = _obj(tiny2_cm, "KR", dev)
    c = "chrA"
    cid = obj.unbalLib.ensure(c)
    M = float(np.quantile(mat.positive_population(c), 0.96))
    w = cn.gaussian_weights(sigma)
    for idx in (1, 4, 11):
        fr = pl.load_frame(mat, c, idx)
        G = im.saturate(fr["D"], M)
        for b in (0, 3, 5):
            d = dev.debug_stages(cid, M, sigma, bf, idx, b, variant)
            gray = im.gray_image(G, im.BRIGHTNESS[b], bf, variant)
            st = cn.canny(gray, sigma, w, None, return_stages=True)
            assert d["S"] == fr["S"]
            assert np.array_equal(d["gray"], gray)
            assert np.array_equal(d["smoothed"], st["smoothed"])
            assert np.array_equal(d["mag"], st["mag"])
            assert np.array_equal(d["low"].astype(bool), st["low"])
            assert np.array_equal(d["edges"].astype(bool), st["edges"])
            assert np.array_equal(d["vert"], sg.vertical_line(st["edges"]))


def _raw_oracle(mat, c, M, minH, maxW, sigma, bf, frames):
    w = cn.gaussian_weights(sigma)
    out = {}
    for idx in frames:
        fr = pl.load_frame(mat, c, idx)
        if fr is None:
            out[idx] = None
            continue
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
        out[idx] = rows
    return out


@pytest.mark.parametrize("cfg,norm,minH,maxW", [("tiny", "None", 10, 8), ("tiny2", "KR", 10, 8), ("tiny", "None", 5, 12)])
def test_raw_candidates_match_oracle(dev, cfg, norm, minH, maxW):
    cm = synth.make_config(cfg)
    obj, mat, names, _, _ = _obj(cm, norm, dev)
    for c in names:
        cid = obj.unbalLib.ensure(c)
        mp = np.quantile(mat.positive_population(c), [0.95, 0.99])
        cand = dev.extract(cid, mp, 2.0, minH, maxW, 3)
        nfr = pl.n_frames(mat.nbins(c))
        for qi in range(2):
            want = _raw_oracle(mat, c, mp[qi], minH, maxW, 2.0, 3, range(nfr))
            for idx in range(nfr):
                sel = cand[(cand["q"] == qi) & (cand["frame"] == idx)]
                got = [(int(r["bright"]), int(r["x"]), int(r["y"]), int(r["w"]), int(r["h"]), int(r["pos1"]),
                        int(r["pos2"]), int(r["pos3"]), int(r["pos4"])) for r in sel]
                assert got == (want[idx] or []), (c, qi, idx)
        # sharding by window range returns the same records
        a = dev.extract(cid, mp, 2.0, minH, maxW, 3, frame_lo=0, frame_hi=nfr // 2)
        b = dev.extract(cid, mp, 2.0, minH, maxW, 3, frame_lo=nfr // 2, frame_hi=nfr)
        both = np.concatenate([a, b])
        both = both[np.lexsort((both["k"], both["ud"], both["bright"], both["frame"], both["q"]))]
        assert np.array_equal(both, cand)


def _cmp_golden(unf, fil, g):
    for df, pre in ((unf, "unf_"), (fil, "fil_")):
        assert len(df) == len(g[pre + "pos1"])
        for k in ("pos1", "pos2", "pos3", "pos4", "length", "width"):
            assert df[k].astype(np.int64).tolist() == g[pre + k].tolist(), (pre, k)
        assert df["chr"].astype(str).tolist() == g[pre + "chr"].tolist()
        assert df["maxpixel"].tolist() == g[pre + "maxpixel"].tolist()
        for k in ("Mean", "pvalue", "Stripiness"):
            np.testing.assert_allclose(df[k].to_numpy(dtype=float), g[pre + k], rtol=RTOL, atol=0, equal_nan=True)


def test_compute_matches_reference_tiny(dev, tiny_cm, golden_dir):
    g = np.load(os.path.join(golden_dir, "tiny_compute.npz"))
    unf, fil, st = cp.run_stages(tiny_cm, norm="None", maxpixel="0.95,0.96,0.97,0.98,0.99", device=dev,
                                 return_stages=True)
    _cmp_golden(unf, fil, g)
    assert np.array_equal(st["MP"]["T1"], g["MP_T1"])
    np.testing.assert_allclose(st["EV"]["T1"], g["EV_T1"], rtol=1e-12)
    for name, t in zip(BG, st["bg"]):
        np.testing.assert_allclose(t[:, ::25], g["bg_" + name], rtol=1e-12, atol=1e-12, equal_nan=True)
    for i, t in enumerate(st["per_q"]):
        for k in ("pos1", "pos2", "pos3", "pos4", "num", "x", "y", "h", "w"):
            assert t[k].astype(np.int64).tolist() == g["q%d_%s" % (i, k)].tolist(), (i, k)
        np.testing.assert_allclose(t["pvalue"].to_numpy(dtype=float), g["q%d_pvalue" % i], rtol=RTOL)


def test_compute_matches_reference_tiny2_kr(dev, tiny2_cm, golden_dir):
    g = np.load(os.path.join(golden_dir, "tiny2_compute.npz"))
    unf, fil = cp.run_stages(tiny2_cm, norm="KR", maxpixel="0.95,0.98", device=dev)
    _cmp_golden(unf, fil, g)


def test_compute_subset_of_chromosomes(dev, tiny2_cm):
    """-k chrB: the background still samples every chromosome of the file (getStripe.py:242,276)."""
    unf, fil = cp.run_stages(tiny2_cm, norm="KR", chrom="chrB", maxpixel="0.97", device=dev)
    ounf, ofil = driver.compute(tiny2_cm, norm="KR", chrom="chrB", maxpixel="0.97")
    assert len(unf) == len(ounf["pos1"]) > 0
    for k in ("pos1", "pos2", "pos3", "pos4"):
        assert unf[k].astype(np.int64).tolist() == list(ounf[k])
    for k in ("Mean", "pvalue", "Stripiness"):
        np.testing.assert_allclose(unf[k].to_numpy(dtype=float), np.array(ounf[k], float), rtol=RTOL, equal_nan=True)


@pytest.mark.parametrize("fast", ["auto", "1"])
def test_score_matches_reference(dev, tiny2_cm, golden_dir, fast, monkeypatch):
    """`stripenn score` against the unmodified reference's own output; fast = "1" sends the regular
    stripes through k_score_fast (a table this small stays with the warp kernel otherwise)."""
    if fast == "1":
        monkeypatch.setenv("SNN_SCORE_FAST", "1")
    dev.score_routing()
    g = np.load(os.path.join(golden_dir, "tiny2_score.npz"))
    df = pd.DataFrame({"chr": g["chr"], "pos1": g["pos1"], "pos2": g["pos2"], "chr2": g["chr2"],
                       "pos3": g["pos3"], "pos4": g["pos4"]})
    res = score_table(tiny2_cm, df, norm="KR", device=dev)
    for k in ("pvalue_added", "Stripiness_added", "O_Mean_added", "O_Sum_added", "O/E_Mean_added", "O/E_Total_added"):
        np.testing.assert_allclose(res[k].to_numpy(dtype=float), g[k.replace("/", "_")], rtol=RTOL, atol=0,
                                   equal_nan=True)
    routed = dev.score_routing()
    assert (routed[0] > 0) == (fast == "1"), routed


def test_small_and_empty_windows(dev):
    """Ragged inputs: a chromosome shorter than one window, windows inside an all-zero gap
    (skipped: <= 10 non-empty columns) and a chromosome without pixels."""
    cm = synth.make_map([("s", 230 * 5000 - 17), ("e", 900 * 5000)], 5000, seed=11, only=["s"])
    cm.chroms["e"] = type(cm.chroms["s"])("e", 900 * 5000, 900, np.zeros(0, np.int32), np.zeros(0, np.int32),
                                          np.zeros(0, np.int32))
    mat = OracleMatrix(cm, False)
    lib = gs.DeviceMatrix(cm, False, dev)
    cid = lib.ensure("s")
    mp = np.quantile(mat.positive_population("s"), [0.95])
    cand = dev.extract(cid, mp, 2.0, 10, 8, 3)
    want = _raw_oracle(mat, "s", mp[0], 10, 8, 2.0, 3, range(pl.n_frames(230)))
    got = {}
    for r in cand:
        got.setdefault(int(r["frame"]), []).append((int(r["bright"]), int(r["x"]), int(r["y"]), int(r["w"]),
                                                    int(r["h"]), int(r["pos1"]), int(r["pos2"]), int(r["pos3"]),
                                                    int(r["pos4"])))
    for idx, rows in want.items():
        assert got.get(idx, []) == (rows or [])
    cid = lib.ensure("e")
    assert len(dev.extract(cid, np.array([5.0]), 2.0, 10, 8, 3)) == 0
    assert np.isnan(dev.quantiles(cid, [0.5])).all()
    assert np.array_equal(dev.expected(cid), np.zeros(400))


@pytest.mark.parametrize("nkeep,sigma", [(11, 2.0), (13, 2.0), (16, 2.0), (17, 2.0), (18, 2.5), (21, 2.5), (30, 2.0)])
def test_narrow_windows_match_oracle(dev, nkeep, sigma, monkeypatch):
    """Windows with fewer non-empty columns than the Gaussian kernel is wide (S < 2 radius + 1:
    every pixel is a border pixel for both filter axes) and just above.  Smoothed image and
    magnitude bit-exact, for both axis-1 kernels."""
    base = synth.make_map([("n", 300 * 5000)], 5000, seed=5)
    px = base.chroms["n"]
    keep_bins = np.sort(np.random.default_rng(nkeep).choice(np.arange(20, 280), size=nkeep, replace=False))
    ok = np.isin(px.bin1, keep_bins) & np.isin(px.bin2, keep_bins)
    # make sure every kept bin has a pixel (its diagonal), so exactly nkeep columns are non-empty
    b1 = np.concatenate([px.bin1[ok], keep_bins.astype(np.int32)])
    b2 = np.concatenate([px.bin2[ok], keep_bins.astype(np.int32)])
    ct = np.concatenate([px.count[ok], np.full(nkeep, 3, np.int32)])
    o = np.lexsort((b2, b1))
    first = np.ones(len(o), bool)
    first[1:] = (b1[o][1:] != b1[o][:-1]) | (b2[o][1:] != b2[o][:-1])
    o = o[first]
    base.chroms["n"] = type(px)("n", px.size_bp, px.nbins, b1[o].astype(np.int32), b2[o].astype(np.int32),
                                ct[o].astype(np.int32))
    mat = OracleMatrix(base, False)
    lib = gs.DeviceMatrix(base, False, dev)
    cid = lib.ensure("n")
    M = 4.0
    fr = pl.load_frame(mat, "n", 0)
    assert fr["S"] == nkeep
    w = cn.gaussian_weights(sigma)
    G = im.saturate(fr["D"], M)
    for b in (0, 5):
        d = dev.debug_stages(cid, M, sigma, 3, 0, b)
        gray = im.gray_image(G, im.BRIGHTNESS[b], 3)
        st = cn.canny(gray, sigma, w, None, return_stages=True)
        assert d["S"] == nkeep
        assert np.array_equal(d["smoothed"], st["smoothed"]), (nkeep, sigma, b)
        assert np.array_equal(d["mag"], st["mag"])
        monkeypatch.setenv("SNN_GAUSS_HD", "1")            # the direct-store axis-1 kernel (radius 8 only)
        d = dev.debug_stages(cid, M, sigma, 3, 0, b)
        monkeypatch.delenv("SNN_GAUSS_HD")
        assert np.array_equal(d["smoothed"], st["smoothed"]) and np.array_equal(d["mag"], st["mag"])


def test_equal_maxpixel_values_share_one_search(dev, tiny_cm, monkeypatch):
    """Raw-count quantiles that coincide: one image chain per distinct value, candidates replicated
    — identical to searching every value separately (STRIPENN_B200_NO_MP_DEDUP=1)."""
    obj, mat, names, _, _ = _obj(tiny_cm, "None", dev)
    cid = obj.unbalLib.ensure("T1")
    qs = [0.95, 0.9501, 0.97, 0.9701, 0.9702, 0.99]
    mp = np.quantile(mat.positive_population("T1"), qs)
    assert len(np.unique(mp)) < len(mp)                       # integers: neighbours coincide
    a = dev.extract(cid, mp, 2.0, 10, 8, 3)
    monkeypatch.setenv("STRIPENN_B200_NO_MP_DEDUP", "1")
    b = dev.extract(cid, mp, 2.0, 10, 8, 3)
    assert len(a) > 0 and np.array_equal(a, b)
    assert set(np.unique(a["q"]).tolist()) == set(range(len(qs)))


def test_load_errors(dev):
    with pytest.raises(_lib.StripennB200Error) as e:
        dev.load_chrom(0, 100, 5000, 500000, [5, 3], [6, 4], [1, 1])       # not sorted
    assert "sorted" in str(e.value)
    with pytest.raises(_lib.StripennB200Error):
        dev.load_chrom(0, 100, 5000, 500000, [5], [3], [1])                # lower triangle
    with pytest.raises(_lib.StripennB200Error) as e:
        dev.quantiles(7, [0.5])
    assert e.value.code == _lib.ESTATE


def test_generic_and_specialised_kernels_agree(dev, tiny2_cm, monkeypatch):
    """SNN_FORCE_GENERIC=1 selects the unspecialised image kernels (image.cu); the default path
    uses image_fast.cu.  Both must give the same bits."""
    from stripenn_b200.engine import Device
    monkeypatch.setenv("SNN_FORCE_GENERIC", "1")
    gen = Device(0)
    monkeypatch.delenv("SNN_FORCE_GENERIC")
    try:
        mat = OracleMatrix(tiny2_cm, "KR")
        c = "chrB"
        px = tiny2_cm.chroms[c]
        w = tiny2_cm.weight_vector(c, "KR")
        for d in (dev, gen):
            d.load_chrom(1, px.nbins, 5000, px.size_bp, px.bin1, px.bin2, px.count, w)
        mp = np.quantile(mat.positive_population(c), [0.95, 0.97, 0.99])
        a = dev.extract(1, mp, 2.0, 10, 8, 3)
        b = gen.extract(1, mp, 2.0, 10, 8, 3)
        assert len(a) > 0 and np.array_equal(a, b)
        for idx in (0, 3):
            for bi in (1, 4):
                x = dev.debug_stages(1, mp[1], 2.0, 3, idx, bi)
                y = gen.debug_stages(1, mp[1], 2.0, 3, idx, bi)
                for k in ("gray", "smoothed", "mag", "low", "edges", "vert"):
                    assert np.array_equal(x[k], y[k]), (idx, bi, k)
    finally:
        gen.close()


def test_fused_gaussian_equals_literal_form(dev, tiny2_cm, monkeypatch):
    """k_gauss_vm / k_gauss_hm accumulate with FMAs and fall back to scipy's literal
    multiply-then-add form next to a float32 rounding tie (gauss_near_tie, image_fast.cu).
    SNN_GAUSS_TIE_K=0x10000000 forces the literal form for every output: the smoothed images of
    every window and brightness of a chromosome (23 M accumulations, a handful of real near-ties)
    must be the same bits either way."""
    obj, mat, names, _, _ = _obj(tiny2_cm, "KR", dev)
    cid = obj.unbalLib.ensure("chrA")
    M = float(np.quantile(mat.positive_population("chrA"), 0.97))
    nwin = pl.n_frames(mat.nbins("chrA"))
    fused = {}
    for idx in range(nwin):
        for b in range(6):
            fused[idx, b] = dev.debug_stages(cid, M, 2.0, 3, idx, b)["smoothed"]
    monkeypatch.setenv("SNN_GAUSS_TIE_K", str(0x10000000))
    npx = 0
    for (idx, b), sm in fused.items():
        lit = dev.debug_stages(cid, M, 2.0, 3, idx, b)["smoothed"]
        assert np.array_equal(sm, lit), (idx, b)
        npx += sm.size
    assert npx > 5_000_000
    # and the whole search of the chromosome agrees as well
    mp = np.quantile(mat.positive_population("chrA"), [0.95, 0.99])
    lit_c = dev.extract(cid, mp, 2.0, 10, 8, 3)
    monkeypatch.delenv("SNN_GAUSS_TIE_K")
    assert np.array_equal(dev.extract(cid, mp, 2.0, 10, 8, 3), lit_c)


def test_direct_store_gaussian_kernel_equals_default(dev, tiny2_cm, monkeypatch):
    """SNN_GAUSS_HD=1 selects k_gauss_hd (axis-1 pass that divides and stores from registers, no
    second shared-memory tile) instead of k_gauss_hm: same bits for every window and brightness of
    a chromosome, with the tie check on and with the literal form forced, and the same candidates."""
    obj, mat, names, _, _ = _obj(tiny2_cm, "KR", dev)
    cid = obj.unbalLib.ensure("chrB")
    M = float(np.quantile(mat.positive_population("chrB"), 0.96))
    nwin = pl.n_frames(mat.nbins("chrB"))
    want = {(idx, b): dev.debug_stages(cid, M, 2.0, 3, idx, b) for idx in range(nwin) for b in (0, 2, 5)}
    mp = np.quantile(mat.positive_population("chrB"), [0.95, 0.99])
    cand = dev.extract(cid, mp, 2.0, 10, 8, 3)
    monkeypatch.setenv("SNN_GAUSS_HD", "1")
    for tie in (None, str(0x10000000)):
        if tie:
            monkeypatch.setenv("SNN_GAUSS_TIE_K", tie)
        for (idx, b), w in want.items():
            d = dev.debug_stages(cid, M, 2.0, 3, idx, b)
            S = w["S"]
            assert d["S"] == S
            assert np.array_equal(d["smoothed"][:S, :S], w["smoothed"][:S, :S]), (idx, b, tie)
            assert np.array_equal(d["edges"], w["edges"]) and np.array_equal(d["vert"], w["vert"])
    monkeypatch.delenv("SNN_GAUSS_TIE_K")
    assert len(cand) > 0 and np.array_equal(dev.extract(cid, mp, 2.0, 10, 8, 3), cand)
    # other radii have no direct-store instance: the switch must be harmless there
    a = dev.debug_stages(cid, M, 2.5, 3, 1, 1)["smoothed"]
    monkeypatch.delenv("SNN_GAUSS_HD")
    assert np.array_equal(a, dev.debug_stages(cid, M, 2.5, 3, 1, 1)["smoothed"])


@pytest.mark.parametrize("cfg,norm", [("tiny2", "KR"), ("tiny1k", "None"), ("tiny10k", "weight")])
def test_warp_and_cta_scoring_kernels_agree(dev, cfg, norm, monkeypatch):
    """snn_score routes in-band windows to the warp-per-stripe kernel (score_warp.cu) and the
    rest to the CTA-per-stripe kernel (score.cu); SNN_FORCE_GENERIC=1 sends everything to the
    latter.  Same stripes (random ones, tall ones, ones clamped at both chromosome ends, NaN-
    weight bins under KR) through both, one multi-chromosome call each: counts-based outputs
    identical, sums within 1e-12."""
    from stripenn_b200.engine import Device
    cm = synth.make_config(cfg)
    bal = cp.resolve_norm(cm, norm, "compute", False)
    monkeypatch.setenv("SNN_FORCE_GENERIC", "1")
    gen = Device(0)
    monkeypatch.delenv("SNN_FORCE_GENERIC")
    try:
        r = cm.binsize
        names = [c for c in cm.chromnames if c in cm.chroms]
        tab = synth.make_stripe_table(cm, 400, chroms=names)
        pos = np.stack([tab["pos1"], tab["pos2"], tab["pos3"], tab["pos4"]], axis=1).astype(np.int64)
        chrs = np.asarray(tab["chr"]).astype(str)
        rng = np.random.default_rng(5)
        extra, extra_c = [], []
        for c in names:                       # clamps at both ends, tall and wide stripes, off-diagonal ones
            nb, size = cm.nbins(c), cm.chroms[c].size_bp
            for a, w, hh in ((0, 3, 30), (1, 5, 70), (2, 8, 150), (nb - 40, 4, 39), (nb - 200, 6, 199),
                             (nb // 2, 2, 310), (nb // 2, 40, 60), (nb // 3, 70, 20)):
                a = int(max(a, 0))
                p1, p2 = a * r + 1, min((a + w) * r, size)
                extra.append((p1, p2, p1, min((a + hh) * r, size)))          # down
                extra.append((p1, p2, max((a + w - hh) * r + 1, 1), p2))      # up
                extra.append((p1, p2, max(p1 - 7 * r, 1), min(p2 + 9 * r, size)))  # neither (stale rows)
                extra_c += [c] * 3
        pos = np.concatenate([pos, np.array(extra, dtype=np.int64)])
        chrs = np.concatenate([chrs, np.array(extra_c)])
        perm = rng.permutation(len(pos))
        pos, chrs = pos[perm], chrs[perm]
        outs = []
        for d in (dev, gen):
            lib = gs.DeviceMatrix(cm, bal, d)
            obj = gs.getStripe(lib, r, 10, 8, 2.0, cm.chromnames, names, cm.chromsizes,
                               [cm.chroms[c].size_bp for c in names], 1, 3, 123456789, 0)
            bg = obj.nulldist()
            ex = obj.mpmean()
            df = pd.DataFrame({"chr": chrs, "pos1": pos[:, 0], "pos2": pos[:, 1], "pos3": pos[:, 2], "pos4": pos[:, 3]})
            pv = np.array(obj.pvalue(*bg, df), dtype=float)
            st, om, ot = (np.array(v, dtype=float) for v in obj.scoringstripes(df, ex))
            gm, gsum = (np.array(v, dtype=float) for v in obj.getMean(df))
            outs.append((pv, st, om, ot, gm, gsum))
        a, b = outs
        assert np.isfinite(a[0]).sum() > 300 and np.isfinite(a[1]).sum() > 300
        np.testing.assert_array_equal(a[0], b[0])                             # p-values: ratios of counts
        for x, y in zip(a[1:], b[1:]):
            np.testing.assert_allclose(x, y, rtol=1e-12, atol=0, equal_nan=True)
    finally:
        gen.close()


@pytest.mark.parametrize("cfg,norm", [("tiny1k", "None"), ("tiny10k", "weight")])
def test_other_resolutions_match_oracle(dev, cfg, norm):
    """1 kb (bs = 50: 110-column scoring windows, background rows that wrap around the fetched
    unit) and 10 kb (bs = 5, uneven bu/bd split) against the oracle, end to end."""
    cm = synth.make_config(cfg)
    unf, fil, st = cp.run_stages(cm, norm=norm, maxpixel="0.96,0.99", device=dev, return_stages=True)
    ounf, ofil, ost = driver.compute(cm, norm=norm, maxpixel="0.96,0.99", return_stages=True)
    for a, b in zip(st["bg"], ost["bg"]):
        assert a.shape == b.shape
        np.testing.assert_allclose(a, b, rtol=1e-12, atol=1e-12, equal_nan=True)
    assert len(unf) == len(ounf["pos1"]) > 0
    for k in ("pos1", "pos2", "pos3", "pos4"):
        assert unf[k].astype(np.int64).tolist() == list(ounf[k]), k
    assert unf["maxpixel"].tolist() == list(ounf["maxpixel"])
    for k in ("Mean", "pvalue", "Stripiness"):
        np.testing.assert_allclose(unf[k].to_numpy(dtype=float), np.array(ounf[k], float), rtol=RTOL, atol=0,
                                   equal_nan=True)
    assert fil["pos1"].astype(np.int64).tolist() == list(ofil["pos1"])
    # score mode on the same map
    tab = synth.make_stripe_table(cm, 60)
    df = pd.DataFrame({"chr": tab["chr"], "pos1": tab["pos1"], "pos2": tab["pos2"], "chr2": tab["chr"],
                       "pos3": tab["pos3"], "pos4": tab["pos4"]})
    res = score_table(cm, df, norm=norm, device=dev)
    o = driver.score(cm, tab, norm=norm)
    for k in o:
        np.testing.assert_allclose(res[k].to_numpy(dtype=float), np.array(o[k], float), rtol=RTOL, atol=0,
                                   equal_nan=True)


def test_host_threads_give_identical_tables(tiny2_cm):
    """streams=2: two host threads / contexts on one GPU (ThreadComm) vs the single-context run."""
    a_unf, a_fil = cp.run_stages(tiny2_cm, norm="KR", maxpixel="0.95,0.98", streams=1)
    b_unf, b_fil = cp.run_stages(tiny2_cm, norm="KR", maxpixel="0.95,0.98", streams=2)
    assert len(a_unf) > 0
    assert a_unf.reset_index(drop=True).equals(b_unf.reset_index(drop=True))
    assert a_fil.reset_index(drop=True).equals(b_fil.reset_index(drop=True))


def test_window_ranges_of_one_chromosome_on_three_contexts(tiny_cm, golden_dir):
    """plan_units cuts a single chromosome into window ranges (BASELINE configs 1 and 4 on several
    GPUs); here three contexts on one GPU: tables identical to the single-context run and to the
    reference's golden tables."""
    g = np.load(os.path.join(golden_dir, "tiny_compute.npz"))
    a_unf, a_fil = cp.run_stages(tiny_cm, norm="None", streams=1)
    b_unf, b_fil = cp.run_stages(tiny_cm, norm="None", streams=3)
    assert a_unf.reset_index(drop=True).equals(b_unf.reset_index(drop=True))
    assert a_fil.reset_index(drop=True).equals(b_fil.reset_index(drop=True))
    for k in ("pos1", "pos2", "pos3", "pos4"):
        assert b_unf[k].astype(np.int64).tolist() == g["unf_" + k].tolist()
    assert b_unf["maxpixel"].tolist() == g["unf_maxpixel"].tolist()
    for k in ("Mean", "pvalue", "Stripiness"):
        np.testing.assert_allclose(b_unf[k].to_numpy(dtype=float), g["unf_" + k], rtol=RTOL, equal_nan=True)


@pytest.mark.parametrize("cfg,norm", [("tiny2", "KR"), ("tiny10k", "None")])
def test_stripes_at_chromosome_ends_match_oracle(dev, cfg, norm):
    """flanks clamped to bp 1 / the chromosome length, bin 1 / Nbins-1 in
    scoringstripes, stripes touching the first and last bins — each stripe scored by the oracle
    on its own; where the oracle (like the reference) raises, the CUDA path answers NaN."""
    cm = synth.make_config(cfg)
    bal = cp.resolve_norm(cm, norm, "compute", False)
    names = [c for c in cm.chromnames if c in cm.chroms]
    r = cm.binsize
    rows, chrs = [], []
    for c in names:
        nb, size = cm.nbins(c), cm.chroms[c].size_bp
        for a, w, hh in ((0, 3, 30), (1, 5, 70), (2, 8, 150), (3, 2, 12), (nb - 40, 4, 39), (nb - 200, 6, 199),
                         (nb - 13, 3, 12), (nb - 31, 9, 30), (nb - 120, 5, 119)):
            p1, p2 = a * r + 1, min((a + w) * r, size)
            rows.append((p1, p2, p1, min((a + hh) * r, size)))               # down
            if a + w - hh >= 0:
                rows.append((p1, p2, (a + w - hh) * r + 1, p2))              # up
            chrs += [c] * (len(rows) - len(chrs))
    pos = np.array(rows, dtype=np.int64)
    obj, mat, names_, all_names, all_sizes = _obj(cm, norm, dev)
    bg = obj.nulldist()
    EV = obj.mpmean()
    df = pd.DataFrame({"chr": chrs, "pos1": pos[:, 0], "pos2": pos[:, 1], "pos3": pos[:, 2], "pos4": pos[:, 3]})
    got = dict(pvalue=np.array(obj.pvalue(*bg, df), float))
    got["stripiness"], got["oe_mean"], got["oe_total"] = (np.array(v, float) for v in obj.scoringstripes(df, EV))
    got["o_mean"], got["o_sum"] = (np.array(v, float) for v in obj.getMean(df))
    bgo = bgm.null_distribution(mat, all_names, all_sizes, 123456789)
    EVo = bgm.expected_by_distance(mat, names)
    want = {k: np.full(len(pos), np.nan) for k in got}
    from snn_oracle import scoring as osc
    raised = 0
    for i in range(len(pos)):
        one = {"chr": [chrs[i]], "pos1": [int(pos[i, 0])], "pos2": [int(pos[i, 1])], "pos3": [int(pos[i, 2])],
               "pos4": [int(pos[i, 3])]}
        for keys, fn in ((("pvalue",), lambda: (osc.pvalues(mat, bgo, one),)),
                         (("stripiness", "oe_mean", "oe_total"), lambda: osc.stripiness(mat, one, EVo)),
                         (("o_mean", "o_sum"), lambda: osc.observed_mean(mat, one))):
            try:
                vals = fn()
            except (ValueError, IndexError):
                raised += 1
                continue
            for k, v in zip(keys, vals):
                want[k][i] = float(v[0])
    assert np.isfinite(want["pvalue"]).sum() > len(pos) // 2
    for k in got:
        np.testing.assert_allclose(got[k], want[k], rtol=RTOL, atol=0, equal_nan=True, err_msg=k)


@pytest.mark.parametrize("mode", ["a", "1"])
def test_fused_chain_equals_unfused_kernels(dev, tiny2_cm, monkeypatch, mode):
    """Image chain variants (image_fused.cu): SNN_FUSED=a runs k_gray_gv (gray + axis-0 Gaussian in one
    kernel) in front of the unfused axis-1 and Sobel kernels, SNN_FUSED=1 adds k_gh_nms (axis 1 +
    Sobel + NMS from TMA tiles), SNN_FUSED=0 is the five unfused kernels of image_fast.cu.  Every
    per-pass product (gray, smoothed, magnitude, masks) of every window and brightness of a
    chromosome, and the candidates of the whole chromosome, must be the same bits — with the tie
    check on and with the literal form forced for every output."""
    obj, mat, names, _, _ = _obj(tiny2_cm, "KR", dev)
    cid = obj.unbalLib.ensure("chrA")
    mp = np.quantile(mat.positive_population("chrA"), [0.95, 0.97, 0.99])
    nwin = pl.n_frames(mat.nbins("chrA"))
    monkeypatch.setenv("SNN_FUSED", mode)
    fused = {(idx, b): dev.debug_stages(cid, mp[1], 2.0, 3, idx, b) for idx in range(nwin) for b in range(6)}
    cand = dev.extract(cid, mp, 2.0, 10, 8, 3)
    monkeypatch.setenv("SNN_GAUSS_TIE_K", str(0x10000000))
    for (idx, b) in ((0, 0), (3, 2), (nwin - 1, 5)):
        lit = dev.debug_stages(cid, mp[1], 2.0, 3, idx, b)
        for k in ("gray", "smoothed", "mag", "low", "edges", "vert"):
            assert np.array_equal(fused[idx, b][k], lit[k]), (idx, b, k, "literal")
    monkeypatch.delenv("SNN_GAUSS_TIE_K")
    monkeypatch.setenv("SNN_FUSED", "0")
    for (idx, b), x in fused.items():
        y = dev.debug_stages(cid, mp[1], 2.0, 3, idx, b)
        assert x["S"] == y["S"]
        for k in ("gray", "smoothed", "mag", "low", "edges", "vert"):
            assert np.array_equal(x[k], y[k]), (idx, b, k)
    assert len(cand) > 0 and np.array_equal(dev.extract(cid, mp, 2.0, 10, 8, 3), cand)
