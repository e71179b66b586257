"""GPU (-m gpu): size-independent properties at a BASELINE.json-sized chromosome (config 1:
chr16 at 5 kb, 19,642 bins, 99 search windows, 5 maxpixel values) where the oracle is too slow
to run in full: exact scale invariance, window-range sharding == whole, idempotent dedup,
quantiles == numpy on the host population, and oracle parity on sampled windows."""
import os

import numpy as np
import pytest

from snn_oracle import pipeline as pl
from snn_oracle.matrix import OracleMatrix
from stripenn_b200 import compute as cp
from stripenn_b200 import getstripe as gs
from stripenn_b200 import synth

pytestmark = pytest.mark.gpu
QS = [0.95, 0.96, 0.97, 0.98, 0.99]


@pytest.fixture(scope="module")
def cfg1():
    return synth.make_config("cfg1")


def test_cfg1_properties(device, cfg1):
    for cid in list(device.loaded):
        device.unload_chrom(cid)
    cm = cfg1
    mat = OracleMatrix(cm, False)
    lib = gs.DeviceMatrix(cm, False, device)
    cid = lib.ensure("16")
    mp = device.quantiles(cid, QS)
    assert np.array_equal(mp, np.quantile(mat.positive_population("16"), QS))
    cand = device.extract(cid, mp, 2.0, 10, 8, 3)
    assert len(cand) > 500
    # sortedness of the emission order
    key = np.lexsort((cand["k"], cand["ud"], cand["bright"], cand["frame"], cand["q"]))
    assert np.array_equal(key, np.arange(len(cand)))
    # window-range shards (what a multi-GPU split of one chromosome does) == whole
    parts = [device.extract(cid, mp, 2.0, 10, 8, 3, frame_lo=lo, frame_hi=hi) for lo, hi in ((0, 13), (13, 50), (50, 99))]
    both = np.concatenate(parts)
    both = both[np.lexsort((both["k"], both["ud"], both["bright"], both["frame"], both["q"]))]
    assert np.array_equal(both, cand)
    # fused Gaussian accumulation (FMA + tie check) == the literal form, over ~1e9 accumulations
    os.environ["SNN_GAUSS_TIE_K"] = str(0x10000000)
    try:
        assert np.array_equal(device.extract(cid, mp, 2.0, 10, 8, 3), cand)
    finally:
        del os.environ["SNN_GAUSS_TIE_K"]
    # the direct-store axis-1 Gaussian kernel (A/B candidate) finds the same candidates
    os.environ["SNN_GAUSS_HD"] = "1"
    try:
        assert np.array_equal(device.extract(cid, mp, 2.0, 10, 8, 3), cand)
    finally:
        del os.environ["SNN_GAUSS_HD"]
    # oracle parity on a sample of windows (the full chromosome takes ~10 min on one core)
    for idx in (0, 17, 33, 98):
        fr = pl.load_frame(mat, "16", idx)
        for qi in (0, 4):
            if fr is None:
                want = []
            else:
                t = pl.stripe_search(fr, mp[qi], QS[qi], "16", 10, 8, 2.0, 3)
                want = sorted(zip(t["x"], t["y"], t["w"], t["h"]))
            sel = cand[(cand["q"] == qi) & (cand["frame"] == idx)]
            got_all = set((int(r["x"]), int(r["y"]), int(r["w"]), int(r["h"])) for r in sel)
            assert set(want) <= got_all      # survivors of the per-window dedup are among the raw candidates
    # exact scale invariance: doubling every count doubles M and leaves (M-D)/M bit-identical
    cm2 = synth.make_config("cfg1")
    for c in cm2.chroms.values():
        c.count = (c.count * 2).astype(np.int32)
    lib2 = gs.DeviceMatrix(cm2, False, device)
    device.unload_chrom(cid)
    cid2 = lib2.ensure("16")
    mp2 = device.quantiles(cid2, QS)
    assert np.array_equal(mp2, 2 * mp)
    cand2 = device.extract(cid2, mp2, 2.0, 10, 8, 3)
    for k in ("q", "frame", "bright", "ud", "k", "x", "y", "w", "h", "pos1", "pos2", "pos3", "pos4"):
        assert np.array_equal(cand[k], cand2[k]), k
    assert np.array_equal(cand2["total"], 2 * cand["total"])
    device.unload_chrom(cid2)


def test_cfg1_end_to_end_invariants(device, cfg1):
    for cid in list(device.loaded):
        device.unload_chrom(cid)
    unf, fil = cp.run_stages(cfg1, norm="None", chrom="16", maxpixel="0.95,0.96,0.97,0.98,0.99", device=device)
    assert len(unf) > 50 and 0 < len(fil) <= len(unf)
    assert (fil["pvalue"] < 0.1).all()
    s = fil["Stripiness"].to_numpy()
    assert np.all(s[:-1] >= s[1:])
    assert ((unf["pvalue"] >= 0.001 - 1e-12) & (unf["pvalue"] <= 1)).all()
    assert (unf["length"] == unf["pos4"] - unf["pos3"] + 1).all()
    # planted stripes are found: most planted anchors have a called stripe within 3 bins
    planted = cfg1.planted["16"]
    x1 = (unf["pos1"].to_numpy() - 1) // 5000
    x2 = unf["pos2"].to_numpy() // 5000
    hit = sum(bool(np.any((x1 - 4 <= p["anchor"]) & (p["anchor"] <= x2 + 4))) for p in planted)
    assert hit >= 0.5 * len(planted)
    # RemoveRedundant is idempotent on its own output (no overlapping pair is left)
    for cid in list(device.loaded):
        device.unload_chrom(cid)
