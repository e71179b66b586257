"""SURVEY.md §8f "next" rows: --mask (getStripe.py:612-639, 701-707) and -s
(getStripe.getQuantile_slow, getStripe.py:107-158).

CPU part: the oracle against the golden fixtures produced by the UNMODIFIED reference
(tests/golden/make_golden_next.py).  GPU part (-m gpu): the CUDA path, through the C ABI,
against the same fixtures and against the oracle."""
import os

import numpy as np
import pandas as pd
import pytest

from snn_oracle import background as bgm
from snn_oracle import pipeline as pl
from snn_oracle import scoring as sc
from snn_oracle.matrix import OracleMatrix
from stripenn_b200 import compute as cp
from stripenn_b200 import synth

RTOL = 1e-5
SLOW3 = [("s1", 700 * 5000 + 100), ("s2", 600 * 5000 + 200), ("s3", 650 * 5000 + 300), ("s3b", 520 * 5000 + 400),
         ("s4", 6100 * 5000 - 77), ("s5", 800 * 5000 + 50)]


@pytest.fixture(scope="module")
def mask_golden(golden_dir):
    return np.load(os.path.join(golden_dir, "tiny2_mask.npz"))


@pytest.fixture(scope="module")
def slow_golden(golden_dir):
    return np.load(os.path.join(golden_dir, "slow_quant.npz"))


@pytest.fixture(scope="module")
def slow3_cm():
    return synth.make_map(SLOW3, 5000, seed=11, balanced=True)


def _mask_table(g):
    pos = g["pos"]
    return {"chr": [str(c) for c in g["chr"]], "pos1": pos[:, 0].tolist(), "pos2": pos[:, 1].tolist(),
            "pos3": pos[:, 2].tolist(), "pos4": pos[:, 3].tolist()}


# ------------------------------------------------------------------------------------------- CPU
def test_oracle_mask_matches_reference(tiny2_cm, mask_golden):
    g = mask_golden
    mat = OracleMatrix(tiny2_cm, "KR")
    EV = {c: g["EV_" + c] for c in tiny2_cm.chroms}
    tab = _mask_table(g)
    for k, mask in enumerate(g["masks"]):
        s, m, t = sc.stripiness(mat, tab, EV, str(mask), on_index_error="nan")
        raises = g["m%d_raises" % k]
        want = g["m%d_values" % k]
        assert (~raises).sum() > 0
        assert np.all(np.isnan(np.array(s, float)[raises]))
        ok = ~raises
        for got, ref in zip((s, m, t), want):
            np.testing.assert_allclose(np.array(got, float)[ok], ref[ok], rtol=1e-12, equal_nan=True)
        # and the oracle raises where the reference does
        for i in np.nonzero(raises)[0]:
            one = {c: [v[int(i)]] for c, v in tab.items()}
            with pytest.raises(IndexError):
                sc.stripiness(mat, one, EV, str(mask))
    assert sum(int(g["m%d_raises" % k].sum()) for k in range(len(g["masks"]))) > 0


def test_oracle_slow_quantiles_match_reference(tiny2_cm, slow3_cm, slow_golden):
    g = slow_golden
    qs = g["qs"].tolist()
    differs = 0
    for norm in ("None", "KR"):
        mat = OracleMatrix(tiny2_cm, False if norm == "None" else norm)
        MP = pl.quantiles_slow(mat, list(tiny2_cm.chroms), qs)
        for c in tiny2_cm.chroms:
            assert np.array_equal(MP[c], g["tiny2_%s_%s" % (norm, c)]), (norm, c)
            differs += int(not np.array_equal(g["tiny2_%s_%s" % (norm, c)], g["tiny2_%s_%s_fast" % (norm, c)]))
        mat = OracleMatrix(slow3_cm, False if norm == "None" else norm)
        MP = pl.quantiles_slow(mat, [c[0] for c in SLOW3], qs)
        for c, _ in SLOW3:
            assert np.array_equal(MP[c], g["slow3_%s_%s" % (norm, c)]), (norm, c)
        assert np.array_equal(pl.quantiles_slow(mat, ["s5"], qs)["s5"], g["slow3_%s_only_s5" % norm])
    assert differs > 0      # -s is NOT the default path's answer


def test_slow_selection_geometry(slow3_cm):
    """Host decomposition of the genome-wide slices into per-chromosome sources."""
    from stripenn_b200 import getstripe as gs

    class NoDevice:
        dev = None
        cm = slow3_cm
    o = gs.getStripe.__new__(gs.getStripe)
    o.unbalLib = NoDevice()
    jobs, src = o.slow_selection([c[0] for c in SLOW3])
    assert jobs == [c[0] for c in SLOW3]
    by_job = {}
    for name, job, s, w, F, c0, c1 in src:
        by_job.setdefault(jobs[job], []).append((name, s, w, F, c0, c1))
    assert by_job["s1"] == [("s1", 1, 25000000 // 701, 700, 1, 701)]
    # s4: sum(ceil) = 2475, ceil(cum/r) + 1 = 2473 -> starts 2 bins inside s3b
    names4 = [x[0] for x in by_job["s4"]]
    assert names4 == ["s3b", "s4"]
    s3b = by_job["s4"][0]
    assert s3b[1] == 521 - 2 and s3b[4] == 519
    s4 = by_job["s4"][1]
    assert s4[1] == -2 and s4[2] == 25000000 // 6100 and s4[3] == 6099


# ------------------------------------------------------------------------------------------- GPU
def _obj(cm, norm, dev, chrom=("all",)):
    bal = cp.resolve_norm(cm, norm, "compute", False)
    all_names, all_sizes, names, sizes = cp.compute_chromosomes(cm, list(chrom))
    o = cp.make_object(cm, bal, dev, all_names, all_sizes, names, sizes, 10, 8, 2.0, 1, 3, 123456789, 0)
    return o, names


@pytest.fixture()
def dev(device):
    for cid in list(device.loaded):
        device.unload_chrom(cid)
    yield device
    device.set_mask(None)
    for cid in list(device.loaded):
        device.unload_chrom(cid)


@pytest.mark.gpu
@pytest.mark.parametrize("generic", [False, True])
def test_gpu_mask_matches_reference(dev, tiny2_cm, mask_golden, generic, monkeypatch):
    g = mask_golden
    if generic:
        monkeypatch.setenv("SNN_FORCE_GENERIC", "1")
        from stripenn_b200.engine import Device
        d = Device(0)
    else:
        d = dev
    try:
        obj, names = _obj(tiny2_cm, "KR", d)
        EV = {c: g["EV_" + c].tolist() for c in tiny2_cm.chroms}
        pos = g["pos"]
        df = pd.DataFrame({"chr": [str(c) for c in g["chr"]], "pos1": pos[:, 0], "pos2": pos[:, 1],
                           "pos3": pos[:, 2], "pos4": pos[:, 3]})
        base = np.array(obj.scoringstripes(df, EV, "0")[0])
        for k, mask in enumerate(g["masks"]):
            s, m, t = (np.array(v) for v in obj.scoringstripes(df, EV, str(mask)))
            raises = g["m%d_raises" % k]
            want = g["m%d_values" % k]
            assert np.all(np.isnan(s[raises])) and np.all(np.isnan(m[raises])) and np.all(np.isnan(t[raises]))
            ok = ~raises
            for got, ref in zip((s, m, t), want):
                np.testing.assert_allclose(got[ok], ref[ok], rtol=RTOL, equal_nan=True)
            assert np.sum(~np.isclose(s[ok], base[ok], rtol=1e-9, equal_nan=True)) > 0   # the mask did something
        # the mask is cleared afterwards, and getMean / p-values never see it
        again = np.array(obj.scoringstripes(df, EV, "0")[0])
        assert np.array_equal(again, base, equal_nan=True)
        # a mask on another chromosome name changes nothing
        other = np.array(obj.scoringstripes(df, EV, "nochr:1-1000000")[0])
        assert np.array_equal(other, base, equal_nan=True)
    finally:
        if generic:
            d.close()


@pytest.mark.gpu
def test_gpu_mask_in_compute_pipeline(dev, tiny2_cm):
    """run_stages with --mask == oracle driver with the same mask (rows the reference dies on: NaN)."""
    from snn_oracle import driver
    mask = "chrA:2000000-2060000"
    unf, fil = cp.run_stages(tiny2_cm, norm="KR", maxpixel="0.95,0.98", mask=mask, device=dev)
    ounf, _ = driver.compute(tiny2_cm, norm="KR", maxpixel="0.95,0.98", mask="0")
    assert list(unf["pos1"]) == list(ounf["pos1"]) and list(unf["pos4"]) == list(ounf["pos4"])
    mat = OracleMatrix(tiny2_cm, "KR")
    EV = bgm.expected_by_distance(mat, list(tiny2_cm.chroms))
    tab = {k: list(unf[k]) for k in ("chr", "pos1", "pos2", "pos3", "pos4")}
    s, _, _ = sc.stripiness(mat, tab, EV, mask, on_index_error="nan")
    np.testing.assert_allclose(np.array(unf["Stripiness"], float), np.array(s, float), rtol=RTOL, equal_nan=True)


@pytest.mark.gpu
def test_gpu_slow_quantiles_match_reference(dev, tiny2_cm, slow3_cm, slow_golden):
    g = slow_golden
    qs = g["qs"].tolist()
    for norm in ("None", "KR"):
        obj, names = _obj(tiny2_cm, norm, dev)
        MP = obj.getQuantile_slow(tiny2_cm, names, qs)
        for c in names:
            assert np.array_equal(MP[c], g["tiny2_%s_%s" % (norm, c)]), (norm, c)
        for cid in list(dev.loaded):
            dev.unload_chrom(cid)
        obj, names = _obj(slow3_cm, norm, dev)
        MP = obj.getQuantile_slow(slow3_cm, names, qs)
        for c in names:
            assert np.array_equal(MP[c], g["slow3_%s_%s" % (norm, c)]), (norm, c)
        for cid in list(dev.loaded):
            dev.unload_chrom(cid)
        obj, names = _obj(slow3_cm, norm, dev, chrom=("s5",))
        assert np.array_equal(obj.getQuantile_slow(slow3_cm, ["s5"], qs)["s5"], g["slow3_%s_only_s5" % norm])
        for cid in list(dev.loaded):
            dev.unload_chrom(cid)


@pytest.mark.gpu
def test_gpu_slow_compute_runs_and_differs(dev, tiny2_cm):
    from snn_oracle import driver
    unf, _ = cp.run_stages(tiny2_cm, norm="KR", maxpixel="0.95,0.98", slow=True, device=dev)
    ounf, _ = driver.compute(tiny2_cm, norm="KR", maxpixel="0.95,0.98", slow=True)
    assert list(unf["pos1"]) == list(ounf["pos1"]) and list(unf["pos2"]) == list(ounf["pos2"])
    assert list(unf["pos3"]) == list(ounf["pos3"]) and list(unf["pos4"]) == list(ounf["pos4"])
    for k in ("Mean", "pvalue", "Stripiness"):
        np.testing.assert_allclose(np.array(unf[k], float), np.array(ounf[k], float), rtol=RTOL, equal_nan=True)
