"""SURVEY.md §8f row 1 — `.cool` / `.mcool` ingestion (reference: cooler.Cooler(cool) at
stripenn.py:80, score.py:24, seeimage.py:34).  cooler/h5py/libhdf5 are not in the image, so the
adaptor ContactMap.from_cooler is driven through tests/fake_cooler.py, a stand-in with cooler's
public API (genome-wide bin ids, cis AND trans pixels in the pixel table, bins table with the
weight columns)."""
import sys

import numpy as np
import pandas as pd
import pytest

import fake_cooler
from stripenn_b200 import synth
from stripenn_b200.contactmap import ContactMap, open_contact_map

URI = "fake.mcool::resolutions/5000"


@pytest.fixture()
def cooler_module(monkeypatch, tiny2_cm):
    # trans pixels (chrA x chrB) that a real pixel table holds and stripenn never reads
    lo_b = tiny2_cm.nbins("chrA")
    trans = pd.DataFrame({"bin1_id": np.array([3, 3, 800, 2399], dtype=np.int64),
                          "bin2_id": np.array([lo_b + 1, lo_b + 700, lo_b, lo_b + 1499], dtype=np.int64),
                          "count": np.array([5, 1, 2, 9], dtype=np.int32)})
    fake_cooler.REGISTRY.clear()
    fake_cooler.REGISTRY[URI] = (tiny2_cm, trans)
    monkeypatch.setitem(sys.modules, "cooler", fake_cooler)
    yield fake_cooler
    fake_cooler.REGISTRY.clear()


def test_from_cooler_round_trip(cooler_module, tiny2_cm):
    cm = open_contact_map(URI)
    assert cm.binsize == tiny2_cm.binsize
    assert cm.chromnames == tiny2_cm.chromnames
    assert np.array_equal(cm.chromsizes, tiny2_cm.chromsizes)
    assert cm.bin_columns() == tiny2_cm.bin_columns()
    for name, want in tiny2_cm.chroms.items():
        got = cm.chroms[name]
        assert got.nbins == want.nbins and got.size_bp == want.size_bp
        # chromosome-local bins, trans pixels dropped, cooler's (bin1, bin2) order kept
        assert np.array_equal(got.bin1, want.bin1) and got.bin1.dtype == np.int32
        assert np.array_equal(got.bin2, want.bin2)
        assert np.array_equal(got.count, want.count)
        for col in tiny2_cm.weights:
            assert np.array_equal(cm.weights[col][name], tiny2_cm.weights[col][name], equal_nan=True)
            # divisive columns (KR, VC, SQRT_VC) are inverted, `weight` is not (cooler's rule)
            a, b = cm.weight_vector(name, col), tiny2_cm.weight_vector(name, col)
            assert np.array_equal(a, b, equal_nan=True)


def test_from_cooler_subset_and_unknown_uri(cooler_module):
    cm = ContactMap.from_cooler(URI, chroms=["chrB"])
    assert list(cm.chroms) == ["chrB"] and cm.chromnames == ["chrA", "chrB"]
    with pytest.raises(OSError):
        open_contact_map("missing.cool")


def test_oracle_agrees_on_a_cooler_backed_map(cooler_module, tiny2_cm):
    """The host layer sees the same matrix whichever way the map was opened."""
    from snn_oracle import pipeline as pl
    from snn_oracle.matrix import OracleMatrix
    a = OracleMatrix(open_contact_map(URI), "KR")
    b = OracleMatrix(tiny2_cm, "KR")
    qs = [0.95, 0.99]
    assert np.array_equal(pl.quantiles(a, ["chrB"], qs)["chrB"], pl.quantiles(b, ["chrB"], qs)["chrB"])
    assert np.array_equal(a.window("chrA", 100, 180, 90, 200), b.window("chrA", 100, 180, 90, 200), equal_nan=True)


@pytest.mark.gpu
def test_gpu_compute_from_a_cooler_uri(cooler_module, tiny2_cm, tmp_path):
    """`stripenn compute --cool X.mcool::resolutions/5000` end to end == the same run on the .npz."""
    from stripenn_b200 import compute as cp
    out1, out2 = str(tmp_path / "a"), str(tmp_path / "b")
    p = str(tmp_path / "tiny2.npz")
    tiny2_cm.save_npz(p)
    cp.compute(URI, out1, "KR", "all", 2.0, 10, 8, "0.95,0.98", 2, 0.1, "0", False, 3, 123456789, assume_yes=True)
    cp.compute(p, out2, "KR", "all", 2.0, 10, 8, "0.95,0.98", 2, 0.1, "0", False, 3, 123456789, assume_yes=True)
    for f in ("result_unfiltered.tsv", "result_filtered.tsv"):
        assert open(out1 + "/" + f).read() == open(out2 + "/" + f).read()
