"""GPU (-m gpu): the fast scoring kernel (csrc/score_fast.cu) against the kernels it replaces.

`k_score_fast` takes the REGULAR stripes (pos2 / pos4 on bin boundaries, complete flanks, window
inside the band); `k_score_warp` / `k_score` keep the rest and are pinned to the oracle and the
reference goldens by the other GPU tests.  Here the same stripes go through both families
(`SNN_SCORE_FAST=0` switches the fast kernel off, `=1` forces it) and must agree: p-value, Stripiness, O/E mean and
O/E total bit for bit (same operations in the same order); observed mean / sum to 1e-13 (the fast
kernel adds row sums, the warp kernel accumulates per lane across rows).  Routing is checked through
`snn_score_routing`, so a silent fall-back to the old kernels fails the test."""
import os

import numpy as np
import pytest

from stripenn_b200 import compute as cp
from stripenn_b200 import getstripe as gs
from stripenn_b200 import synth

pytestmark = pytest.mark.gpu
KEYS_EXACT = ("pvalue", "stripiness", "oe_mean", "oe_total")
KEYS_CLOSE = ("o_mean", "o_sum")


def _unload_all(dev):
    for cid in list(dev.loaded):
        dev.unload_chrom(cid)


@pytest.fixture()
def dev(device):
    _unload_all(device)
    yield device
    os.environ.pop("SNN_SCORE_FAST", None)
    _unload_all(device)


def _setup(dev, cm, norm):
    bal = cp.resolve_norm(cm, norm, "compute", False)
    all_names, all_sizes, names, sizes = cp.compute_chromosomes(cm, ["all"])
    names = [c for c in names if c in cm.chroms]
    lib = gs.DeviceMatrix(cm, bal, dev)
    obj = gs.getStripe(lib, cm.binsize, 10, 8, 2.0, all_names, names, all_sizes,
                       [cm.chroms[c].size_bp for c in names], 1, 3, 123456789, 0)
    obj.all_chromnames = names
    obj.all_chromsizes = np.array([cm.chroms[c].size_bp for c in names])
    bg = obj.nulldist()
    obj._install_background(bg)
    ex = obj.mpmean()
    ids = np.array([cm.chromnames.index(c) for c in names], dtype=np.int32)
    exs = np.stack([np.asarray(ex[c], dtype=np.float64) for c in names])
    return names, ids, exs


def _both(dev, ids, index, pos, exs, want_pvalue=True, expected=True, stale=None):
    os.environ["SNN_SCORE_FAST"] = "0"
    dev.score_routing()
    old = dev.score_multi(ids, index, pos, exs if expected else None, want_pvalue, stale=stale)
    r_old = dev.score_routing()
    os.environ["SNN_SCORE_FAST"] = "1"
    new = dev.score_multi(ids, index, pos, exs if expected else None, want_pvalue, stale=stale)
    r_new = dev.score_routing()
    assert r_old[0] == 0
    return old, new, r_new


def _compare(old, new, keys_exact=KEYS_EXACT, keys_close=KEYS_CLOSE):
    for k in keys_exact:
        a, b = old[k], new[k]
        same = (a == b) | (np.isnan(a) & np.isnan(b))
        assert same.all(), "%s: %d of %d differ, first %s" % (k, (~same).sum(), len(a), np.nonzero(~same)[0][:5])
    for k in keys_close:
        np.testing.assert_allclose(new[k], old[k], rtol=1e-13, atol=0, equal_nan=True, err_msg=k)


def _table(cm, names, n, seed=0):
    tab = synth.make_stripe_table(cm, n, chroms=names, seed=synth.DEFAULT_SEED + seed)
    pos = np.stack([tab["pos1"], tab["pos2"], tab["pos3"], tab["pos4"]], axis=1).astype(np.int64)
    lut = {c: k for k, c in enumerate(names)}
    index = np.array([lut[c] for c in np.asarray(tab["chr"]).astype(str)], dtype=np.int32)
    return pos, index


def _kr_map():
    chroms, binsize, _ = synth.config_chroms("cfg3")
    keep = ("chr20", "chr21", "chr22")
    return synth.make_map([c for c in chroms if c[0] in keep], binsize, balanced=True)


def test_fast_equals_warp_kernel_kr(dev):
    """20,000 config-5-shaped stripes on three KR-balanced config-3 chromosomes (2 % NaN-weight bins:
    dead columns and deleted rows are common), every height class."""
    cm = _kr_map()
    names, ids, exs = _setup(dev, cm, "KR")
    pos, index = _table(cm, names, 20000)
    old, new, routed = _both(dev, ids, index, pos, exs)
    assert routed[0] == len(pos), routed            # every stripe of the bench table is regular
    _compare(old, new)
    assert np.isfinite(new["pvalue"]).all()
    h = -(-pos[:, 3] // cm.binsize) - pos[:, 2] // cm.binsize
    assert (h <= 32).any() and ((h > 32) & (h <= 64)).any() and ((h > 64) & (h <= 128)).any() and (h > 128).any()


def test_fast_pvalue_only_and_stripiness_only(dev):
    """The two single-purpose calls `compute` makes (getStripe.pvalue, then scoringstripes)."""
    cm = _kr_map()
    names, ids, exs = _setup(dev, cm, "KR")
    pos, index = _table(cm, names, 4000, seed=1)
    old, new, routed = _both(dev, ids, index, pos, exs, want_pvalue=True, expected=False)
    assert routed[0] == len(pos)
    _compare(old, new, keys_exact=("pvalue",), keys_close=())
    old, new, routed = _both(dev, ids, index, pos, exs, want_pvalue=False, expected=True)
    assert routed[0] == len(pos)
    _compare(old, new, keys_exact=("stripiness", "oe_mean", "oe_total"))


def test_fast_dense_nan_bins(dev):
    """A quarter of the bins masked: most windows lose columns, many lose every row of a part,
    some have no finite cell at all."""
    cm = _kr_map()
    rng = np.random.default_rng(7)
    for col in ("weight", "KR"):
        for c in cm.weights[col]:
            w = cm.weights[col][c]
            w[rng.random(len(w)) < 0.25] = np.nan
    for c in cm.weights["weight"]:          # a run of masked bins: windows without any finite cell
        cm.weights["weight"][c][3000:3400] = np.nan
        cm.weights["KR"][c][3000:3400] = np.nan
    names, ids, exs = _setup(dev, cm, "KR")
    pos, index = _table(cm, names, 6000, seed=2)
    extra = []
    r = cm.binsize
    for a in range(3050, 3300, 10):       # stripes inside the masked run
        extra.append([a * r + 1, (a + 5) * r, a * r + 1, (a + 40) * r])
    pos = np.concatenate([pos, np.array(extra, dtype=np.int64)])
    index = np.concatenate([index, np.zeros(len(extra), dtype=np.int32)])
    old, new, routed = _both(dev, ids, index, pos, exs)
    assert routed[0] == len(pos)
    _compare(old, new)
    assert np.isnan(new["stripiness"]).any() and np.isfinite(new["stripiness"]).any()


def test_fast_raw_counts_and_irregular_neighbours(dev):
    """--norm None (no weights, no NaN) at 10 kb (bs = 5); irregular stripes in the same call are
    routed to the old kernels and the outputs of both groups land in the right rows."""
    chroms, binsize, _ = synth.config_chroms("cfg2")
    keep = ("18", "19")
    cm = synth.make_map([c for c in chroms if c[0] in keep], binsize, balanced=False)
    names, ids, exs = _setup(dev, cm, "None")
    pos, index = _table(cm, names, 6000, seed=3)
    r = cm.binsize
    odd = np.arange(0, len(pos), 7)
    pos[odd, 1] += 1234                                  # pos2 off the bin boundary: not regular
    near = np.arange(3, len(pos), 11)                    # stripes at the chromosome start: clipped left flank
    for i in near:
        w = (pos[i, 1] - pos[i, 0]) // r + 1
        pos[i] = [2 * r + 1, (2 + w) * r, 2 * r + 1, (2 + 30) * r]
    stale = np.full((len(pos), 2), -1, dtype=np.int32)
    mid = np.arange(5, len(pos), 13)                     # neither end on the diagonal: stale-row rule
    for i in mid:
        pos[i, 2] += 3 * r
        pos[i, 3] += 3 * r
        stale[i] = [i % 2, (i * 7) % 400]
    old, new, routed = _both(dev, ids, index, pos, exs, stale=stale)
    assert routed[0] > 0 and routed[1] + routed[2] > 0 and sum(routed) == len(pos), routed
    _compare(old, new)


def test_fast_one_kb_wide_flanks(dev):
    """1 kb bins: bs = 50, the widest flanks (two 32-bit halves of the dead-column masks)."""
    cm = synth.make_map([("chrT", 9_000_000)], 1000, balanced=True)
    names, ids, exs = _setup(dev, cm, "KR")
    pos, index = _table(cm, names, 3000, seed=4)
    old, new, routed = _both(dev, ids, index, pos, exs)
    assert routed[0] == len(pos), routed
    _compare(old, new)


def test_fast_auto_rule(dev):
    """Unset SNN_SCORE_FAST: a small table does not pay for the full band (warp kernel), a large one
    does; once the bands are resident the small table uses them too."""
    cm = _kr_map()
    names, ids, exs = _setup(dev, cm, "KR")
    small_pos, small_index = _table(cm, names, 300, seed=5)
    big_pos, big_index = _table(cm, names, 30000, seed=6)
    os.environ.pop("SNN_SCORE_FAST", None)
    dev.score_routing()
    a = dev.score_multi(ids, small_index, small_pos, exs, True)
    assert dev.score_routing()[0] == 0
    dev.score_multi(ids, big_index, big_pos, exs, True)
    assert dev.score_routing()[0] == len(big_pos)
    b = dev.score_multi(ids, small_index, small_pos, exs, True)
    assert dev.score_routing()[0] == len(small_pos)
    _compare(a, b)


def test_fast_tree_search_equals_row_search(dev):
    """The 5-ary float32 key tree over the sorted background rows against the float64 halving search
    (SNN_SCORE_TREE=0), and with every lane forced through the tree's repeat path (=2)."""
    cm = _kr_map()
    names, ids, exs = _setup(dev, cm, "KR")
    pos, index = _table(cm, names, 8000, seed=8)
    os.environ["SNN_SCORE_FAST"] = "1"
    try:
        os.environ["SNN_SCORE_TREE"] = "0"
        plain = dev.score_multi(ids, index, pos, exs, True)
        os.environ["SNN_SCORE_TREE"] = "2"
        forced = dev.score_multi(ids, index, pos, exs, True)
        os.environ.pop("SNN_SCORE_TREE")
        tree = dev.score_multi(ids, index, pos, exs, True)
    finally:
        os.environ.pop("SNN_SCORE_TREE", None)
    _compare(plain, tree, keys_exact=KEYS_EXACT + KEYS_CLOSE, keys_close=())
    _compare(plain, forced, keys_exact=KEYS_EXACT + KEYS_CLOSE, keys_close=())
