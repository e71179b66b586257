"""CPU checks of the two arguments the fast scoring kernel (csrc/score_fast.cu) rests on.

1. The division sequence `f_div` (two Newton corrections of RN(a * RN(1/b))) equals IEEE `/` — the C
   checker tools/check_fdiv.c, run here on a reduced range (the full run, every count pair up to 26,000
   and 2e9 random numerators, is recorded in profiles/r02v_check_fdiv.json).
2. The background-row search: a 5-ary descent over float32 keys rounded toward -inf, against the
   threshold rounded toward +inf, ends at q = #(key < RU(v)); q >= #(entry < v), with equality unless an
   entry lies within one float32 ulp above v; the kernel confirms with ONE float64 entry (row[q-1] < v)
   and repeats the search on the float64 row otherwise.  Restated in numpy (same layout as
   background.cu: k_bg_sort) and compared with np.searchsorted on rows built to hit the ambiguous
   cases: duplicates, thresholds equal to entries that float32 cannot represent, thresholds one ulp
   off an entry, +inf padding, empty rows."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fdiv_identity(tmp_path):
    exe = tmp_path / "fdiv"
    subprocess.run(["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-o", str(exe),
                    os.path.join(ROOT, "tools", "check_fdiv.c"), "-lm"], check=True)
    out = subprocess.run([str(exe), "3000", "4000000"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    assert '"bad_pairs": 0' in out.stdout and '"bad_small_divisor": 0' in out.stdout


# ---- the tree of k_bg_sort and the descent of k_score_fast, in numpy -----------------------------
def _round_down32(x):
    f = x.astype(np.float32)
    up = f.astype(np.float64) > x
    return np.where(up, np.nextafter(f, np.float32(-np.inf)), f).astype(np.float32)


def _round_up32(v):
    with np.errstate(over="ignore"):
        f = np.float32(v)
    if float(f) < v:
        f = np.nextafter(f, np.float32(np.inf))
    return f


def _levels(ncol):
    levels, reach = 1, 5
    while reach < ncol + 1:
        reach *= 5
        levels += 1
    return levels


def _build_tree(row_sorted, ncol):
    """Per level l: nodes q < 5^l, keys at sorted positions (5q + r + 1) * 5^(levels-1-l) - 1."""
    levels = _levels(ncol)
    keys = _round_down32(row_sorted)
    tree = []
    stride = 5 ** (levels - 1)
    for l in range(levels):
        q = np.arange(5 ** l)[:, None]
        posn = (5 * q + np.arange(4)[None, :] + 1) * stride - 1
        node = np.full(posn.shape, np.float32(np.inf), dtype=np.float32)
        ok = posn < ncol
        node[ok] = keys[posn[ok]]
        tree.append(node)
        stride //= 5
    return tree


def _descent(tree, vu):
    q = 0
    for node in tree:
        q = 5 * q + int(np.sum(node[q] < vu))
    return q


def _count_less(row_sorted, ncol, tree, v):
    """What the kernel computes: (#(row < v), whether the float64 repeat path was needed)."""
    q = _descent(tree, _round_up32(v))
    assert q <= ncol
    if q > 0 and not (row_sorted[q - 1] < v):
        return int(np.searchsorted(row_sorted[:ncol], v, side="left")), True
    return q, False


def _rows():
    rng = np.random.default_rng(42)
    yield np.sort(rng.normal(0.0, 3.0, 1000))
    yield np.sort(np.round(rng.normal(0.0, 1.0, 1000), 1))                      # heavy duplicates
    yield np.sort(np.concatenate([np.full(300, 0.1), np.full(300, 0.1 + 2 ** -40), rng.random(400)]))
    yield np.sort(np.concatenate([rng.random(700), np.full(300, np.inf)]))      # +inf padding (NaN entries)
    yield np.full(1000, np.inf)                                                 # empty row
    yield np.sort(rng.normal(0.0, 1e-3, 37))                                    # short row
    yield np.sort(rng.lognormal(0.0, 2.0, 3124))                                # longest five-level row
    yield np.sort(np.float64(np.float32(rng.random(1000))))                     # float32-representable entries


@pytest.mark.parametrize("k", range(8))
def test_tree_descent_counts_like_searchsorted(k):
    row = list(_rows())[k]
    ncol = len(row)
    tree = _build_tree(row, ncol)
    rng = np.random.default_rng(k)
    finite = row[np.isfinite(row)]
    probes = [-np.inf, np.inf, -1e300, 1e300, 0.0]
    if len(finite):
        pick = finite[rng.integers(0, len(finite), 200)]
        probes += list(pick)                                     # thresholds equal to entries
        probes += list(np.nextafter(pick, np.inf)) + list(np.nextafter(pick, -np.inf))
        probes += list(np.float64(np.float32(pick)))             # float32 neighbours of entries
        probes += list(rng.uniform(finite.min() - 1, finite.max() + 1, 300))
    repeats = 0
    for v in probes:
        got, repeated = _count_less(row, ncol, tree, float(v))
        want = int(np.searchsorted(row, v, side="left"))
        assert got == want, (k, v, got, want)
        # the descent alone never undercounts: q >= #(entry < v)
        assert _descent(tree, _round_up32(float(v))) >= want
        repeats += repeated
    if k == 7:          # float32-representable entries and thresholds: keys are exact, no repeat ever
        assert repeats == 0


def test_tree_levels_match_the_library_rule():
    assert _levels(1000) == 5 and _levels(624) == 4 and _levels(625) == 5 and _levels(3124) == 5 and _levels(3125) == 6
