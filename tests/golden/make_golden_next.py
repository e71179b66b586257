"""Golden fixtures for the SURVEY.md §8f "next" rows, from the UNMODIFIED reference run through
oracle/refharness (build container only, needs /root/reference):

  tiny2_mask.npz   getStripe.scoringstripes(df, EV, mask) on 'tiny2' (--norm KR) for 160 synthetic
                   stripes and two masks.  The reference's masking() raises IndexError whenever a
                   masked bin equals int(pos1/r), int(pos2/r) or int(pos4/r) of a stripe (the whole
                   run dies); those rows are found with the oracle, CONFIRMED to raise in the
                   reference one by one, and recorded in `<mask>_raises`; the other rows carry the
                   reference's values.
  slow_quant.npz   getStripe.getQuantile_slow (-s) on 'tiny2' (None and KR) and on a
                   six-chromosome map (SLOW3; rebuilt by the tests with synth.make_map(SLOW3, 5000,
                   seed=11, balanced=True)) whose sizes make the reference's genome-wide row/column
                   arithmetic start inside the PREVIOUS chromosome (getStripe.py:139-146).

    python tests/golden/make_golden_next.py
"""
import os
import sys
import tempfile

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

from refharness import run_reference as rr  # noqa: E402
from snn_oracle import background as bgm, scoring as sc  # noqa: E402
from snn_oracle.matrix import OracleMatrix  # noqa: E402
from stripenn_b200 import synth  # noqa: E402

# sizes with small fractional bins: sum(ceil) - ceil(sum) = 3 at s4/s5, so their windows start two
# bins inside the previous chromosome; s4 has 6,100 bins > 5,000, so w = floor(25e6/L) < L and the
# reference's per-block "last row lost" slices matter
SLOW3 = [("s1", 700 * 5000 + 100), ("s2", 600 * 5000 + 200), ("s3", 650 * 5000 + 300), ("s3b", 520 * 5000 + 400),
         ("s4", 6100 * 5000 - 77), ("s5", 800 * 5000 + 50)]


def main():
    tmp = tempfile.mkdtemp(prefix="snn_golden_next_")
    cm2 = synth.make_config("tiny2")
    p2 = os.path.join(tmp, "tiny2.npz")
    cm2.save_npz(p2)
    names = list(cm2.chroms.keys())

    # ---- --mask
    tab = synth.make_stripe_table(cm2, 160)
    df = pd.DataFrame({"chr": np.asarray(tab["chr"]).astype(str), "pos1": tab["pos1"], "pos2": tab["pos2"],
                       "chr2": np.asarray(tab["chr"]).astype(str), "pos3": tab["pos3"], "pos4": tab["pos4"]})
    # masks: a fixed one, three rows inside the first chrB stripe, two left-flank columns and the
    # bin right of the body (the right part's shifted index) of the first chrA stripe
    r = cm2.binsize
    mat = OracleMatrix(cm2, "KR")
    base = np.array(sc.stripiness(mat, {c: df[c].tolist() for c in df.columns},
                                  bgm.expected_by_distance(mat, names), "0")[0], dtype=float)
    fin = np.isfinite(base)
    iB = int(np.nonzero((df["chr"].to_numpy() == names[-1]) & fin)[0][0])
    iA = int(np.nonzero((df["chr"].to_numpy() == names[0]) & fin)[0][0])
    yB = int(df["pos3"].iloc[iB] // r)
    xA, xeA = int(df["pos1"].iloc[iA] // r), int(df["pos2"].iloc[iA] // r)
    masks = [names[0] + ":2000000-2060000",
             "%s:%d-%d" % (names[-1], (yB + 3) * r + 7, (yB + 5) * r + 11),
             "%s:%d-%d" % (names[0], (xA - 3) * r, (xA - 2) * r + 1),
             "%s:%d-%d" % (names[0], (xeA + 1) * r + 5, (xeA + 2) * r)]
    obj, lib = rr.make_object(p2, norm="KR")
    EV = obj.mpmean()
    out = {"chr": df["chr"].to_numpy().astype(str), "pos": df[["pos1", "pos2", "pos3", "pos4"]].to_numpy(np.int64),
           "masks": np.array(masks)}
    for c in obj.chromnames:
        out["EV_" + c] = np.asarray(EV[c], dtype=np.float64)
    for k, mask in enumerate(masks):
        raises = np.zeros(len(df), dtype=bool)
        for i in range(len(df)):
            try:
                sc.stripiness(mat, {c: [df[c].iloc[i]] for c in df.columns}, EV, mask)
            except IndexError:
                raises[i] = True
        # confirm every oracle-predicted IndexError in the reference, row by row
        for i in np.nonzero(raises)[0]:
            try:
                obj.scoringstripes(df.iloc[[i]].reset_index(drop=True), EV, mask)
            except IndexError:
                continue
            raise AssertionError("reference did not raise on row %d with mask %s" % (i, mask))
        ok = np.nonzero(~raises)[0]
        g, mm, tt = obj.scoringstripes(df.iloc[ok].reset_index(drop=True), EV, mask)
        full = np.full((3, len(df)), np.nan)
        full[0, ok], full[1, ok], full[2, ok] = g, mm, tt
        g0, _, _ = obj.scoringstripes(df.iloc[ok].reset_index(drop=True), EV, "0")
        changed = int(np.sum(~np.isclose(np.array(g, float), np.array(g0, float), rtol=1e-12, equal_nan=True)))
        out["m%d_raises" % k] = raises
        out["m%d_values" % k] = full
        print("mask", mask, "rows", len(df), "raise", int(raises.sum()), "changed by the mask", changed)
    np.savez_compressed(os.path.join(HERE, "tiny2_mask.npz"), **out)

    # ---- -s
    out = {"qs": np.array([0.95, 0.96, 0.97, 0.98, 0.99])}
    qs = out["qs"].tolist()
    for norm in ("None", "KR"):
        obj, lib = rr.make_object(p2, norm=norm)
        MP = obj.getQuantile_slow(lib, obj.chromnames, qs)
        MPf = obj.getQuantile_original(lib, obj.chromnames, qs)
        for c in obj.chromnames:
            out["tiny2_%s_%s" % (norm, c)] = np.asarray(MP[c], dtype=np.float64)
            out["tiny2_%s_%s_fast" % (norm, c)] = np.asarray(MPf[c], dtype=np.float64)
    cm3 = synth.make_map(SLOW3, 5000, seed=11, balanced=True)
    p3 = os.path.join(tmp, "slow3.npz")
    cm3.save_npz(p3)
    for norm in ("None", "KR"):
        obj, lib = rr.make_object(p3, norm=norm)
        MP = obj.getQuantile_slow(lib, obj.chromnames, qs)
        for c in obj.chromnames:
            out["slow3_%s_%s" % (norm, c)] = np.asarray(MP[c], dtype=np.float64)
        # -k with a subset: the genome-wide offsets still come from ALL chromosomes of the file
        obj, lib = rr.make_object(p3, norm=norm, chrom="s5")
        MP = obj.getQuantile_slow(lib, obj.chromnames, qs)
        out["slow3_%s_only_s5" % norm] = np.asarray(MP["s5"], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "slow_quant.npz"), **out)
    print("slow_quant", {k: v.tolist() for k, v in out.items() if k.startswith("slow3")})


if __name__ == "__main__":
    main()
