"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference
(/root/reference/src/stripenn) through oracle/refharness on the synthetic contact maps of
stripenn_b200.synth.  Only runs in the build container (needs /root/reference); the GPU box
uses the committed .npz files.

    python tests/golden/make_golden.py            # ~4 min on one core

Fixtures (all small, compressed):
  tiny_compute.npz     stripenn.compute on 'tiny' (1 chromosome, 2,400 bins, --norm None,
                       5 maxpixel quantiles): both result tables + the stage outputs
                       (quantiles, mpmean, a column subset of the four background tables,
                       per-maxpixel candidate tables with p-values)
  tiny2_compute.npz    same on 'tiny2' (2 chromosomes, --norm KR, NaN-weight bins), 2 quantiles
  tiny2_score.npz      score.getScore on 'tiny2' (--norm KR) for 160 synthetic stripes
  tiny_stages.npz      per-brightness-pass intermediates of StripeSearch for two search windows
                       (gray float32 input of canny, edge map, verticalLine output)
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

from refharness import run_reference as rr  # noqa: E402
from stripenn_b200 import synth  # noqa: E402

BG_COLS = slice(0, None, 25)
SKIP_EXISTING = "--all" not in sys.argv
TABLE_COLS = ("chr", "pos1", "pos2", "chr2", "pos3", "pos4", "length", "width", "Mean",
              "maxpixel", "pvalue", "Stripiness")


def table_arrays(df, prefix, cols=TABLE_COLS):
    out = {}
    for c in cols:
        v = df[c].to_numpy()
        if c in ("chr", "chr2", "maxpixel"):
            v = v.astype(str)
        elif v.dtype == object:      # pandas keeps list-built columns as object
            f = np.array([float(x) for x in v])
            v = f.astype(np.int64) if np.all(f == np.floor(f)) and c not in ("Mean", "pvalue", "Stripiness") else f
        out[prefix + c] = v
    return out


def staged_compute(npz, norm, maxpixel):
    """What stripenn.compute does (stripenn.py:120-152), stage by stage, keeping the
    intermediate products."""
    import pandas as pd
    m = rr.setup()
    gs = m["getStripe"].getStripe
    obj, lib = rr.make_object(npz, norm=norm)
    qs = list(map(float, maxpixel.split(",")))
    MP = obj.getQuantile_original(lib, obj.chromnames, qs)
    EV = obj.mpmean()
    bg = obj.nulldist()
    per_q = []
    res = pd.DataFrame()
    for i, q in enumerate(qs):
        t = obj.extract(MP, i, q, *bg)
        per_q.append(t)
        res = pd.concat([res, t])
    res = obj.RemoveRedundant(res, "pvalue")
    s, _, _ = obj.scoringstripes(res, EV, "0")
    out = {"qs": np.array(qs)}
    for c in obj.chromnames:
        out["MP_" + c] = np.asarray(MP[c], dtype=np.float64)
        out["EV_" + c] = np.asarray(EV[c], dtype=np.float64)
    for name, t in zip(("blu", "bru", "bld", "brd"), bg):
        out["bg_" + name] = np.asarray(t)[:, BG_COLS]
        out["bg_" + name + "_shape"] = np.array(np.asarray(t).shape)
        out["bg_" + name + "_nansum"] = np.nansum(np.asarray(t), axis=1)
        out["bg_" + name + "_nnan"] = np.isnan(np.asarray(t)).sum(axis=1)
    for i, t in enumerate(per_q):
        out.update(table_arrays(t, "q%d_" % i, ("chr", "pos1", "pos2", "pos3", "pos4",
                                                 "Mean", "maxpixel", "pvalue", "num",
                                                 "x", "y", "h", "w")))
    return out


def main():
    tmp = tempfile.mkdtemp(prefix="snn_golden_")
    # ---- tiny (norm None, 5 quantiles)
    cm = synth.make_config("tiny")
    p = os.path.join(tmp, "tiny.npz")
    cm.save_npz(p)
    mp = "0.95,0.96,0.97,0.98,0.99"
    if not (SKIP_EXISTING and os.path.exists(os.path.join(HERE, "tiny_compute.npz"))):
        unf, fil = rr.run_compute(p, maxpixel=mp)
        out = {}
        out.update(table_arrays(unf, "unf_"))
        out.update(table_arrays(fil, "fil_"))
        out.update(staged_compute(p, "None", mp))
        np.savez_compressed(os.path.join(HERE, "tiny_compute.npz"), **out)
        print("tiny_compute", len(unf), len(fil))

    # ---- stage dumps: two windows of 'tiny', maxpixel 0.97
    obj, lib = rr.make_object(p, norm="None")
    MP = obj.getQuantile_original(lib, obj.chromnames, [0.97])
    M = MP["T1"][0]
    import pandas as pd
    st = {"M": np.float64(M)}
    for idx in (0, 5):
        start, end = (0 if idx == 0 else idx * 200 - 100), min((idx + 1) * 200 + 99, 2399)
        # the same window construction search_frame does (getStripe.py:794-821)
        r = 5000
        sa = (np.arange(start, end + 1) * r + 1)
        ea = sa + r - 1
        if ea[-1] >= obj.chromsizes[0]:
            ea[-1] = obj.chromsizes[0]
        loc = "T1:%d-%d" % (sa[0], ea[-1])
        D = obj.unbalLib.fetch(loc, loc)
        D = np.nan_to_num(D)
        nz = np.where(np.sum(D, axis=0) != 0)[0]
        D = D[np.ix_(nz, nz)]
        with rr.StageRecorder() as rec:
            res = obj.StripeSearch(D, idx, start, end, M, 0.97, "T1", len(nz), list(sa[nz]), list(ea[nz]))
        st["f%d_S" % idx] = np.int64(len(nz))
        st["f%d_nz" % idx] = nz.astype(np.int32)
        for b in range(6):
            g = rec.gray[b]
            if b in (0, 4):
                st["f%d_gray%d" % (idx, b)] = g
            st["f%d_graysum%d" % (idx, b)] = np.float64(g.astype(np.float64).sum())
            st["f%d_edges%d" % (idx, b)] = np.packbits(rec.edges[b].astype(bool))
            st["f%d_vert%d" % (idx, b)] = np.packbits(rec.vert[b].astype(bool))
        for c in ("x", "y", "w", "h"):
            st["f%d_res_%s" % (idx, c)] = res[c].to_numpy().astype(np.int64) if len(res) else np.zeros(0, np.int64)
    np.savez_compressed(os.path.join(HERE, "tiny_stages.npz"), **st)
    print("tiny_stages")

    # ---- tiny2 (KR balanced, 2 chromosomes)
    cm2 = synth.make_config("tiny2")
    p2 = os.path.join(tmp, "tiny2.npz")
    cm2.save_npz(p2)
    mp2 = "0.95,0.98"
    unf, fil = rr.run_compute(p2, maxpixel=mp2, norm="KR")
    out = {}
    out.update(table_arrays(unf, "unf_"))
    out.update(table_arrays(fil, "fil_"))
    out.update(staged_compute(p2, "KR", mp2))
    np.savez_compressed(os.path.join(HERE, "tiny2_compute.npz"), **out)
    print("tiny2_compute", len(unf), len(fil))

    # ---- score mode on tiny2
    tab = synth.make_stripe_table(cm2, 160)
    coord = os.path.join(tmp, "coords.tsv")
    with open(coord, "w") as f:
        for i in range(len(tab["pos1"])):
            f.write("%s\t%d\t%d\t%s\t%d\t%d\n" % (tab["chr"][i], tab["pos1"][i], tab["pos2"][i],
                                                   tab["chr"][i], tab["pos3"][i], tab["pos4"][i]))
    res = rr.run_score(p2, coord, os.path.join(tmp, "scored.tsv"), norm="KR")
    out = {}
    for c in res.columns:
        v = res[c].to_numpy()
        out[c.replace("/", "_")] = v.astype(str) if v.dtype == object or c.startswith("chr") else v
    np.savez_compressed(os.path.join(HERE, "tiny2_score.npz"), **out)
    print("tiny2_score", len(res))


if __name__ == "__main__":
    main()
