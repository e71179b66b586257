"""ORACLE (test infrastructure, never on the product path) — the ``compute`` and ``score``
drivers restated on top of the stage functions of this package.

Follows reference src/stripenn/stripenn.py:64-163 (compute: norm resolution :80-92,
chromosome filter :96, -k handling :104-116, stage order :120-159) and
src/stripenn/score.py:7-61 (getScore).  No files are written here; the product's host layer
owns the TSV/log layout.
"""
import numpy as np

from . import background as bgm
from . import pipeline as pl
from . import scoring as sc
from .matrix import OracleMatrix

OUT_COLS = ("chr", "pos1", "pos2", "chr2", "pos3", "pos4", "length", "width", "Mean",
            "maxpixel", "pvalue", "Stripiness")


def resolve_norm(cm, norm, mode="compute"):
    """stripenn.py:82-92 / score.py:27-35 ('weight' -> True only in compute)."""
    cols = cm.bin_columns()
    if norm == "None":
        return False
    if mode == "compute" and norm == "weight":
        return True
    if norm not in cols:
        return False
    return norm


def compute_chromosomes(cm, chrom="all"):
    """stripenn.py:94-116 -> (all_names, all_sizes, names, sizes)."""
    names = list(cm.chromnames)
    keep = [i for i in range(len(names))
            if 'JH5' not in names[i] and 'GL4' not in names[i] and 'RANDOM' not in names[i]
            and names[i] not in ("M", "chrM", "Y", "chrY")]
    all_names = [names[i] for i in keep]
    all_sizes = np.asarray(cm.chromsizes)[keep]
    sel = chrom.split(",")
    if sel[0] == "all":
        return all_names, all_sizes, all_names, all_sizes
    idx = [all_names.index(c) for c in sel if c in all_names]
    return all_names, all_sizes, sel, all_sizes[idx]


def score_chromosomes(cm):
    """score.py:37-46."""
    names = list(cm.chromnames)
    ok = [i for i in range(len(names))
          if names[i] != "Y" and "JH" not in names[i] and "RANDOM" not in names[i]]
    nm = [names[i] for i in ok]
    sz = np.asarray(cm.chromsizes)[ok]
    big = np.nonzero(sz > 1000000)[0]
    return [nm[i] for i in big], sz[big]


def compute(cm, norm="None", chrom="all", canny=2.0, minL=10, maxW=8,
            maxpixel="0.95,0.96,0.97,0.98,0.99", pvalue=0.1, mask="0", bfilter=3,
            seed=123456789, gray_variant=0, return_stages=False, frame_filter=None, slow=False):
    qs = list(map(float, maxpixel.split(",")))
    bal = resolve_norm(cm, norm)
    all_names, all_sizes, names, sizes = compute_chromosomes(cm, chrom)
    mat = OracleMatrix(cm, bal)
    MP = pl.quantiles_slow(mat, names, qs) if slow else pl.quantiles(mat, names, qs)
    EV = bgm.expected_by_distance(mat, names)
    bg = bgm.null_distribution(mat, all_names, all_sizes, seed)
    per_q = []
    for qi, q in enumerate(qs):
        t = pl.extract_candidates(mat, names, MP, qi, q, minL, maxW, canny, bfilter,
                                  gray_variant, frame_filter)
        t["pvalue"] = sc.pvalues(mat, bg, t)
        per_q.append(t)
    allq = pl.concat_tables(per_q)
    final = pl.remove_redundant(allq, "pvalue")
    s, _, _ = sc.stripiness(mat, final, EV, mask)
    final["Stripiness"] = s
    unf = {k: list(final[k]) for k in OUT_COLS}
    pv = np.array(unf["pvalue"], dtype=np.float64)
    st = np.array(unf["Stripiness"], dtype=np.float64)
    sel = np.nonzero(pv < pvalue)[0]
    # DataFrame.sort_values(ascending=False) is quicksort on the negated key with NaN last
    order = sel[np.argsort(-st[sel], kind="quicksort")]
    fil = {k: [unf[k][i] for i in order] for k in OUT_COLS}
    if return_stages:
        return unf, fil, dict(MP=MP, EV=EV, bg=bg, per_q=per_q)
    return unf, fil


def score(cm, table, norm="None", seed=123456789, mask="0"):
    """getScore: table has chr,pos1..pos4 (+ anything).  Returns the six appended columns."""
    bal = resolve_norm(cm, norm, mode="score")
    names, sizes = score_chromosomes(cm)
    mat = OracleMatrix(cm, bal)
    EV = bgm.expected_by_distance(mat, names)
    bg = bgm.null_distribution(mat, names, sizes, seed)
    pv = sc.pvalues(mat, bg, table)
    mean, total = sc.observed_mean(mat, table)
    s, moe, toe = sc.stripiness(mat, table, EV, mask)
    return {"pvalue_added": pv, "Stripiness_added": s, "O_Mean_added": mean,
            "O_Sum_added": total, "O/E_Mean_added": moe, "O/E_Total_added": toe}


def seeimage(cm, position, maxpixel="0.95,0.96,0.97,0.98,0.99", norm="KR", slow=False):
    """seeimage.py:32-97 restated: -> [(file-name suffix "_<position>_<q>qt.png", green(=blue)
    channel float64)] per maxpixel value; the red channel is 1.0 everywhere.  Norm resolution
    :41-50 (no 'weight' -> True special case), chromosome filter :59-62 (> 500 kb, nothing else),
    quantiles of the position's chromosome only :70-74, window = fetch(position, position) :75,
    saturation :79-88 (same arithmetic as image.saturate)."""
    import sys
    from . import image as im
    qs = list(map(float, maxpixel.split(",")))
    bal = resolve_norm(cm, norm, mode="seeimage")
    chrom = position.split(":")[0]
    if chrom not in cm.chromnames:
        sys.exit("Invalid chromosome name.")
    if not np.any(np.asarray(cm.chromsizes) > 500000):
        sys.exit("Exit: All chromosomes are shorter than 50kb.")
    mat = OracleMatrix(cm, bal)
    MP = pl.quantiles_slow(mat, [chrom], qs) if slow else pl.quantiles(mat, [chrom], qs)
    if ":" in position:
        s, e = position.rsplit(":", 1)[1].split("-")
        start, end = int(s.replace(",", "")), int(e.replace(",", ""))
    else:
        start, end = 0, mat.chromsize(chrom)
    A = mat.fetch_bp(chrom, start, end, start, end)
    return [("_" + position + "_" + str(q) + "qt.png", im.saturate(A, MP[chrom][i]))
            for i, q in enumerate(qs)]
