"""Config-size parity fixture (VERDICT r01 "parity at config size"): the UNMODIFIED reference
(/root/reference/src/stripenn through oracle/refharness) on `mid` — a 12,050-bin KR-balanced
chromosome with NaN-weight bins plus a chromosome shorter than 500 bins (`itera` = 0 in nulldist:
the reference drops it from the sample budget, getStripe.py:249-283) — with the default five
maxpixel quantiles, i.e. `compute()` as stripenn.py:120-152 runs it on BASELINE configs 1/3.

    SNN_REF_PROCS=8 python tests/golden/make_golden_mid.py      # ~6 min on 8 cores

Writes tests/golden/mid_compute.npz: both result tables of stripenn.compute, and from a staged
run the quantiles, expected values, a column subset of the background tables and the
per-maxpixel candidate tables with p-values.
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, HERE)

from refharness import run_reference as rr  # noqa: E402
from stripenn_b200 import synth  # noqa: E402
import make_golden as mg  # noqa: E402


def main():
    tmp = tempfile.mkdtemp(prefix="snn_golden_mid_")
    cm = synth.make_config("mid")
    p = os.path.join(tmp, "mid.npz")
    cm.save_npz(p)
    mp = "0.95,0.96,0.97,0.98,0.99"
    unf, fil = rr.run_compute(p, maxpixel=mp, norm="KR")
    out = {}
    out.update(mg.table_arrays(unf, "unf_"))
    out.update(mg.table_arrays(fil, "fil_"))
    out.update(mg.staged_compute(p, "KR", mp))
    np.savez_compressed(os.path.join(HERE, "mid_compute.npz"), **out)
    print("mid_compute", len(unf), len(fil))


if __name__ == "__main__":
    main()
