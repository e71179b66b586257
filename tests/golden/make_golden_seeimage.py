"""Golden fixture for `stripenn seeimage` from the UNMODIFIED reference
(src/stripenn/seeimage.py:32-97) run through oracle/refharness — the matplotlib.pyplot stub
records the float image handed to ``imshow`` and the file name handed to ``savefig`` (build
container only, needs /root/reference):

  seeimage.npz   for each case k: position, norm, slow flag, maxpixel string, and per maxpixel
                 value i the green channel (== blue channel; red is 1.0 everywhere, checked here)
                 `k<k>_g<i>`, the output path suffix `k<k>_name<i>` and the title.
                 Cases: tiny2 / KR window holding NaN-weight bins; tiny / None 520-bin window
                 (wider than the 448-bin band: far pixels); tiny2 / KR with -s.

    python tests/golden/make_golden_seeimage.py
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

CASES = [  # (map, position, maxpixel, norm, slow)
    ("tiny2", "chrA:2000000-2800000", "0.95,0.98", "KR", False),
    ("tiny", "T1:5000000-7600000", "0.97", "None", False),
    ("tiny2", "chrB:1000000-1600000", "0.96", "KR", True),
    ("tiny2", "chrB:6900001-7499923", "0.99", "bogus", False),     # unknown norm -> None; last (partial) bin
]


def main():
    rr.setup()
    from stripenn import seeimage as see          # reference module, unmodified
    import matplotlib.pyplot as plt               # the recording stub
    tmp = tempfile.mkdtemp(prefix="snn_golden_see_")
    paths = {}
    for name in ("tiny", "tiny2"):
        paths[name] = os.path.join(tmp, name + ".npz")
        synth.make_config(name).save_npz(paths[name])
    out = {"ncases": np.array(len(CASES))}
    for k, (mp_, pos, mpx, norm, slow) in enumerate(CASES):
        del plt.CALLS[:]
        see.seeimage(paths[mp_], pos, mpx, norm, os.path.join(tmp, "heat"), slow, 123456789)
        out["k%d_map" % k] = np.array(mp_)
        out["k%d_position" % k] = np.array(pos)
        out["k%d_maxpixel" % k] = np.array(mpx)
        out["k%d_norm" % k] = np.array(norm)
        out["k%d_slow" % k] = np.array(slow)
        assert len(plt.CALLS) == len(mpx.split(","))
        for i, c in enumerate(plt.CALLS):
            img = c["img"]
            assert img.dtype == np.float64 and np.all(img[:, :, 0] == 1.0)
            assert np.array_equal(img[:, :, 1], img[:, :, 2], equal_nan=True)
            out["k%d_g%d" % (k, i)] = img[:, :, 1]
            out["k%d_name%d" % (k, i)] = np.array(os.path.basename(c["path"]))
            out["k%d_title%d" % (k, i)] = np.array(c["title"])
            print(k, pos, norm, slow, img.shape, "nan px", int(np.isnan(img[:, :, 1]).sum()),
                  "distinct", len(np.unique(img[:, :, 1])), os.path.basename(c["path"]))
    np.savez_compressed(os.path.join(HERE, "seeimage.npz"), **out)
    print("seeimage.npz", os.path.getsize(os.path.join(HERE, "seeimage.npz")), "bytes")


if __name__ == "__main__":
    main()
