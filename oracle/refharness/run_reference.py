"""Run the UNMODIFIED reference (``/root/reference/src/stripenn``) on a synthetic ``.npz``
contact map, through three stub packages (``oracle/refharness/stubs``: matplotlib, skimage,
cooler).  This is how the golden fixtures under ``tests/golden`` were produced and how the
numpy restatement in ``oracle/snn_oracle`` is pinned (tests/test_oracle_vs_reference.py).

It can only run where ``/root/reference`` exists (the build container) — nothing on the GPU
box imports this module.

RNG convention: the reference's background sampler draws from
``self.prng`` inside joblib tasks; with ``-n >= 2`` every chromosome task receives a pickled
copy of the *unadvanced* ``random.Random(seed)``.  We reproduce that deterministically by
replacing ``getStripe.Parallel`` with a sequential executor that round-trips the
``main_null_calc`` task through cloudpickle before calling it (identical to loky with batch
size 1), and runs every other task in-process.

TEST INFRASTRUCTURE ONLY — never imported by the product path.
"""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
STUBS = os.path.join(HERE, "stubs")
REFERENCE_SRC = os.environ.get("STRIPENN_REFERENCE_SRC", "/root/reference/src")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_SRC, "stripenn", "getStripe.py"))


class _SeqParallel:
    """Drop-in for joblib.Parallel used by getStripe (getStripe.py:223,276,482,527,782,842,1189)."""

    def __init__(self, n_jobs=None, **_kw):
        self.n_jobs = n_jobs

    # tasks that neither draw random numbers nor touch shared state: with SNN_REF_PROCS=N they run on
    # N forked workers (same results as the sequential loop; used for the config-size goldens and
    # by bench.py's reference arm, which is allowed all host cores)
    PURE = ("search_frame", "iterate_idx", "calc", "available_cols")

    def __call__(self, tasks):
        import cloudpickle
        tasks = list(tasks)
        procs = int(os.environ.get("SNN_REF_PROCS", "1") or 1)
        if procs > 1 and len(tasks) > 1 and all(getattr(t[0], "__name__", "") in self.PURE for t in tasks):
            return _fork_map(tasks, procs)
        out = []
        for fn, args, kwargs in tasks:
            if getattr(fn, "__name__", "") == "main_null_calc":
                fn = cloudpickle.loads(cloudpickle.dumps(fn))
            out.append(fn(*args, **kwargs))
        return out


_FORK_TASKS = None


def _fork_call(i):
    fn, args, kwargs = _FORK_TASKS[i]
    return fn(*args, **kwargs)


def _fork_map(tasks, procs):
    """Run closures on forked workers without pickling them: the task list is inherited through
    fork, only indices go in and results come back."""
    global _FORK_TASKS
    import multiprocessing as mp
    _FORK_TASKS = tasks
    try:
        with mp.get_context("fork").Pool(min(procs, len(tasks))) as pool:
            return pool.map(_fork_call, range(len(tasks)), chunksize=1)
    finally:
        _FORK_TASKS = None


_MODS = None


def setup():
    """Put stubs + reference on sys.path, import the reference, patch Parallel/tqdm."""
    global _MODS
    if _MODS is not None:
        return _MODS
    if not reference_available():
        raise RuntimeError("reference sources not found under %s" % REFERENCE_SRC)
    for p in (STUBS, REFERENCE_SRC):
        if p not in sys.path:
            sys.path.insert(0, p)
    import warnings
    warnings.filterwarnings("ignore")
    from stripenn import getStripe as gs  # noqa: E402  (reference module, unmodified)
    from stripenn import stripenn as drv  # noqa: E402
    from stripenn import score as sc  # noqa: E402
    from stripenn import ImageProcessing as ip  # noqa: E402
    gs.Parallel = _SeqParallel
    gs.tqdm = lambda x, *a, **k: x
    _MODS = dict(getStripe=gs, stripenn=drv, score=sc, ImageProcessing=ip)
    return _MODS


DEFAULTS = dict(norm="None", chrom="all", canny=2.0, minL=10, maxW=8,
                maxpixel="0.95,0.96,0.97,0.98,0.99", numcores=2, pvalue=0.1, mask="0",
                slow=False, bfilter=3, seed=123456789)


def run_compute(npz_path: str, outdir: str | None = None, **opts):
    """``stripenn.compute`` end to end (stripenn.py:64-163).  Returns (unfiltered, filtered)
    as pandas DataFrames read back from the TSVs the reference wrote."""
    import pandas as pd
    m = setup()
    o = dict(DEFAULTS)
    o.update(opts)
    tmp = None
    if outdir is None:
        tmp = tempfile.mkdtemp(prefix="snnref_")
        outdir = os.path.join(tmp, "out")
    m["stripenn"].compute(npz_path, outdir, o["norm"], o["chrom"], o["canny"], o["minL"],
                          o["maxW"], o["maxpixel"], o["numcores"], o["pvalue"], o["mask"],
                          o["slow"], o["bfilter"], o["seed"])
    od = outdir if outdir.endswith("/") else outdir + "/"
    unf = pd.read_csv(od + "result_unfiltered.tsv", sep="\t", dtype={"chr": str, "chr2": str})
    fil = pd.read_csv(od + "result_filtered.tsv", sep="\t", dtype={"chr": str, "chr2": str})
    return unf, fil


def run_score(npz_path: str, coord_path: str, out_path: str, norm="None", numcores=2,
              seed=123456789, mask="0"):
    """``score.getScore`` end to end (score.py:7-61)."""
    import pandas as pd
    m = setup()
    m["score"].getScore(npz_path, coord_path, norm, numcores, seed, out_path, mask)
    return pd.read_csv(out_path, sep="\t", dtype={"chr": str, "chr2": str})


def make_object(npz_path: str, norm="None", chrom="all", canny=2.0, minL=10, maxW=8,
                numcores=2, bfilter=3, seed=123456789):
    """Build the reference ``getStripe`` object exactly as ``compute`` does
    (stripenn.py:80-120) so stages can be called one by one."""
    m = setup()
    import cooler
    lib = cooler.Cooler(npz_path)
    cols = lib.bins().columns
    if norm == "None":
        bal = False
    elif norm == "weight":
        bal = True
    elif norm not in cols:
        bal = False
    else:
        bal = norm
    names = lib.chromnames
    sizes = lib.chromsizes
    keep = [i for i in range(len(names)) if 'JH5' not in names[i] and 'GL4' not in names[i]
            and 'RANDOM' not in names[i] and names[i] not in ('M', 'chrM', 'Y', 'chrY')]
    all_names = [names[i] for i in keep]
    all_sizes = sizes[keep]
    chromnames, chromsizes = all_names, all_sizes
    chroms = chrom.split(',')
    if chroms[0] != 'all':
        idx = [all_names.index(c) for c in chroms if c in all_names]
        chromnames = chroms
        chromsizes = all_sizes[idx]
    sel = lib.matrix(balance=bal)
    obj = m["getStripe"].getStripe(sel, lib.binsize, minL, maxW, canny, all_names, chromnames,
                                   all_sizes, chromsizes, numcores, bfilter, seed)
    return obj, lib


class StageRecorder:
    """Records the per-pass intermediates of StripeSearch by wrapping the module-level
    callables it uses (no reference source is modified): gray (input of feature.canny),
    edges (its output), vert (output of ImageProcessing.verticalLine)."""

    def __init__(self):
        self.gray, self.edges, self.vert = [], [], []

    def __enter__(self):
        m = setup()
        gs = m["getStripe"]
        self._gs = gs
        self._canny = gs.feature.canny
        self._vl = gs.ImageProcessing.verticalLine
        rec = self

        def canny(img, *a, **k):
            out = rec._canny(img, *a, **k)
            rec.gray.append(np.array(img, copy=True))
            rec.edges.append(np.array(out, copy=True))
            return out

        def vline(mm, *a, **k):
            out = rec._vl(mm, *a, **k)
            rec.vert.append(np.array(out, copy=True))
            return out

        class _F:  # shadow `feature` inside getStripe's namespace only
            pass
        f = _F()
        f.canny = canny
        self._feature = gs.feature
        gs.feature = f

        class _IP:
            pass
        ip = _IP()
        for k in dir(self._gs.ImageProcessing):
            if not k.startswith("__"):
                setattr(ip, k, getattr(self._gs.ImageProcessing, k))
        ip.verticalLine = vline
        self._ip = gs.ImageProcessing
        gs.ImageProcessing = ip
        return self

    def __exit__(self, *exc):
        self._gs.feature = self._feature
        self._gs.ImageProcessing = self._ip
        return False
