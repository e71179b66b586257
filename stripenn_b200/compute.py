"""``stripenn compute`` driver — mirrors reference src/stripenn/stripenn.py:64-163 stage by stage
(norm resolution :80-92, chromosome filter :96, -k handling :104-116, stage order :120-152,
output layout :149-159) on top of the B200 :class:`getStripe`.
"""
from __future__ import annotations

import os
import shutil
import sys
import time
import warnings

import numpy as np
import pandas as pd

from . import getstripe as gs
from .contactmap import ContactMap, open_contact_map
from .engine import Device


def _empty_directory(path):
    """Remove everything inside `path`, reporting (not raising) what cannot be removed."""
    with os.scandir(path) as entries:
        for entry in entries:
            try:
                if entry.is_dir(follow_symlinks=False):
                    shutil.rmtree(entry.path)
                else:
                    os.unlink(entry.path)
            except Exception as e:
                print('Failed to delete %s with the reason: %s' % (entry.path, e))


def makeOutDir(outdir, assume_yes=False):
    """Output directory handling of stripenn.py:12-42: a new directory is created; an existing one is
    emptied after the user confirms with Y/y (``assume_yes`` answers for them), anything else ends
    the run.  The prompt and messages are the reference's."""
    outdir = outdir if outdir.endswith('/') else outdir + '/'
    if not os.path.exists(outdir):
        os.makedirs(outdir, exist_ok=True)
        return
    answer = 'Y'
    if not assume_yes:
        print('\n%s exists. Do you want to remove all files and save new results in this folder? [Y/n]' % (outdir))
        answer = input()
    if answer in ('Y', 'y'):
        print('All directories and files in %s will be deleted.' % (outdir))
        _empty_directory(outdir)
        return
    print('Input another output directory. Exit.' if answer in ('n', 'N') else 'Type Y or n.\nExit.')
    sys.exit()


def addlog(cool, out, norm, chrom, canny, minL, maxW, maxpixel, numcores, pvalue, mask, bfilter):
    """stripenn.py:45-61 — same keys, same order."""
    if out[-1] != '/':
        out += '/'
    with open(out + "stripenn.log", 'w') as f:
        f.write('cool: ' + str(cool) + '\n')
        f.write('out: ' + out + '\n')
        f.write('norm: ' + str(norm) + '\n')
        f.write('chrom: ' + str(chrom) + '\n')
        f.write('canny: ' + str(canny) + '\n')
        f.write('minL: ' + str(minL) + '\n')
        f.write('maxW: ' + str(maxW) + '\n')
        f.write('maxpixel: ' + str(maxpixel) + '\n')
        f.write('num_cores: ' + str(numcores) + '\n')
        f.write('pvalue: ' + str(pvalue) + '\n')
        f.write('mask: ' + str(mask) + '\n')
        f.write('blur filter: ' + str(bfilter) + '\n')


def resolve_norm(cm: ContactMap, norm, mode="compute", verbose=True):
    """stripenn.py:82-92 / score.py:27-35."""
    possible = cm.bin_columns()
    if norm == 'None':
        return False
    if mode == "compute" and norm == 'weight':
        return True
    if norm not in possible:
        if verbose:
            print('Possible normalization methods are:')
            print('None')
            for n in range(3, len(possible)):
                print(possible[n])
            print("Invalid normalization method. Normalization method is forced to None")
        return False
    return norm


def compute_chromosomes(cm: ContactMap, chroms):
    """stripenn.py:94-116 -> (all_chromnames, all_chromsizes, chromnames, chromsizes)."""
    names = list(cm.chromnames)
    sizes = np.asarray(cm.chromsizes)
    keep = [i for i in range(len(names))
            if 'JH5' not in names[i] and 'GL4' not in names[i] and 'RANDOM' not in names[i]
            and names[i] != 'M' and names[i] != "chrM" and names[i] != "Y" and names[i] != "chrY"]
    all_names = [names[i] for i in keep]
    all_sizes = sizes[keep]
    if len(all_names) == 0:
        sys.exit("Exit: All chromosomes are shorter than 50kb.")
    chromnames, chromsizes = all_names, all_sizes
    if chroms[0] != 'all':
        idx, warnflag = [], False
        for item in chroms:
            if item in all_names:
                idx.append(all_names.index(item))
            else:
                warnings.warn('\nThere is no chromosomes called ' + str(item) +
                              ' in the provided .cool file or it is shorter than 50kb.')
                warnflag = True
        if warnflag:
            warnings.warn('\nThe possible chromosomes are: ' + ', '.join(all_names))
        chromnames = chroms
        chromsizes = all_sizes[idx]
    return all_names, all_sizes, chromnames, chromsizes


class StageClock:
    """Wall-clock laps of the host pipeline (the reference prints minutes per stage; we keep ms)."""

    def __init__(self):
        self.t = time.perf_counter()
        self.ms = {}

    def lap(self, name):
        now = time.perf_counter()
        self.ms[name] = self.ms.get(name, 0.0) + (now - self.t) * 1e3
        self.t = now


# communicators of the sharded pipeline (re-exported: tests and bench.py use them from here)
from .comm import (HybridComm, LocalComm, ThreadComm, TorchComm, background_layout,  # noqa: E402,F401
                   run_threads)


def lpt_partition(names, sizes, world):
    """Longest-processing-time-first assignment of chromosomes to ranks (work ~ bins).
    Deterministic: ties keep file order.  Returns {chrom: rank}."""
    order = sorted(range(len(names)), key=lambda i: (-int(sizes[i]), i))
    load = [0] * world
    owner = {}
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        owner[names[i]] = r
        load[r] += int(sizes[i])
    return owner


def plan_units(names, sizes_bp, binsize, world, split=True, selected=None, tpp=1):
    """Work units of the sharded pipeline and their owners.  `world` virtual ranks = `world // tpp`
    processes (one per GPU) x `tpp` host threads; virtual rank = process * tpp + thread.

    Units are whole chromosomes, except that a chromosome holding more than 1.25 x a virtual rank's
    fair share of the SEARCHED bins is cut into window ranges (getStripe.py:841-842: windows of one chromosome
    are independent tasks) — a single-chromosome run (BASELINE configs 1 and 4, or `-k chr1` on a
    whole-genome file) then still uses every GPU.  Assignment is longest-processing-time-first
    (work ~ windows) over the GPUs, then over the threads of each GPU: what has to balance is the
    work per GPU — with 16 virtual ranks as bins, chr1 alone exceeded a bin's fair share and the
    GPU holding chr1 + chr2 carried 1.3 x the mean (profiles/bench_r02n_n8*.json).
    `selected` = the -k chromosomes (default: all): only they are searched and only they count for
    the fair share; the other chromosomes of the file still need an owner for the background stage
    (getStripe.py:242,276 samples every chromosome) and are dealt out afterwards as units without
    windows (search = False).
    Returns a list of dicts {chrom, piece, npieces, lo, hi, rank, search}; piece 0's rank is the
    chromosome's PRIMARY (it runs the chromosome-wide stages and the host steps)."""
    import math
    tpp = max(1, min(int(tpp), world))
    nproc = max(1, world // tpp)
    sel = set(names if selected is None else selected)
    nb = [int(math.ceil(int(s) / binsize)) for s in sizes_bp]
    total = max(sum(nb[i] for i, c in enumerate(names) if c in sel), 1)
    fair = total / max(world, 1)          # per virtual rank: a lone chromosome also feeds every thread of a GPU
    units, rest = [], []
    for i, c in enumerate(names):
        if c not in sel:
            rest.append(dict(chrom=c, piece=0, npieces=1, lo=0, hi=0, order=i, rank=0, search=False))
            continue
        nfr = max(1, int(math.ceil(nb[i] / 200)))
        k = 1
        if split and world > 1 and nb[i] > 1.25 * fair:
            k = max(1, min(world, nfr, int(round(nb[i] / fair))))
        for p in range(k):
            lo, hi = (nfr * p) // k, (nfr * (p + 1)) // k
            units.append(dict(chrom=c, piece=p, npieces=k, lo=lo, hi=hi, order=i, rank=0, search=True))
    pload = [0] * nproc                                  # windows per GPU
    tload = [[0] * tpp for _ in range(nproc)]            # windows per host thread of a GPU
    for u in sorted(units, key=lambda u: (-(u["hi"] - u["lo"]), u["order"], u["piece"])):
        p = min(range(nproc), key=lambda k: (pload[k], k))
        t = min(range(tpp), key=lambda k: (tload[p][k], k))
        u["rank"] = p * tpp + t
        pload[p] += u["hi"] - u["lo"]
        tload[p][t] += u["hi"] - u["lo"]
    # background-only chromosomes: their cost (row flags + a share of 1000 samples) is small next to
    # a search; spread them by size on counters of their own
    bload = [0] * nproc
    for u in sorted(rest, key=lambda u: (-nb[u["order"]], u["order"])):
        p = min(range(nproc), key=lambda k: (bload[k], pload[k], k))
        t = min(range(tpp), key=lambda k: (tload[p][k], k))
        u["rank"] = p * tpp + t
        bload[p] += nb[u["order"]]
    return sorted(units + rest, key=lambda u: (u["order"], u["piece"]))


# fixed-dtype records the ranks exchange (never pickled objects)
FINAL_DTYPE = np.dtype([("q", "<i8"), ("ck", "<i8"), ("pos1", "<i8"), ("pos2", "<i8"), ("pos3", "<i8"),
                        ("pos4", "<i8"), ("total", "<f8"), ("h", "<i8"), ("w", "<i8"), ("pvalue", "<f8"),
                        ("s", "<f8")])
PIECE_DTYPE = np.dtype([("ck", "<i4"), ("piece", "<i4"), ("rec", gs.CANDIDATE_DTYPE)])


def make_object(cm, bal, dev, all_names, all_sizes, chromnames, chromsizes, minL, maxW, canny, numcores,
                bfilter, seed, gray_variant):
    lib = gs.DeviceMatrix(cm, bal, dev)
    return gs.getStripe(lib, cm.binsize, minL, maxW, canny, all_names, chromnames, all_sizes, chromsizes,
                        numcores, bfilter, seed, gray_variant)


class DeviceQueue:
    """One worker thread that issues every device call of a pipeline, in submission order.  The
    calls block inside the C library (GIL released) while the GPU works, so the submitting
    thread is free to do the host-side steps (sampling plan, RemoveRedundant, table assembly)
    of earlier results at the same time.  A context is only ever touched by this thread while
    the queue is open — the C ABI's one-thread-at-a-time rule holds."""

    def __init__(self, trace=False):
        from concurrent.futures import ThreadPoolExecutor
        self._ex = ThreadPoolExecutor(max_workers=1, thread_name_prefix="snn-device")
        self.trace = [] if trace else None       # (call, submitted, started, finished) in seconds

    def submit(self, fn, *args, **kw):
        if self.trace is None:
            return self._ex.submit(fn, *args, **kw)
        t_sub = time.perf_counter()
        name = getattr(fn, "__name__", "call")

        def timed():
            t0 = time.perf_counter()
            try:
                return fn(*args, **kw)
            finally:
                self.trace.append((name, t_sub, t0, time.perf_counter()))
        return self._ex.submit(timed)

    def close(self):
        self._ex.shutdown(wait=True)


def filtered_sorted(table, pvals, stripiness, pcut):
    """``table[table.pvalue < pcut].sort_values(by=['Stripiness'], ascending=False)``
    (stripenn.py:153-154) as ONE take instead of a boolean take plus a sorted take: pandas' nargsort
    restated on the indices of the filtered rows (quicksort on the reversed non-NaN keys, result
    reversed, NaN rows last) — same rows, same tie order, same index labels
    (tests/test_host_logic.py compares with the pandas expression)."""
    idx = np.flatnonzero(np.asarray(pvals, dtype=np.float64) < pcut)
    key = np.asarray(stripiness, dtype=np.float64)[idx]
    nan = np.isnan(key)
    good = idx[~nan][::-1]
    srt = good[key[~nan][::-1].argsort(kind="quicksort")][::-1]
    return table.take(np.concatenate([srt, idx[nan]]))


def run_pipeline(obj, cm, qs, chromnames, pvalue, mask, slow, comm=None, say=lambda *a, **k: None,
                 return_stages=False, group=5, ahead=2, adapt_group=True):
    """Stages 1-5 of compute (stripenn.py:120-152) for the units this (virtual) rank owns
    (plan_units: whole chromosomes, window ranges of over-sized ones).  `obj` is a getStripe-shaped
    object, `comm` a communicator of stripenn_b200.comm.

    Exchanges between ranks (typed arrays, see comm.py): the availability counts (one int64 per
    background chromosome); the background columns — device to device: every GPU's k_background
    writes into its slot range of one buffer, one in-place all-gather completes it everywhere,
    k_bg_sort reads it (with a Device that lacks the device-resident calls, e.g. the oracle-backed
    test double, the columns travel as float64 arrays instead); the raw candidates of a chromosome
    that was cut into window ranges; the final rows.  The searches of every rank start before the
    first exchange and keep the GPU busy through all of them — only the p-values wait for the
    background.

    Every device call goes through one DeviceQueue in an order that keeps the GPU busy while this
    thread does the host steps: maxpixel values, expected values and row flags of all
    chromosomes; the first candidate searches; the background samples (their sampling plan is
    drawn on the host meanwhile); then, chromosome by chromosome, search -> RemoveRedundant(size)
    on the host -> p-values (in groups of `group` chromosomes) -> RemoveRedundant(pvalue) on the
    host -> stripiness.  RemoveRedundant never pairs rows of different chromosomes
    (getStripe.py:1127), so pruning per chromosome and re-ordering at the end gives exactly the
    tables of the stage-by-stage flow (getStripe.extract_all + RemoveRedundant + scoringstripes)."""
    comm = comm or LocalComm()
    clock = StageClock()
    all_names = list(obj.all_chromnames)
    chromnames = list(chromnames)
    units = plan_units(all_names, obj.all_chromsizes, obj.resol, comm.world, selected=chromnames, tpp=comm.tpp)
    primary = {u["chrom"]: u["rank"] for u in units if u["piece"] == 0}
    mine = lambda c: primary.get(c, 0) == comm.rank
    my_sel = [c for c in chromnames if mine(c)]
    # search units of this rank, in -k order; a split chromosome's windows are searched by several
    # ranks (each holds the chromosome and derives the same maxpixel values) and merged by its primary
    my_units = sorted((u for u in units if u["search"] and u["rank"] == comm.rank),
                      key=lambda u: (chromnames.index(u["chrom"]), u["piece"]))
    split_chroms = [c for c in chromnames if any(u["chrom"] == c and u["npieces"] > 1 for u in units)]
    quant_sel = list(dict.fromkeys(my_sel + [u["chrom"] for u in my_units]))
    labels = [str(q * 100) + '%' for q in qs]
    cidx = {c: i for i, c in enumerate(chromnames)}
    dev_bg = hasattr(obj.dev, "background_dev")       # device-resident background columns
    dq = DeviceQueue(trace=bool(os.environ.get("SNN_TRACE")))
    t_pipeline = time.perf_counter()
    try:
        say('1. Maximum pixel value calculation ...')
        quant = obj.getQuantile_slow if slow else obj.getQuantile_original
        names2 = obj.background_chromosomes()
        todo = list(my_units)
        f_cand, f_mp = [], {}

        def submit_search():
            if todo:
                u = todo.pop(0)
                rng = None if u["npieces"] == 1 else (u["lo"], u["hi"])
                f_cand.append((u, dq.submit(lambda u=u, rng=rng: obj.search_candidates(
                    u["chrom"], f_mp[u["chrom"]].result()[u["chrom"]], rng))))
        # Every upload starts now, on the copy stream.  The maxpixel values of the chromosome that is
        # searched first come from a call of their own and its search is queued right behind them:
        # those kernels run while the other chromosomes are still arriving.
        first = [my_units[0]["chrom"]] if my_units else quant_sel[:1]
        rest = [c for c in quant_sel if c not in first]
        dq.submit(obj.prefetch, list(dict.fromkeys(first + rest + [c for c in names2 if mine(c)])))
        # With several processes every rank waits for every other rank's availability counts
        # (exchange 1) and background columns (exchange 2): the cheap row-flag kernels go FIRST there,
        # before the quantile passes and the first search — once a search of the other host thread
        # fills the GPU, every small kernel of this thread waits about a millisecond for SM space
        # (profiles/bench_r02o_n8.json: exchange 1 completed 24 ms into a 51 ms step).  Alone, the
        # first search starts at once and hides the uploads of the other chromosomes.
        early_flags = comm.nproc > 1
        f_avail = None
        if early_flags:
            say('3. Background distribution estimation ...')
            f_avail = dq.submit(obj.available_counts, [c for c in names2 if mine(c)])
        if first:
            f_mp[first[0]] = dq.submit(quant, cm, first, qs)
        say('4. Finding candidate stripes from each chromosome ...')
        submit_search()
        if rest:
            f = dq.submit(quant, cm, rest, qs)
            for c in rest:
                f_mp[c] = f
        say('2. Expected value calculation ...')
        f_ev = {c: dq.submit(obj.expected_values, c) for c in my_sel}
        if not early_flags:
            say('3. Background distribution estimation ...')
            f_avail = dq.submit(obj.available_counts, [c for c in names2 if mine(c)])
        if not early_flags:           # with several processes the background columns go in front of search #2
            for _ in range(ahead - 1):
                submit_search()
        # exchange 1: availability counts, one int64 per background chromosome (-1 = not mine)
        got = f_avail.result()
        counts = np.array([got.get(c, -1) for c in names2], dtype=np.int64)
        merged = np.max(np.stack(comm.all_gather_fixed(counts)), axis=0) if names2 else counts
        avail = {c: int(merged[i]) for i, c in enumerate(names2)}
        clock.lap('quantiles+expected+flags')
        names2b, budget = obj.sample_budget(names2, avail)
        n_of = {c: int(budget[names2b.index(c)]) for c in names2b}
        ncol = sum(n_of.values())
        # sampling plan on the host while the first searches run; its device part goes in next
        bg_host = None
        if dev_bg:
            slot0, per_proc, slots = background_layout(names2b, n_of, primary, comm.tpp, comm.nproc)
            d_buf = comm.bg_buffer(obj.dev, per_proc * comm.nproc)
            f_cols = [dq.submit(obj.background_columns_dev, c, obj.plan_chromosome(c, n_of[c]), d_buf, slot0[c],
                                n_of[c]) for c in names2b if mine(c)]
            for _ in range(ahead - 1 if early_flags else 1):
                submit_search()
            for f in f_cols:
                f.result()
            comm.bg_exchange(per_proc)                  # exchange 2: one all-gather between the GPUs
            f_bg = dq.submit(obj.install_background_dev, d_buf, ncol, slots)
            if return_stages:
                raw_bg = dq.submit(obj.dev.device_read, d_buf, (per_proc * comm.nproc, 4, 400)).result()
                full = (np.concatenate([raw_bg[slot0[c]:slot0[c] + n_of[c]] for c in names2b]).transpose(1, 2, 0)
                        if names2b else np.zeros((4, 400, 0)))
                bg_host = tuple(np.ascontiguousarray(full[t]) for t in range(4))
        else:
            f_cols = {c: dq.submit(obj.background_columns, c, obj.plan_chromosome(c, n_of[c]))
                      for c in names2b if mine(c)}
            for _ in range(ahead - 1 if early_flags else 1):
                submit_search()
            # exchange 2 (host form): sample-major float64 rows, chromosome index in front
            rows = [np.concatenate([np.full((f.result().shape[2], 1), names2b.index(c), dtype=np.float64),
                                    f.result().transpose(2, 0, 1).reshape(-1, 1600)], axis=1)
                    for c, f in f_cols.items()]
            mine_rows = np.concatenate(rows) if rows else np.zeros((0, 1601))
            allrows = np.concatenate(comm.all_gather_arrays(mine_rows))
            allrows = allrows[np.argsort(allrows[:, 0], kind="stable")]        # names2b order, draws in order
            full = allrows[:, 1:].reshape(-1, 4, 400).transpose(1, 2, 0) if len(allrows) else np.zeros((4, 400, 0))
            bg_host = tuple(np.ascontiguousarray(full[t]) for t in range(4))
            f_bg = dq.submit(obj._install_background, bg_host)
        clock.lap('background')
        MP = {c: f_mp[c].result()[c] for c in quant_sel}
        EV = {c: f_ev[c].result() for c in my_sel}
        # ---- search -> prune -> p-values (grouped) -> prune by p-value -> stripiness
        kept, pend, f_pv = [], [], []      # kept: (chrom, records, positions)
        final, f_st, all_pv = [], [], []
        # p-values in groups of `group` chromosomes, but never so large that a rank with few search
        # units computes all of them after its last search: with 8 GPUs x 2 host threads a thread owns
        # one or two chromosomes, and the chain last search -> p-values -> RemoveRedundant ->
        # stripiness -> tables is the serial tail of the whole step
        if adapt_group:
            group = max(1, min(group, (len(my_units) + 2) // 3))
        pv_next = [0]

        def consume_pvalues(block):
            """RemoveRedundant(pvalue) and the stripiness call of every p-value group that has
            finished (`block`: of every group), in submission order."""
            while pv_next[0] < len(f_pv) and (block or f_pv[pv_next[0]][1].done()):
                grp, f = f_pv[pv_next[0]]
                pv_next[0] += 1
                pv = f.result()
                off = 0
                rows = []
                for c, cand, pos in grp:
                    p = pv[off:off + len(cand)]
                    off += len(cand)
                    all_pv.append(p)
                    keep = gs.remove_redundant_mask(np.zeros(len(cand), np.int32), cand["frame"].astype(np.int64),
                                                    pos, cand["h"].astype(np.float64), cand["w"].astype(np.float64),
                                                    np.ascontiguousarray(p, dtype=np.float64), "pvalue")
                    rows.append((c, cand[keep], pos[keep], p[keep]))
                f_st.append((rows, dq.submit(obj.stripiness_of, [(c, pos) for c, _, pos, _ in rows], EV, mask)))

        def flush_group():
            if pend:
                grp = list(pend)
                del pend[:]
                f_pv.append((grp, dq.submit(obj.candidate_pvalues, grp)))
        def take(c, raw):
            cand, pos = obj.prune_candidates(raw)
            if len(cand):
                kept.append((c, cand, pos))
                pend.append((c, cand, pos))
            if len(pend) >= group:
                flush_group()
        i = 0
        pieces = []
        while i < len(f_cand):
            u, f = f_cand[i]
            i += 1
            raw = f.result()
            submit_search()
            if u["npieces"] == 1:
                take(u["chrom"], raw)
            else:
                rec = np.empty(len(raw), dtype=PIECE_DTYPE)
                rec["ck"], rec["piece"], rec["rec"] = cidx[u["chrom"]], u["piece"], raw
                pieces.append(rec)
            consume_pvalues(False)         # earlier groups: prune and queue their stripiness behind the next search
        if split_chroms:
            # exchange 3 (only when a chromosome was cut): the window ranges of a chromosome meet
            # at its primary, in window order, before the cross-window pruning
            sent = np.concatenate(pieces) if pieces else np.zeros(0, dtype=PIECE_DTYPE)
            gathered = np.concatenate(comm.all_gather_arrays(sent))
            for c in split_chroms:
                if not mine(c):
                    continue
                g = gathered[gathered["ck"] == cidx[c]]
                raw = g["rec"][np.argsort(g["piece"], kind="stable")]
                take(c, raw[np.argsort(raw["q"], kind="stable")])   # (q, window, ...) order of one call
        flush_group()
        f_bg.result()
        clock.lap('search+prune')
        consume_pvalues(True)
        clock.lap('pvalue+dedup')
        say('5. Stripiness calculation ...')
        for rows, f in f_st:
            st = f.result()
            off = 0
            for c, cand, pos, p in rows:
                final.append((c, cand, p, st[off:off + len(cand)]))
                off += len(cand)
        clock.lap('stripiness')
    finally:
        dq.close()
    # ---- result tables.  Exchange 4: fixed-dtype rows (maxpixel index, chromosome index, bp
    # coordinates, window sum, size, p-value, Stripiness); the DataFrames are built ONCE per process
    # from the merged rows.
    final = [r for r in final if len(r[1])]
    part = np.zeros(sum(len(r[1]) for r in final), dtype=FINAL_DTYPE)
    off = 0
    for c, cand, p, st in final:
        v = part[off:off + len(cand)]
        v["q"], v["ck"] = cand["q"], cidx[c]
        for k in ("pos1", "pos2", "pos3", "pos4", "total", "h", "w"):
            v[k] = cand[k]
        v["pvalue"], v["s"] = p, st
        off += len(cand)
    parts = comm.all_gather_arrays(part)

    def build():
        m = np.concatenate(parts)
        if not len(m):
            empty = pd.DataFrame(columns=[k for k in gs.RESULT_COLUMNS if k not in
                                          ('total', 'num', 'start', 'end', 'x', 'y', 'h', 'w', 'medpixel')] +
                                 ['pvalue', 'Stripiness'])
            return empty, empty[empty['pvalue'] < pvalue]
        # reference row order: maxpixel-major, then chromosomes in -k order, then emission order
        # (rows of one chromosome come from one rank, already in emission order): ONE stable sort of a
        # small integer key (16-bit keys take numpy's radix sort), and every column is taken through the
        # order from its field — re-ordering the 86-byte records first and pulling the fields out of the
        # copy was a third of the table time (3.5 -> 2.6 ms for the 6,353 rows of a cfg3 genome)
        nck = max(len(chromnames), 1)
        key = m["q"].astype(np.int64) * nck + m["ck"]
        order = np.argsort(key.astype(np.int16) if key.max(initial=0) < 32000 else key, kind="stable")
        col = lambda k: m[k][order]
        q, ck = col("q"), col("ck")
        # string columns: take from a small typed array (pandas would convert an object array
        # element by element); same dtype as a DataFrame built from Python strings
        sdt = pd.Series(np.array(["x"], dtype=object)).dtype
        chrs = pd.array([str(c) for c in chromnames], dtype=sdt).take(ck)
        p1, p2, p3, p4 = col("pos1"), col("pos2"), col("pos3"), col("pos4")
        with np.errstate(divide="ignore", invalid="ignore"):
            mean = col("total") / col("h") / col("w")                # getStripe.py:1097
        pv, st = col("pvalue"), col("s")
        table = pd.DataFrame({
            'chr': chrs, 'pos1': p1, 'pos2': p2, 'chr2': chrs, 'pos3': p3, 'pos4': p4,
            'length': p4 - p3 + 1, 'width': p2 - p1 + 1, 'Mean': mean,
            'maxpixel': pd.array(list(labels), dtype=sdt).take(q), 'pvalue': pv,
            'Stripiness': st}, copy=False)
        return table, filtered_sorted(table, pv, st, pvalue)
    result_table, res_filter = comm.once(build)
    clock.lap('tables')
    obj.last_host_ms = clock.ms
    if dq.trace is not None:
        obj.last_trace = [(n, (a - t_pipeline) * 1e3, (b - t_pipeline) * 1e3, (c - t_pipeline) * 1e3)
                          for n, a, b, c in dq.trace]
        obj.last_trace_total_ms = (time.perf_counter() - t_pipeline) * 1e3
    if return_stages:
        per_q = obj.candidate_tables(kept, np.concatenate(all_pv) if all_pv else np.zeros(0), qs)
        return result_table, res_filter, dict(MP=MP, EV=EV, bg=bg_host, per_q=per_q)
    return result_table, res_filter


def run_stages(cm: ContactMap, norm='KR', chrom='all', canny=2.0, minL=10, maxW=8,
               maxpixel='0.95,0.96,0.97,0.98,0.99', numcores=1, pvalue=0.1, mask='0', slow=False,
               bfilter=3, seed=123456789, device=None, device_id=0, gray_variant=0, verbose=False,
               return_stages=False, comm=None, streams=1):
    """Stages 1-5 of ``compute`` on an in-memory contact map: everything between "matrix
    available" and "result tables on the host" (the span the Mpixels/s metric is defined on).
    With a multi-rank `comm` every rank processes the chromosomes it owns and all ranks return
    the same full tables.  Returns (result_unfiltered, result_filtered) [+ the stage products]."""
    say = print if verbose else (lambda *a, **k: None)
    qs = list(map(float, maxpixel.split(','))) if isinstance(maxpixel, str) else list(maxpixel)
    chroms = chrom.split(',') if isinstance(chrom, str) else list(chrom)
    bal = resolve_norm(cm, norm, "compute", verbose)
    all_names, all_sizes, chromnames, chromsizes = compute_chromosomes(cm, chroms)
    if streams > 1 and comm is None and device is None:
        # `streams` host threads, each with its own context on the same GPU (see ThreadComm)
        def one(tc):
            d = Device(device_id)
            try:
                o = make_object(cm, bal, d, all_names, all_sizes, chromnames, chromsizes, minL, maxW, canny,
                                numcores, bfilter, seed, gray_variant)
                return run_pipeline(o, cm, qs, chromnames, pvalue, mask, slow, tc,
                                    say if tc.rank == 0 else (lambda *a, **k: None), return_stages)
            finally:
                d.close()
        return run_threads(one, streams)[0]
    own = device is None
    dev = device if device is not None else Device(device_id)
    try:
        obj = make_object(cm, bal, dev, all_names, all_sizes, chromnames, chromsizes, minL, maxW, canny,
                          numcores, bfilter, seed, gray_variant)
        return run_pipeline(obj, cm, qs, chromnames, pvalue, mask, slow, comm, say, return_stages)
    finally:
        if own:
            dev.close()


def compute(cool, out, norm, chrom, canny, minL, maxW, maxpixel, numcores, pvalue, mask, slow, bfilter,
            seed, device_id=0, assume_yes=False):
    """Drop-in for stripenn.compute (stripenn.py:64): same arguments, same files written."""
    np.seterr(divide='ignore', invalid='ignore')
    t_start = time.time()
    if out[-1] != '/':
        out += '/'
    makeOutDir(out, assume_yes)
    addlog(cool, out, norm, chrom, canny, minL, maxW, maxpixel, numcores, pvalue, mask, bfilter)
    print('Result will be stored in %s' % (out))
    chroms = chrom.split(',')
    cm = open_contact_map(cool, None if chroms[0] == 'all' else None)
    result_table, res_filter = run_stages(cm, norm, chrom, canny, minL, maxW, maxpixel, numcores, pvalue,
                                          mask, slow, bfilter, seed, device_id=device_id, verbose=True)
    result_table.to_csv(out + 'result_unfiltered.tsv', sep="\t", header=True, index=False)
    res_filter.to_csv(out + 'result_filtered.tsv', sep="\t", header=True, index=False)
    print('\n' + str(round((time.time() - t_start) / 60, 3)) + 'min taken.')
    print('Check the result stored in %s' % (out))
    return 0
