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


class LocalComm:
    """Single-process stand-in for the two host-side exchanges of the sharded pipeline."""
    rank = 0
    world = 1

    def all_gather(self, obj):
        return [obj]


class TorchComm:
    """torch.distributed (nccl or gloo) as the exchange: the payloads are the per-chromosome
    availability counts, the background columns (12.8 MB in total) and the candidate tables —
    host objects, gathered with all_gather_object.  No data-path collective exists (SURVEY.md
    §8e): every (chromosome, window) unit is independent."""

    def __init__(self):
        import torch.distributed as dist
        self.dist = dist
        self.rank = dist.get_rank()
        self.world = dist.get_world_size()

    def all_gather(self, obj):
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out


class ThreadComm:
    """Several host threads feeding ONE GPU through separate contexts/streams ("virtual ranks"):
    while one thread post-processes its chromosome on the host (table assembly, dedup), the
    kernels of the others keep the GPU busy.  Same exchange points as the multi-GPU pipeline,
    implemented with a barrier.  Create with ThreadComm.group(k)."""

    def __init__(self, rank, world, shared):
        self.rank, self.world, self._sh = rank, world, shared

    @classmethod
    def group(cls, world):
        import threading
        shared = {"slots": [None] * world, "barrier": threading.Barrier(world)}
        return [cls(r, world, shared) for r in range(world)]

    def all_gather(self, obj):
        sh = self._sh
        sh["slots"][self.rank] = obj
        sh["barrier"].wait()
        out = list(sh["slots"])
        sh["barrier"].wait()
        return out

    def once(self, fn):
        """fn() evaluated by rank 0 only; every rank gets the same object (threads share memory,
        and the GIL would serialise identical copies of the work anyway)."""
        sh = self._sh
        if self.rank == 0:
            try:
                sh["once"] = (fn(), None)
            except BaseException as e:  # noqa: BLE001 - re-raised on every rank
                sh["once"] = (None, e)
        sh["barrier"].wait()
        val, err = sh["once"]
        sh["barrier"].wait()
        if err is not None:
            raise err
        return val


def run_threads(fn, world):
    """Run fn(comm) on `world` host threads (one ThreadComm each); returns the per-rank results
    and re-raises the first exception."""
    import threading
    comms = ThreadComm.group(world)
    res, err = [None] * world, [None] * world

    def work(i):
        try:
            res[i] = fn(comms[i])
        except BaseException as e:  # noqa: BLE001 - re-raised below
            err[i] = e
            comms[i]._sh["barrier"].abort()

    ts = [threading.Thread(target=work, args=(i,)) for i in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    for e in err:
        if e is not None and not isinstance(e, threading.BrokenBarrierError):
            raise e
    for e in err:
        if e is not None:
            raise e
    return res


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


def plan_units(names, sizes_bp, binsize, world, split=True):
    """Work units of the sharded pipeline: whole chromosomes, except that a chromosome holding
    more than 1.25 x its fair share of the bins (total / world) is cut into window ranges
    (getStripe.py:841-842: windows of one chromosome are independent tasks) — a single-chromosome
    run (BASELINE configs 1 and 4) then still uses every GPU.  Units are assigned longest first
    (work ~ windows).  Returns a list of dicts {chrom, piece, npieces, lo, hi, rank}; piece 0's
    rank is the chromosome's PRIMARY (it runs the chromosome-wide stages and the host steps)."""
    import math
    nb = [int(math.ceil(int(s) / binsize)) for s in sizes_bp]
    total = max(sum(nb), 1)
    fair = total / max(world, 1)
    units = []
    for i, c in enumerate(names):
        nfr = max(1, int(math.ceil(nb[i] / 200)))
        k = 1
        if split and world > 1 and nb[i] > 1.25 * fair:
            k = max(1, min(world, nfr, int(round(nb[i] / fair))))
        for p in range(k):
            lo, hi = (nfr * p) // k, (nfr * (p + 1)) // k
            units.append(dict(chrom=c, piece=p, npieces=k, lo=lo, hi=hi, order=i, rank=0))
    load = [0] * world
    for u in sorted(units, key=lambda u: (-(u["hi"] - u["lo"]), u["order"], u["piece"])):
        r = min(range(world), key=lambda k: (load[k], k))
        u["rank"] = r
        load[r] += u["hi"] - u["lo"]
    return units


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

    def __init__(self):
        from concurrent.futures import ThreadPoolExecutor
        self._ex = ThreadPoolExecutor(max_workers=1, thread_name_prefix="snn-device")

    def submit(self, fn, *args, **kw):
        return self._ex.submit(fn, *args, **kw)

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
                 return_stages=False, group=5, ahead=2):
    """Stages 1-5 of compute (stripenn.py:120-152) for the units this rank owns (plan_units:
    whole chromosomes, window ranges of over-sized ones), with the two host-side exchanges the
    background needs (+ one for the window ranges of cut chromosomes).  `obj` is a
    getStripe-shaped object.

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
    units = plan_units(all_names, obj.all_chromsizes, obj.resol, comm.world)
    primary = {u["chrom"]: u["rank"] for u in units if u["piece"] == 0}
    mine = lambda c: primary.get(c, 0) == comm.rank
    my_sel = [c for c in chromnames if mine(c)]
    # search units of this rank, in -k order; a split chromosome's windows are searched by several
    # ranks (each holds the chromosome and derives the same maxpixel values) and merged by its primary
    my_units = sorted((u for u in units if u["rank"] == comm.rank and u["chrom"] in chromnames),
                      key=lambda u: (list(chromnames).index(u["chrom"]), u["piece"]))
    split_chroms = [c for c in chromnames if any(u["chrom"] == c and u["npieces"] > 1 for u in units)]
    quant_sel = list(dict.fromkeys(my_sel + [u["chrom"] for u in my_units]))
    labels = [str(q * 100) + '%' for q in qs]
    dq = DeviceQueue()
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
        say('3. Background distribution estimation ...')
        f_avail = dq.submit(obj.available_counts, [c for c in names2 if mine(c)])
        for _ in range(ahead - 1):
            submit_search()
        avail = {}
        for part in comm.all_gather(f_avail.result()):
            avail.update(part)
        clock.lap('quantiles+expected+flags')
        names2b, budget = obj.sample_budget(names2, avail)
        # sampling plan on the host while the first searches run; its device part goes in next
        f_cols = {c: dq.submit(obj.background_columns, c, obj.plan_chromosome(c, int(budget[names2b.index(c)])))
                  for c in names2b if mine(c)}
        submit_search()
        allcols = {}
        for part in comm.all_gather({c: f.result() for c, f in f_cols.items()}):
            allcols.update(part)
        full = np.concatenate([allcols[c] for c in names2b], axis=2) if names2b else np.zeros((4, 400, 0))
        bg = (full[0], full[1], full[2], full[3])
        f_bg = dq.submit(obj._install_background, bg)
        clock.lap('background')
        MP = {c: f_mp[c].result()[c] for c in quant_sel}
        EV = {c: f_ev[c].result() for c in my_sel}
        # ---- search -> prune -> p-values (grouped) -> prune by p-value -> stripiness
        kept, pend, f_pv = [], [], []      # kept: (chrom, records, positions)

        def flush_group():
            if pend:
                grp = list(pend)
                del pend[:]
                f_pv.append((grp, dq.submit(obj.candidate_pvalues, grp, bg)))
        def take(c, raw):
            cand, pos = obj.prune_candidates(raw)
            if len(cand):
                kept.append((c, cand, pos))
                pend.append((c, cand, pos))
            if len(pend) >= group:
                flush_group()
        i = 0
        pieces = {}
        while i < len(f_cand):
            u, f = f_cand[i]
            i += 1
            raw = f.result()
            submit_search()
            if u["npieces"] == 1:
                take(u["chrom"], raw)
            else:
                pieces[(u["chrom"], u["piece"])] = raw
        if split_chroms:
            # third host-side exchange (only when a chromosome was cut): the window ranges of a
            # chromosome meet at its primary, in window order, before the cross-window pruning
            gathered = {}
            for part in comm.all_gather(pieces):
                gathered.update(part)
            for c in split_chroms:
                if not mine(c):
                    continue
                k = max(u["npieces"] for u in units if u["chrom"] == c)
                raw = np.concatenate([gathered[(c, p)] for p in range(k)])
                take(c, raw[np.argsort(raw["q"], kind="stable")])   # (q, window, ...) order of one call
        flush_group()
        f_bg.result()
        clock.lap('search+prune')
        final, f_st, all_pv = [], [], []
        for grp, f in f_pv:
            pv = f.result()
            off = 0
            rows = []
            for c, cand, pos in grp:
                p = pv[off:off + len(cand)]
                off += len(cand)
                all_pv.append(p)
                keep = gs.remove_redundant_mask(np.zeros(len(cand), np.int32), cand["frame"].astype(np.int64), pos,
                                                cand["h"].astype(np.float64), cand["w"].astype(np.float64),
                                                np.ascontiguousarray(p, dtype=np.float64), "pvalue")
                rows.append((c, cand[keep], pos[keep], p[keep]))
            f_st.append((rows, dq.submit(obj.stripiness_of, [(c, pos) for c, _, pos, _ in rows], EV, mask)))
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
    # ---- result tables.  Ranks exchange plain arrays (maxpixel index, chromosome index, bp
    # coordinates, window sum, size, p-value, Stripiness); the DataFrames are built ONCE from the
    # merged arrays (by one thread when the ranks are host threads of one process).
    final = [r for r in final if len(r[1])]
    cidx = {c: i for i, c in enumerate(chromnames)}
    if final:
        allc = np.concatenate([r[1] for r in final])
        part = dict(q=allc["q"].astype(np.int64),
                    ck=np.concatenate([np.full(len(r[1]), cidx[r[0]], dtype=np.int64) for r in final]),
                    pos=np.stack([allc["pos1"], allc["pos2"], allc["pos3"], allc["pos4"]], axis=1),
                    total=allc["total"].astype(np.float64), h=allc["h"].astype(np.int64),
                    w=allc["w"].astype(np.int64), pvalue=np.concatenate([r[2] for r in final]),
                    s=np.concatenate([r[3] for r in final]))
    else:
        part = None
    parts = [p for p in (comm.all_gather(part) if comm.world > 1 else [part]) if p is not None]

    def build():
        if not parts:
            empty = pd.DataFrame(columns=[k for k in gs.RESULT_COLUMNS if k not in
                                          ('total', 'num', 'start', 'end', 'x', 'y', 'h', 'w', 'medpixel')] +
                                 ['pvalue', 'Stripiness'])
            return empty, empty[empty['pvalue'] < pvalue]
        m = {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}
        # reference row order: maxpixel-major, then chromosomes in -k order, then emission order
        # (rows of one chromosome come from one rank, already in emission order)
        order = np.lexsort((np.arange(len(m["q"])), m["ck"], m["q"]))
        m = {k: v[order] for k, v in m.items()}
        # string columns: take from a small typed array (pandas would convert an object array
        # element by element); same dtype as a DataFrame built from Python strings
        sdt = pd.Series(np.array(["x"], dtype=object)).dtype
        chrs = pd.array([str(c) for c in chromnames], dtype=sdt).take(m["ck"])
        pos = m["pos"]
        with np.errstate(divide="ignore", invalid="ignore"):
            mean = m["total"] / m["h"] / m["w"]                     # getStripe.py:1097
        table = pd.DataFrame({
            'chr': chrs, 'pos1': pos[:, 0], 'pos2': pos[:, 1], 'chr2': chrs, 'pos3': pos[:, 2], 'pos4': pos[:, 3],
            'length': pos[:, 3] - pos[:, 2] + 1, 'width': pos[:, 1] - pos[:, 0] + 1, 'Mean': mean,
            'maxpixel': pd.array(list(labels), dtype=sdt).take(m["q"]), 'pvalue': m["pvalue"],
            'Stripiness': m["s"]}, copy=False)
        return table, filtered_sorted(table, m["pvalue"], m["s"], pvalue)
    result_table, res_filter = comm.once(build) if hasattr(comm, "once") else build()
    clock.lap('tables')
    obj.last_host_ms = clock.ms
    if return_stages:
        per_q = obj.candidate_tables(kept, np.concatenate(all_pv) if all_pv else np.zeros(0), qs)
        return result_table, res_filter, dict(MP=MP, EV=EV, bg=bg, per_q=per_q)
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
