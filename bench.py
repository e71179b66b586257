#!/usr/bin/env python
"""Benchmark of the `stripenn compute` hot path (BASELINE.json metric: Mpixels/s of banded Hi-C
scored — extract + score + pvalue — at 1/2/4/8 B200 next to the CPU path).

    python bench.py --gpus N --steps K --warmup W            # the B200 arm (one JSON line)
    python bench.py --impl reference --gpus N --steps K ...  # the CPU arm (oracle port, all cores)

A step = one pass of stages 1-5 of `compute` (maxpixel quantiles, expected values, background
sampling, candidate search over every 400-bin window x 5 maxpixel values x 6 brightness passes,
p-values, dedup, stripiness) over one synthetic genome.  Default workload = BASELINE.json
configs[2], the largest configuration that fits one GPU and the north star's target: synthetic
human whole genome at 5 kb (hg38 chr1-22+X; chrY is dropped by the reference's chromosome
filter), KR-balanced with NaN-weight bins, 606,219 bins searched, 3,044 windows,
-m 0.95,0.96,0.97,0.98,0.99.  Mpixels = Q * sum over windows of (end-start+1)^2 / 1e6
 = 2,410 Mpixels per genome.  --config cfg2 / cfg1 / cfg4 select the others.

`value`: pixels resident in HBM (COO arrays on the device) when the timed region starts;
`e2e`: the same step through the public API with pinned HOST buffers, H2D copies of the pixel
tables and D2H of candidates/results inside the timed region.
Multi-GPU (`--scaling strong`, the default): ONE genome sharded over the GPUs by chromosome —
stripenn_b200.compute.plan_units, longest first, over-sized chromosomes cut into window ranges —
with `--streams` host threads per GPU; the exchanges are typed (comm.py): availability counts,
ONE NCCL all-gather of the background columns between the GPUs, final rows.  `--scaling weak`:
one genome per rank (independent samples, different seeds), no exchange.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "Mpixels/s banded Hi-C scored (extract+score+pvalue)"
UNIT = "Mpixels/s"
QS = "0.95,0.96,0.97,0.98,0.99"
NQ = 5


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--config", default="cfg3")
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"])
    ap.add_argument("--streams", type=int, default=2, help="host threads (contexts/streams) feeding each GPU")
    ap.add_argument("--cpu-seconds", type=float, default=None,
                    help="time budget of one CPU-arm step (default: 20 s for cpu_baseline, 300 s / (steps + warmup) "
                         "clamped to 8..30 s for --impl reference)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-score-mode", action="store_true", help="skip the config-5 side measurement")
    return ap.parse_args()


def env_rank():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def genome_mpixels(cm, names):
    from stripenn_b200 import synth
    return NQ * sum(synth.frame_pixels(cm.nbins(c)) for c in names) / 1e6


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in out.strip().splitlines():
            p = [x.strip() for x in line.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
            except ValueError:
                continue
            for nme, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------ peaks
def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def fp64_peak():
    """FP64 FMA rate of this GPU from the micro-benchmark tools/fp64_peak.cu (MEASURED_PEAKS.json has no
    FP64 figure): run now if the binary was built (__graft_entry__.build()), else the committed
    measurement of this pool (profiles/r02a_fp64_peak.json)."""
    exe = os.path.join(ROOT, "tools", "bin", "fp64_peak")
    if os.path.exists(exe):
        try:
            out = subprocess.run([exe], capture_output=True, text=True, timeout=120).stdout.strip().splitlines()[-1]
            d = json.loads(out)
            return d["rows"], "measured now (tools/fp64_peak.cu, %s, %d SMs, attr clock %d MHz)" % (
                d["gpu"], d["sms"], d["clock_khz_attr"] // 1000)
        except Exception:
            pass
    f = os.path.join(ROOT, "profiles", "r02a_fp64_peak.json")
    if os.path.exists(f):
        try:
            return json.load(open(f))["rows"], "measured earlier on this pool (profiles/r02a_fp64_peak.json)"
        except Exception:
            pass
    return None, "unavailable"


CONFIG_TEXT = {"cfg1": "synthetic chr16-sized map at 5 kb", "cfg2": "synthetic mouse whole genome at 10 kb (chr1-19+X)",
               "cfg3": "synthetic human whole genome at 5 kb, KR-balanced (chrY dropped by the chromosome filter)",
               "cfg4": "synthetic human chr1 at 1 kb"}
NCU_FAMILY = {"gaussian": ("k_gauss_vm", "k_gauss_hm"), "sobel_nms": ("k_sobel_nms",), "gray": ("k_gray6",),
              "search": ("k_search",), "score": ("k_score_warp",), "background": ("k_background",),
              "quantile_select": ("k_radix_hist",), "k_gray_gv": ("k_gray_gv",), "k_gh_nms": ("k_gh_nms",),
              "k_score_fast": ("k_score_fast",)}
# FP64 operations of the REFERENCE per pass pixel (one pixel of one (window, maxpixel, brightness) image),
# counted from its literal operation sequence: what `roofline.achieved` credits.
#   k_gray_gv: saturate 4 (255*(M-D)/M/255) + brightness 1 + cv.filter2D 3x3 (9 mul + 9 add) = 23,
#              + scipy correlate1d along axis 0, 17 taps (8 pair adds + 9 mul + 8 adds) = 25
#   k_gh_nms:  correlate1d along axis 1 = 25, + ndi.sobel: two derivative passes (sub, mul) = 4,
#              two [1 2 1] passes (2 mul + 2 add) = 8 (float32 work — division, magnitude, NMS — not counted)
FP64_FLOPS_PER_PASS_PIXEL = {"k_gray_gv": 48.0, "k_gh_nms": 37.0, "gray": 23.0, "gaussian": 50.0, "sobel_nms": 12.0}
FULL_BATCH_WINDOWS = 51      # windows of one full image-chain launch (1536 // (5 * 6)); ncu captures use this shape


def ncu_traffic(family):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of a kernel family from
    the newest committed `ncu --set full` summary under profiles/ (one capture per kernel, taken on
    a full 51-window batch like the launches of this benchmark); mean over the family's kernels."""
    import glob
    best = None
    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_summary.json"))):
        try:
            k = json.load(open(f)).get("kernels", {})
        except Exception:
            continue
        if all(n in k and "dram_read" in k[n] for n in NCU_FAMILY.get(family, ("?",))):
            unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            tot = [k[n]["dram_read"] * unit.get(k[n].get("dram_read_unit", "Mbyte"), 1e6) +
                   k[n]["dram_write"] * unit.get(k[n].get("dram_write_unit", "Mbyte"), 1e6) for n in NCU_FAMILY[family]]
            best = (sum(tot) / len(tot), os.path.basename(f))
    return best


# ------------------------------------------------------------------------------------------ CPU arm
# The CPU arm times the reference's own implementation of the path on the box's host cores.
#   kind "reference": the UNMODIFIED reference (ysora/stripenn 1.1.65.22) imported from baseline/_ref
#       (installed there with `pip install --no-index --no-deps --target baseline/_ref /root/reference`,
#       git-ignored, travels to the GPU box) or $STRIPENN_REFERENCE_SRC, through oracle/refharness
#       (stub packages for cooler / matplotlib / scikit-image, which the image lacks);
#   kind "port": the numpy restatement oracle/snn_oracle when no reference tree is present.
# A step = the stages the B200 arm times — maxpixel quantiles, expected values, candidate search over
# every window x 5 maxpixel values x 6 brightness passes, p-values, dedup, stripiness — on a bounded
# SAMPLE of the workload: one synthetic chromosome of the same generator and resolution (5 kb, KR
# weights with NaN bins for cfg3), F windows long, F sized to the per-step time budget with the
# windows spread over all host cores (one task per window like the reference's joblib loop).  The
# background tables (nulldist: 1000 samples, one serial task per chromosome, ~1 % of a genome's CPU
# time) are computed ONCE before the steps; their measured time is charged to every step in
# proportion: t_nulldist x (tasks per genome = chromosomes) x (sample pixels / genome pixels) / cores.
GENOME = {"cfg3": (23, 2410.4), "cfg2": (20, 1041.0), "cfg1": (1, 78.0), "cfg4": (1, 995.0)}   # chromosomes, Mpixels


def reference_src():
    for p in (os.environ.get("STRIPENN_REFERENCE_SRC"), os.path.join(ROOT, "baseline", "_ref")):
        if p and os.path.isfile(os.path.join(p, "stripenn", "getStripe.py")):
            return p
    return None


def sample_mpixels(nbins, frames):
    return NQ * sum((min((i + 1) * 200 + 99, nbins - 1) - (0 if i == 0 else i * 200 - 100) + 1) ** 2
                    for i in frames) / 1e6


class CpuArm:
    def __init__(self, config, cores=None):
        from stripenn_b200 import synth
        self.cores = cores or os.cpu_count() or 1
        self.config = config if config in GENOME else "cfg3"
        chroms, self.binsize, self.balanced = synth.config_chroms(self.config)
        self.norm = "KR" if self.balanced else "None"
        self.qs = list(map(float, QS.split(",")))
        self.src = reference_src()
        self.kind = "reference" if self.src else "port"
        self.F = None
        self.t_null = None
        sys.path.insert(0, os.path.join(ROOT, "oracle"))

    # ---- sample construction
    def build(self, F):
        """Sample chromosome of F windows (F * 200 bins), stage objects, background tables."""
        from stripenn_b200 import synth
        self.F = int(F)
        nbins = self.F * 200
        self.cm = synth.make_map([("chr1", nbins * self.binsize - 17)], self.binsize, balanced=self.balanced)
        self.nbins = self.cm.nbins("chr1")
        self.frames = list(range(int(math.ceil(self.nbins / 200))))
        self.mpx = sample_mpixels(self.nbins, self.frames)
        t0 = time.time()
        if self.kind == "reference":
            import tempfile
            os.environ["STRIPENN_REFERENCE_SRC"] = self.src
            os.environ["SNN_REF_PROCS"] = str(self.cores)
            from refharness import run_reference as rr
            self.rr = rr
            tmp = tempfile.mkdtemp(prefix="snn_cpuarm_")
            path = os.path.join(tmp, "sample.npz")
            self.cm.save_npz(path)
            self.obj, self.lib = rr.make_object(path, norm=self.norm, numcores=self.cores)
            self.bg = self.obj.nulldist()
        else:
            from snn_oracle import background as bgm
            from snn_oracle.matrix import OracleMatrix
            from stripenn_b200.compute import resolve_norm
            self.mat = OracleMatrix(self.cm, resolve_norm(self.cm, self.norm, "compute", False))
            self.bg = bgm.null_distribution(self.mat, ["chr1"], np.array([self.cm.chromsizes[0]]), 123456789)
        self.t_null = time.time() - t0

    # ---- one step
    def step(self):
        t0 = time.time()
        if self.kind == "reference":
            import pandas as pd
            obj = self.obj
            MP = obj.getQuantile_original(self.lib, obj.chromnames, self.qs)      # stripenn.py:126
            EV = obj.mpmean()                                                      # :128
            res = pd.DataFrame()
            for i, q in enumerate(self.qs):                                        # :134-138 (p-values inside)
                res = pd.concat([res, obj.extract(MP, i, q, *self.bg)])
            res = obj.RemoveRedundant(res, "pvalue")                               # :144
            obj.scoringstripes(res, EV, "0")                                       # :147
            nrow = len(res)
        else:
            import multiprocessing as mp
            from snn_oracle import background as bgm, pipeline as pl, scoring as sc
            MP = pl.quantiles(self.mat, ["chr1"], self.qs)["chr1"]
            EV = bgm.expected_by_distance(self.mat, ["chr1"])
            _CPU_STATE.update(mat=self.mat, chrom="chr1", qs=self.qs, MP=MP)
            if self.cores > 1:
                with mp.get_context("fork").Pool(self.cores) as pool:
                    got = dict(pool.map(_cpu_window, [(i,) for i in self.frames], chunksize=1))
            else:
                got = dict(_cpu_window((i,)) for i in self.frames)
            per_q = []
            for qi in range(len(self.qs)):
                t = pl.remove_redundant(pl.concat_tables([got[i][qi] for i in self.frames]), "size")
                t["pvalue"] = sc.pvalues(self.mat, self.bg, t)
                per_q.append(t)
            final = pl.remove_redundant(pl.concat_tables(per_q), "pvalue")
            sc.stripiness(self.mat, final, EV)
            nrow = len(final["pos1"])
        dt = time.time() - t0
        nchrom, genome_mpx = GENOME[self.config]
        share = self.t_null * nchrom * (self.mpx / genome_mpx) / self.cores
        return dt + share, dt, share, nrow

    def describe(self, dt, share, nrow):
        what = ("UNMODIFIED reference (stripenn 1.1.65.22 from %s through oracle/refharness)" % os.path.relpath(self.src, ROOT)
                if self.kind == "reference" else "oracle port (numpy restatement, oracle/snn_oracle; no reference tree here)")
        return ("%s on a %d-window synthetic chromosome (%d bins at %d bp, --norm %s, same generator as %s), 5 maxpixel "
                "values, %d host processes = %.1f windows per core: quantiles + expected values + search + p-values + "
                "dedup + stripiness %.1f s per step, + %.2f s = share of the background stage (nulldist %.1f s per "
                "chromosome task, measured once) ; %d stripes" %
                (what, len(self.frames), self.nbins, self.binsize, self.norm, self.config, self.cores,
                 len(self.frames) / self.cores, dt, share, self.t_null, nrow))


def _cpu_window(args):
    """One search window x all maxpixel values with the oracle port (worker process)."""
    idx, = args
    from snn_oracle import pipeline as pl
    st = _CPU_STATE
    fr = pl.load_frame(st["mat"], st["chrom"], idx)
    out = []
    for qi, q in enumerate(st["qs"]):
        if fr is None:
            out.append(pl.empty_table())
        else:
            out.append(pl.stripe_search(fr, st["MP"][qi], q, st["chrom"], 10, 8, 2.0, 3))
    return idx, out


_CPU_STATE = {}


def run_cpu_arm(config, steps, warm, budget_s, cores=None):
    """Returns (per-step records of the timed steps, arm).  The sample is sized from one calibration
    step on `cores` windows so that a step fits `budget_s`."""
    arm = CpuArm(config, cores)
    arm.build(arm.cores)                      # calibration: one window per core
    tot, dt, share, nrow = arm.step()
    per_window_core_s = dt * arm.cores / max(len(arm.frames), 1)
    F = int(min(max(arm.cores, budget_s * arm.cores / max(per_window_core_s, 1e-3)), 8 * arm.cores))
    if F > len(arm.frames):
        arm.build(F)
    recs = []
    for i in range(warm + steps):
        r = arm.step()
        if i >= warm:
            recs.append(r)
    return recs, arm


def cpu_sample(config, budget_s, cores=None):
    """cpu_baseline of the B200 arm's line: one timed step of the CPU arm on a bounded sample."""
    recs, arm = run_cpu_arm(config, 1, 0, budget_s, cores)
    tot, dt, share, nrow = recs[0]
    return {"value": arm.mpx / tot, "unit": UNIT, "cores": arm.cores, "kind": arm.kind,
            "sample": arm.describe(dt, share, nrow)}


def reference_arm(args):
    rank, local, world = env_rank()
    if rank != 0:
        return
    steps, warm = max(1, args.steps), max(0, args.warmup)
    # the whole run must end within a few minutes: ~300 s over all steps, 8-30 s per step
    per = args.cpu_seconds if args.cpu_seconds else max(8.0, min(30.0, 300.0 / (steps + warm)))
    recs, arm = run_cpu_arm(args.config, steps, warm, per)
    tot_s = sum(r[0] for r in recs)
    value = arm.mpx * len(recs) / tot_s
    sample = arm.describe(recs[-1][1], recs[-1][2], recs[-1][3])
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": 1000.0 * tot_s / steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s: %s, --norm %s, -m %s, stages 1-5 of compute; each step = a bounded sample of it"
                                   % (arm.config, CONFIG_TEXT.get(arm.config, ""), arm.norm, QS),
                       "sample": sample, "sample_mpixels": arm.mpx},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": arm.cores, "kind": arm.kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------------------------------ score mode
def score_bench(devs, cm, stream, torch, n_stripes=100000):
    """BASELINE.json config 5 on the resident genome: re-score 100k synthetic candidate stripes
    (p-value + stripiness + O/E + observed mean, one fused kernel) through Device.score.  The
    chromosomes and the background tables are the ones the last compute step left on the device."""
    from stripenn_b200 import synth
    names = [c for c in cm.chromnames if c in cm.chroms]
    tab = synth.make_stripe_table(cm, n_stripes, chroms=names)
    r = cm.binsize
    bs = int(50000 / r)
    pos = np.stack([tab["pos1"], tab["pos2"], tab["pos3"], tab["pos4"]], axis=1).astype(np.int64)
    chrs = np.asarray(tab["chr"]).astype(str)
    groups = {c: np.nonzero(chrs == c)[0] for c in names}
    cid = {c: cm.chromnames.index(c) for c in names}
    owner = {}
    for d in devs:                      # each chromosome lives in the context that processed it
        for k in d.loaded:
            owner[k] = d
    names = [c for c in names if cid[c] in owner]
    ex = {c: owner[cid[c]].expected(cid[c]) for c in names}
    h = -(-pos[:, 3] // r) - pos[:, 2] // r
    w = -(-pos[:, 1] // r) - pos[:, 0] // r
    alg_bytes = float(np.sum(8 * h * (w + 2 * bs)) + 48 * len(pos) + 4 * 400 * 1000 * 8)

    # one multi-chromosome call per context (each chromosome lives in the context that loaded it)
    plan = []
    for d in devs:
        mine = [c for c in names if owner[cid[c]] is d]
        if not mine:
            continue
        idx = np.concatenate([groups[c] for c in mine])
        index = np.concatenate([np.full(len(groups[c]), k, dtype=np.int32) for k, c in enumerate(mine)])
        plan.append((d, np.array([cid[c] for c in mine], dtype=np.int32), index, idx,
                     np.ascontiguousarray(pos[idx]), np.stack([ex[c] for c in mine])))

    def run():
        out = np.empty(len(pos))
        for d, ids, index, idx, p, e in plan:
            out[idx] = d.score_multi(ids, index, p, e, True)["stripiness"]
        return out

    for _ in range(2):
        run()
    for d in devs:
        d.reset_stage_times()
        d.score_routing()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record(stream)
    for _ in range(reps):
        out = run()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    sts = [d.stage_times()["score"] for d in devs]
    st = {"ms": sum(x["ms"] for x in sts), "launches": sum(x["launches"] for x in sts)}
    kms = st["ms"] / reps
    routed = [sum(v) // reps for v in zip(*[d.score_routing() for d in devs])]
    peak, src = measured_peaks()
    ach = alg_bytes / 1e9 / (kms / 1e3)
    tr = ncu_traffic("k_score_fast")
    return {"workload": "config-5-like: %d synthetic stripes (h ~ lognormal(55), w 3..12, +-%d flank bins) on the "
                        "resident %d-bp genome, p-value + stripiness + O/E + observed mean" % (len(pos), bs, r),
            "stripes_per_s_e2e": len(pos) / (ms / 1e3), "ms_e2e": ms, "kernel_ms": kms,
            "kernel_launches": st["launches"] // reps,
            "window_mpixels_per_s": float(np.sum(h * (w + 2 * bs))) / 1e6 / (kms / 1e3),
            "nan_outputs": int(np.isnan(out).sum()),
            "stripes_routed": {"k_score_fast": routed[0], "k_score_warp": routed[1], "k_score": routed[2]},
            "kernel_ms_is": "CUDA events around the scoring kernels of a call (the copies of the stripe table and of "
                            "the results are outside; ms_e2e has them)",
            "roofline": {"kernel": "k_score_fast" if routed[0] >= max(routed[1], routed[2]) else "k_score_warp",
                         "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                         "frac": ach / peak, "algorithmic_bytes": alg_bytes,
                         "algorithmic": "8 B x h x (w + 2 x flank) window bytes + 48 B of results per stripe (DESIGN.md 3) "
                                        "+ the background tables once",
                         "traffic": tr[0] if tr else None, "traffic_source": tr[1] if tr else None,
                         "peak_source": src}}


# ------------------------------------------------------------------------------------------ B200 arm
def native_arm(args):
    import torch
    import torch.distributed as dist
    from stripenn_b200 import compute as cp
    from stripenn_b200 import getstripe as gs
    from stripenn_b200 import synth
    from stripenn_b200.engine import Device, default_band_width

    rank, local, world = env_rank()
    # host threads that wait on the GPU: 2 per host thread per process.  When all ranks together have
    # more of them than half the cores, the library waits on blocking events instead of spinning.
    if world * max(1, args.streams) * 2 > (os.cpu_count() or 1) // 2:
        os.environ.setdefault("SNN_BLOCKING_SYNC", "1")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 arm has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        # high-priority NCCL stream: the one device-side exchange (background columns) must not queue
        # behind the search kernels that keep every GPU full
        try:
            opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local), pg_options=opts)
        except Exception:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    strong = args.scaling == "strong" and world > 1

    chroms, binsize, balanced = synth.config_chroms(args.config)
    names = [c[0] for c in chroms]
    sizes = [c[1] for c in chroms]
    nstreams = max(1, args.streams)                 # host threads (virtual ranks) per GPU
    if strong:
        # the same plan run_pipeline derives: world x nstreams virtual ranks, whole chromosomes
        # longest-first, window ranges of an over-sized chromosome spread over several of them;
        # virtual rank v lives in process v // nstreams
        shell = synth.make_map(chroms, binsize, only=[])
        f_names, f_sizes, _, _ = cp.compute_chromosomes(shell, ["all"])
        units = cp.plan_units(f_names, f_sizes, binsize, world * nstreams, selected=f_names, tpp=nstreams)
        mine = [c for c in names if any(u["chrom"] == c and u["rank"] // nstreams == rank for u in units)]
        seed = synth.DEFAULT_SEED
    else:
        mine = names
        seed = synth.DEFAULT_SEED + rank          # one independent sample per rank
    t0 = time.time()
    cm = synth.make_config(args.config, seed=seed, only=mine)
    gen_s = time.time() - t0
    norm = "KR" if balanced else "None"
    bal = cp.resolve_norm(cm, norm, "compute", False)
    all_names, all_sizes, chromnames, chromsizes = cp.compute_chromosomes(cm, ["all"])
    if strong:   # pixels of the windows THIS process searches (a cut chromosome is shared)
        mpix_rank = NQ * sum(synth.frame_pixels_range(cm.nbins(u["chrom"]), u["lo"], u["hi"]) for u in units
                             if u["search"] and u["rank"] // nstreams == rank and u["chrom"] in chromnames) / 1e6
    else:
        mpix_rank = genome_mpixels(cm, [c for c in chromnames if c in cm.chroms])
    nnz_rank = sum(cm.chroms[c].nnz for c in cm.chroms)

    stream = torch.cuda.current_stream()
    devs, tstreams = [], []
    for k in range(nstreams):
        d = Device(local, timing=True)
        ts = torch.cuda.Stream() if nstreams > 1 else stream
        d.set_stream(ts.cuda_stream)
        devs.append(d)
        tstreams.append(ts)
    dev = devs[0]
    # communicators, one per host thread: threads of a process meet in shared memory; with several
    # processes thread 0 of each talks to the others through torch.distributed (NCCL)
    if strong:
        comms = cp.HybridComm.group(nstreams, cp.TorchComm())
        thread_of = lambda c: c.tc.rank
    elif nstreams > 1:
        comms = cp.ThreadComm.group(nstreams)
        thread_of = lambda c: c.rank
    else:
        comms = [cp.LocalComm()]
        thread_of = lambda c: 0

    # resident copies (HBM) and pinned host copies of the pixel tables
    resident, pinned = {}, {}
    for c, px in cm.chroms.items():
        host = [torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).pin_memory()
                for a in (px.bin1, px.bin2, px.count)]
        pinned[c] = host
        resident[c] = [h.cuda(non_blocking=True) for h in host]
    torch.cuda.synchronize()

    class ResidentMatrix(gs.DeviceMatrix):  # noqa: E306
        """DeviceMatrix whose chromosomes are built from pixel tables already in HBM."""
        def _upload(self, chrom, cid, need, async_):
            px = self.cm.chroms[str(chrom)]
            b1, b2, ct = resident[str(chrom)]
            self.dev.load_chrom_device(cid, px.nbins, self.binsize, px.size_bp, px.nnz, b1.data_ptr(),
                                       b2.data_ptr(), ct.data_ptr(), None, need or self.band_width)

    class PinnedMatrix(gs.DeviceMatrix):
        """DeviceMatrix fed from pinned host buffers: the H2D copies happen inside the timed step
        (asynchronously on the library's copy stream when the pipeline prefetches)."""
        def _upload(self, chrom, cid, need, async_):
            px = self.cm.chroms[str(chrom)]
            b1, b2, ct = (t.numpy() for t in pinned[str(chrom)])
            self.dev.load_chrom(cid, px.nbins, self.binsize, px.size_bp, b1, b2, ct, None,
                                need or self.band_width, async_=async_)

    qs = list(map(float, QS.split(",")))

    host_ms = {}

    def one_rank(d, c):
        t_in = time.perf_counter()
        for cid in list(d.loaded):
            d.unload_chrom(cid)
        lib = matrix_cls_now[0](cm, bal, d)
        obj = gs.getStripe(lib, cm.binsize, 10, 8, 2.0, all_names, chromnames, all_sizes, chromsizes, 1, 3,
                           123456789, 0)
        t0 = time.perf_counter()
        unf, fil = cp.run_pipeline(obj, cm, qs, chromnames, 0.1, '0', False, c)
        if d is devs[0]:
            host_ms.clear()
            host_ms.update(obj.last_host_ms)
            host_ms['unload+setup'] = (t0 - t_in) * 1e3
        return unf, fil

    matrix_cls_now = [None]

    def step(matrix_cls):
        matrix_cls_now[0] = matrix_cls
        if len(comms) == 1:
            return one_rank(dev, comms[0])
        return cp.run_threads(lambda c: one_rank(devs[thread_of(c)], c), comms)[0]

    step_walls = {}

    def timed(matrix_cls, k, w):
        for _ in range(w):
            step(matrix_cls)
        # Full garbage collections walk every container object of the process — with torch, pandas and
        # numpy imported that is 50-100 ms, and CPython runs them often while the process is young: two
        # steps of 75-105 ms in every 8-GPU run of 36 ms steps (profiles/bench_r02s/t_n8.json).  Collect
        # once now and move the survivors out of the collector's sight; the steps' own garbage is
        # still collected.
        import gc
        gc.collect()
        gc.freeze()
        for d in devs:
            d.reset_stage_times()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        walls = []
        for _ in range(k):
            t_s = time.perf_counter()
            res = step(matrix_cls)
            walls.append(round((time.perf_counter() - t_s) * 1e3, 2))
        e1.record(stream)
        torch.cuda.synchronize()
        step_walls[matrix_cls.__name__] = walls
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            dist.barrier()
        return ms, res

    steps, warm = max(1, args.steps), max(0, args.warmup)   # the default (3) satisfies W >= 3
    fp64_rows = fp64_peak() if rank == 0 else (None, "")
    # One clock sampler, on rank 0 (its own GPU), started a second before the work: nvidia-smi's start-up
    # (NVML attaches to every GPU of the box) holds driver locks, and eight of them starting inside
    # the timed region cost two steps of 75-105 ms in a run of 36 ms steps (profiles/bench_r02s_n8.json).
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(1.5)
    if world > 1:
        dist.barrier()
    ms_dev, (unf, fil) = timed(ResidentMatrix, steps, warm)
    clocks = sampler.stop() if rank == 0 else None
    stages = {}
    for d in devs:
        for k, v in d.stage_times().items():
            acc = stages.setdefault(k, {"ms": 0.0, "launches": 0})
            acc["ms"] += v["ms"]
            acc["launches"] += v["launches"]
    host_phases = {k: round(v, 2) for k, v in host_ms.items()}
    for d in devs:
        d.h2d_bytes = d.d2h_bytes = 0
    ms_e2e, _ = timed(PinnedMatrix, steps, 1)
    h2d_step = sum(d.h2d_bytes for d in devs) // (steps + 1)
    d2h_step = sum(d.d2h_bytes for d in devs) // (steps + 1)

    score_mode = None
    if rank == 0 and not args.no_score_mode:
        try:
            score_mode = score_bench(devs, cm, stream, torch)
        except Exception as e:  # secondary measurement must not break the headline
            score_mode = {"error": repr(e)}
    if world > 1:
        t = torch.tensor([mpix_rank, float(nnz_rank)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        mpix_total, nnz_total = float(t[0].item()), float(t[1].item())
    else:
        mpix_total, nnz_total = mpix_rank, float(nnz_rank)
    value = mpix_total * steps / (ms_dev / 1000.0)
    e2e = mpix_total * steps / (ms_e2e / 1000.0)

    shard_text = ""
    if strong:
        per = {}
        for u in units:
            if u["search"]:
                per.setdefault(u["rank"] // nstreams, []).append(
                    u["chrom"] if u["npieces"] == 1 else "%s[%d:%d]" % (u["chrom"], u["lo"], u["hi"]))
        shard_text = "; ".join("gpu%d: %s" % (g, ",".join(v)) for g, v in sorted(per.items()))
    if rank == 0:
        peak, peak_src = measured_peaks()
        launches = int(sum(v["launches"] for v in stages.values()))
        kern_ms = {k: v["ms"] / steps for k, v in stages.items()}
        # algorithmic bytes per step of each kernel family on THIS rank (DESIGN.md "Kernels")
        nb = sum(cm.nbins(c) for c in cm.chroms)
        bw = default_band_width(cm.binsize)
        fpx = mpix_rank * 1e6 / NQ
        n_str = max(len(unf), 1)
        alg = {
            "band_build": nnz_rank * 12 + nb * bw * 8,
            "quantile_select": nnz_rank * 12,
            "expected": nb * 400 * 8,
            "frame_prepare": fpx * 8,
            "gray": fpx * NQ * 6 * (8 + 4),
            "gaussian": fpx * NQ * 6 * 16,
            "sobel_nms": fpx * NQ * 6 * 4,
            "search": fpx * NQ * 6 * 2 / 8,
        }
        alg.update({"k_gray_gv": fpx * 8, "k_gh_nms": 0.0})
        per_stage = {k: {"ms_per_step": round(kern_ms[k], 4),
                         "achieved_gbs": (alg[k] / 1e9) / (kern_ms[k] / 1e3) if alg.get(k) and kern_ms[k] > 0 else None}
                     for k in kern_ms}
        # the dominant KERNEL (a stage that is one kernel), by device time inside the timed region
        single = [k for k in kern_ms if k in ("k_gray_gv", "k_gh_nms", "gray", "sobel_nms", "search", "score",
                                              "background", "expected", "row_flags")]
        top = max(single or list(kern_ms), key=lambda k: kern_ms[k])
        nl = max(stages[top]["launches"] / steps, 1)
        tr = ncu_traffic(top)
        rows, fsrc = fp64_rows
        fpeak = rows["dfma"]["tflops"] if rows else None
        pass_px = fpx * NQ * 6                                     # pass pixels per step on this rank
        batch_px = FULL_BATCH_WINDOWS * 400 * 400                  # window pixels of one full launch
        chain = [k for k in ("k_gray_gv", "k_gh_nms", "gray", "gaussian", "sobel_nms") if kern_ms.get(k, 0) > 0]
        chain_ms = sum(kern_ms[k] for k in chain)
        if top in FP64_FLOPS_PER_PASS_PIXEL and fpeak:
            flops = FP64_FLOPS_PER_PASS_PIXEL[top]
            ach = flops * pass_px / (kern_ms[top] / 1e3) / 1e12
            roofline = {"kernel": top, "bound": "fp64", "achieved": ach, "peak": fpeak, "unit": "TFLOP/s",
                        "frac": ach / fpeak,
                        "algorithmic": "%.0f FP64 operations of the reference per pass pixel (DESIGN.md §3) x %.4g pass "
                                       "pixels per step / kernel time per step" % (flops, pass_px),
                        "algorithmic_flops_per_launch": flops * batch_px * NQ * 6}
        else:
            ach = per_stage[top]["achieved_gbs"]
            roofline = {"kernel": top, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                        "frac": (ach / peak) if ach else None}
        roofline.update({
            "fp64_peak_tflops": fpeak, "fp64_peak_source": fsrc,
            "fp64_pipe_rates_per_clk_per_sm": ({k: rows[k]["ops_per_clk_per_sm_at_attr_clock"] for k in
                                                ("dfma", "dadd", "dmul", "cvt_widen", "cvt_narrow", "sqrtf_rn", "divf_rn")}
                                               if rows else None),
            # HBM view, for information (the image chain is not HBM-bound): the chain's
            # algorithmic bytes are the compacted window read once, 8 B per window pixel
            "hbm": {"peak": peak, "peak_source": peak_src, "unit": "GB/s",
                    "image_chain_kernels": chain,
                    "algorithmic_bytes_per_step": fpx * 8,
                    "achieved": (fpx * 8 / 1e9) / (chain_ms / 1e3) if chain_ms > 0 else None,
                    "frac": ((fpx * 8 / 1e9) / (chain_ms / 1e3) / peak) if chain_ms > 0 else None,
                    "algorithmic_bytes_per_launch": batch_px * 8},
            "traffic": tr[0] if tr else None,
            "traffic_source": ("%s: dram__bytes_read.sum + dram__bytes_write.sum of one %d-window launch of %s (ncu --set "
                               "full)" % (tr[1], FULL_BATCH_WINDOWS, "+".join(NCU_FAMILY[top]))) if tr else None,
            "traffic_over_algorithmic": (tr[0] / (batch_px * 8)) if tr else None,
            "launch_shape": "%d windows x %d maxpixel values x 6 brightness passes of 400 x 400 pixels (full batch)"
                            % (FULL_BATCH_WINDOWS, NQ),
            "launches_per_step": nl, "avg_launch_ms": kern_ms[top] / nl,
            "share_of_step": kern_ms[top] / (ms_dev / steps), "stages": per_stage,
            "gpu_ms_per_step": sum(kern_ms.values())})
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                cpu = cpu_sample(args.config, args.cpu_seconds or 20.0)
            except Exception as e:  # the checker must not break the measurement
                cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": "failed: %r" % e}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warm,
                "ms_per_step": ms_dev / steps, "higher_is_better": True, "scaling": args.scaling,
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "%s: %s (%d chromosomes, %d bins, %.0f Mpixels per genome), --norm %s, "
                                       "-m %s, stages 1-5 of compute" %
                                       (args.config, CONFIG_TEXT.get(args.config, "synthetic map at %d bp" % binsize),
                                        len(sizes), sum(int(math.ceil(s / binsize)) for s in sizes),
                                        mpix_rank if not strong else mpix_total, norm, QS),
                           "host_threads_per_gpu": nstreams,
                           "parallelism": ("1 genome per GPU (independent samples)" if not strong else
                                           "1 genome sharded over %d GPUs x %d host threads: plan_units (chromosomes "
                                           "longest-first, over-sized ones cut into window ranges): %s" %
                                           (world, nstreams, shard_text)),
                           "l2": "inputs larger than L2 (pixel tables %.0f MB per genome)" % (nnz_rank * 12 / 1e6),
                           "stripes_found": int(len(unf)), "stripes_filtered": int(len(fil)),
                           "generate_s": round(gen_s, 1)},
                "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d_step),
                        "d2h_bytes_per_step": int(d2h_step), "ms_per_step": ms_e2e / steps},
                "gpu_launches": launches, "host_phases_ms_last_step": host_phases,
                "host_wall_ms_per_step_rank0": {"resident": step_walls.get("ResidentMatrix"),
                                                "pinned_host": step_walls.get("PinnedMatrix")}, "score_mode": score_mode, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu}
        emit(line)
    for d in devs:
        d.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_OUT = [None]


def emit(line):
    """The ONE JSON line goes to the process's real stdout; everything else any library (or the
    reference's own progress prints) writes to stdout is sent to stderr."""
    out = _OUT[0] or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


if __name__ == "__main__":
    a = parse()
    sys.stdout.flush()
    _OUT[0] = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)                 # C-level writers (NCCL's version banner, child processes)
    sys.stdout = sys.stderr       # Python-level prints
    if a.impl == "reference":
        reference_arm(a)
    else:
        native_arm(a)
    _OUT[0].flush()
