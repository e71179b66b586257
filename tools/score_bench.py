#!/usr/bin/env python
"""Score-mode micro-benchmark (BASELINE.json config 5 shape): N synthetic stripes re-scored on a
resident synthetic genome through Device.score_multi (one call = p-value + stripiness + O/E +
observed mean).  Prints one JSON line with the CUDA-event time of the scoring kernels and the
HBM roofline fraction of their algorithmic bytes.

    python tools/score_bench.py --config cfg2 --stripes 100000 --reps 5
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="cfg2")
    ap.add_argument("--stripes", type=int, default=100000)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--chroms", default="", help="comma list (default: all)")
    ap.add_argument("--tag", default="")
    a = ap.parse_args()
    from stripenn_b200 import compute as cp, getstripe as gs, synth
    from stripenn_b200.engine import Device
    chroms, binsize, balanced = synth.config_chroms(a.config)
    only = a.chroms.split(",") if a.chroms else None
    cm = synth.make_config(a.config, only=only)
    norm = "KR" if balanced else "None"
    bal = cp.resolve_norm(cm, norm, "compute", False)
    all_names, all_sizes, names, sizes = cp.compute_chromosomes(cm, ["all"])
    names = [c for c in names if c in cm.chroms]
    dev = Device(0, timing=True)
    lib = gs.DeviceMatrix(cm, bal, dev)
    obj = gs.getStripe(lib, cm.binsize, 10, 8, 2.0, all_names, names, all_sizes,
                       [cm.chroms[c].size_bp for c in names], 1, 3, 123456789, 0)
    obj.all_chromnames = names                      # background from the resident chromosomes only
    obj.all_chromsizes = np.array([cm.chroms[c].size_bp for c in names])
    bg = obj.nulldist()
    obj._install_background(bg)
    ex = obj.mpmean()
    tab = synth.make_stripe_table(cm, a.stripes, chroms=names)
    r = cm.binsize
    bs = int(50000 / r)
    pos = np.stack([tab["pos1"], tab["pos2"], tab["pos3"], tab["pos4"]], axis=1).astype(np.int64)
    chrs = np.asarray(tab["chr"]).astype(str)
    ids = np.array([cm.chromnames.index(c) for c in names], dtype=np.int32)
    lut = {c: k for k, c in enumerate(names)}
    index = np.array([lut[c] for c in chrs], dtype=np.int32)
    exs = np.stack([np.asarray(ex[c], dtype=np.float64) for c in names])
    h = -(-pos[:, 3] // r) - pos[:, 2] // r
    w = -(-pos[:, 1] // r) - pos[:, 0] // r
    alg_bytes = float(np.sum(8 * h * (w + 2 * bs)) + 48 * len(pos) + 4 * 400 * bg[0].shape[1] * 8)
    for _ in range(2):
        res = dev.score_multi(ids, index, pos, exs, True)
    dev.reset_stage_times()
    t0 = time.perf_counter()
    for _ in range(a.reps):
        res = dev.score_multi(ids, index, pos, exs, True)
    wall = (time.perf_counter() - t0) / a.reps * 1e3
    st = dev.stage_times()["score"]
    kms = st["ms"] / a.reps
    peak = 6650.0
    mp = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(mp):
        try:
            peak = float(json.load(open(mp))["hbm_gbs"])
        except Exception:
            pass
    ach = alg_bytes / 1e9 / (kms / 1e3)
    print(json.dumps({"tag": a.tag, "lib": os.environ.get("STRIPENN_B200_LIB", "default"), "config": a.config,
                      "stripes": len(pos), "kernel_ms": kms, "launches_per_call": st["launches"] // a.reps,
                      "wall_ms_per_call": wall, "algorithmic_bytes": alg_bytes, "achieved_gbs": ach,
                      "frac_of_hbm_peak": ach / peak, "peak_gbs": peak,
                      "nan": {k: int(np.isnan(v).sum()) for k, v in res.items()},
                      "checksum": {k: float(np.nansum(v)) for k, v in res.items()}}), flush=True)
    dev.close()


if __name__ == "__main__":
    main()
