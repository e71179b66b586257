"""Development probe (runs on the GPU box): compares every device stage with the oracle on the
'tiny' / 'tiny2' maps and prints a report instead of stopping at the first mismatch."""
import json
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import numpy as np

from stripenn_b200 import synth
from stripenn_b200.engine import Device
from stripenn_b200 import getstripe as gs
from stripenn_b200.compute import run_stages, compute_chromosomes, resolve_norm
from snn_oracle import pipeline as pl, image as im, canny as cn, segments as sg, background as bgm, scoring as sc
from snn_oracle import driver
from snn_oracle.matrix import OracleMatrix

report = {}


def section(name):
    def deco(fn):
        t = time.time()
        try:
            report[name] = fn()
        except Exception:
            report[name] = "EXC " + traceback.format_exc()
        print("==", name, "(%.1fs)" % (time.time() - t), report[name], flush=True)
    return deco


def main(which):
    cm = synth.make_config(which)
    norm = "KR" if which == "tiny2" else "None"
    bal = resolve_norm(cm, norm, verbose=False)
    all_names, all_sizes, names, sizes = compute_chromosomes(cm, ["all"])
    mat = OracleMatrix(cm, bal)
    dev = Device(0, timing=True)
    lib = gs.DeviceMatrix(cm, bal, dev)
    obj = gs.getStripe(lib, cm.binsize, 10, 8, 2.0, all_names, names, all_sizes, sizes, 1, 3, 123456789)
    qs = [0.95, 0.97, 0.99]
    c0 = names[0]
    cid = lib.ensure(c0)

    @section("fetch")
    def _():
        a = dev.fetch(cid, 100, 400, 150, 500)
        b = mat.window(c0, 100, 400, 150, 500).astype(float)
        return dict(equal=bool(np.array_equal(a, b, equal_nan=True)))

    MPo = pl.quantiles(mat, names, qs)

    @section("quantiles")
    def _():
        MP = obj.getQuantile_original(cm, names, qs)
        return {c: dict(dev=MP[c].tolist(), ora=MPo[c].tolist(), equal=bool(np.array_equal(MP[c], MPo[c]))) for c in names}

    EVo = bgm.expected_by_distance(mat, names)

    @section("expected")
    def _():
        EV = obj.mpmean()
        return {c: float(np.nanmax(np.abs(np.array(EV[c]) - np.array(EVo[c])) / np.maximum(np.abs(np.array(EVo[c])), 1e-300))) for c in names}

    @section("plan")
    def _():
        n2, plans = obj.plan_samples()
        n2o, planso = bgm.plan_samples(mat, all_names, all_sizes, 123456789)
        ok = n2 == n2o and all(np.array_equal(a, np.array(b, dtype=np.int32).reshape(-1, 6)) for a, b in zip(plans, planso))
        return dict(names=n2, equal=bool(ok), n=[len(p) for p in plans])

    bgo = bgm.null_distribution(mat, all_names, all_sizes, 123456789)
    bgd = [None]

    @section("background")
    def _():
        bg = obj.nulldist()
        bgd[0] = bg
        out = {}
        for nm, a, b in zip(("blu", "bru", "bld", "brd"), bg, bgo):
            naneq = bool(np.array_equal(np.isnan(a), np.isnan(b)))
            with np.errstate(invalid="ignore", divide="ignore"):
                err = np.nanmax(np.abs(a - b) / np.maximum(np.abs(b), 1e-12)) if a.shape == b.shape else -1
            out[nm] = dict(shape=list(a.shape), shape_o=list(b.shape), nan_equal=naneq, relerr=float(err),
                           exact=bool(np.array_equal(a, b, equal_nan=True)))
        return out

    @section("stages")
    def _():
        out = {}
        M = float(MPo[c0][1])
        for frame in (0, 3, 5):
            fr = pl.load_frame(mat, c0, frame)
            if fr is None:
                out[frame] = "skipped by oracle"
                continue
            G = im.saturate(fr["D"], M)
            w = cn.gaussian_weights(2.0)
            for bi in (0, 2, 5):
                d = dev.debug_stages(cid, M, 2.0, 3, frame, bi)
                gray = im.gray_image(G, im.BRIGHTNESS[bi], 3)
                st = cn.canny(gray, 2.0, w, None, return_stages=True)
                vert = sg.vertical_line(st["edges"])
                res = dict(S=(d["S"], fr["S"]))
                if d["S"] == fr["S"]:
                    res["gray_neq"] = int((d["gray"] != gray).sum())
                    res["smoothed_neq"] = int((d["smoothed"] != st["smoothed"]).sum())
                    res["mag_neq"] = int((d["mag"] != st["mag"]).sum())
                    res["low_neq"] = int((d["low"].astype(bool) != st["low"]).sum())
                    res["edges_neq"] = int((d["edges"].astype(bool) != st["edges"]).sum())
                    res["vert_neq"] = int((d["vert"] != vert).sum())
                    res["n_edges"] = int(st["edges"].sum())
                out["f%d_b%d" % (frame, bi)] = res
        return out

    @section("extract")
    def _():
        out = {}
        MP = {c: MPo[c] for c in names}
        t = time.time()
        cand = dev.extract(cid, MPo[c0], 2.0, 10, 8, 3)
        out["extract_s"] = time.time() - t
        out["n_raw"] = int(len(cand))
        # oracle raw candidates per (q, frame)
        mism = []
        for qi, q in enumerate(qs):
            for idx in range(pl.n_frames(mat.nbins(c0))):
                fr = pl.load_frame(mat, c0, idx)
                if fr is None:
                    continue
                want = []
                G = im.saturate(fr["D"], MPo[c0][qi])
                w = cn.gaussian_weights(2.0)
                bleed = cn.bleed_over((fr["S"], fr["S"]), *w)
                for bi, b in enumerate(im.BRIGHTNESS):
                    gray = im.gray_image(G, b, 3)
                    edges = cn.canny(gray, 2.0, w, bleed)
                    vert = sg.vertical_line(edges)
                    for (x, y, ww, hh) in sg.search_pass(edges, vert, 10, 8):
                        want.append((bi, x, y, ww, hh))
                sel = cand[(cand["q"] == qi) & (cand["frame"] == idx)]
                got = [(int(r["bright"]), int(r["x"]), int(r["y"]), int(r["w"]), int(r["h"])) for r in sel]
                if got != want:
                    mism.append(dict(q=qi, frame=idx, got=got[:6], want=want[:6], ngot=len(got), nwant=len(want)))
        out["mismatch_frames"] = len(mism)
        out["first"] = mism[:3]
        return out

    @section("compute_end_to_end")
    def _():
        t = time.time()
        unf, fil = run_stages(cm, norm=norm, maxpixel="0.95,0.97,0.99", device=dev)
        td = time.time() - t
        ounf, ofil = driver.compute(cm, norm=norm, maxpixel="0.95,0.97,0.99")
        out = dict(dev_s=td, n=(len(unf), len(ounf["pos1"])), nfil=(len(fil), len(ofil["pos1"])))
        if len(unf) == len(ounf["pos1"]):
            for k in ("pos1", "pos2", "pos3", "pos4"):
                out[k] = bool(list(unf[k]) == list(ounf[k]))
            out["maxpixel"] = bool(list(unf["maxpixel"]) == list(ounf["maxpixel"]))
            for k in ("Mean", "pvalue", "Stripiness"):
                a = np.array(unf[k], float); b = np.array(ounf[k], float)
                with np.errstate(invalid="ignore"):
                    out[k] = float(np.nanmax(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))
                out[k + "_nan_eq"] = bool(np.array_equal(np.isnan(a), np.isnan(b)))
            out["fil_order"] = bool(list(fil["pos1"]) == list(ofil["pos1"]))
        return out

    @section("score_mode")
    def _():
        from stripenn_b200.score import score_table
        import pandas as pd
        tab = synth.make_stripe_table(cm, 120)
        df = pd.DataFrame({"chr": tab["chr"], "pos1": tab["pos1"], "pos2": tab["pos2"], "chr2": tab["chr"],
                           "pos3": tab["pos3"], "pos4": tab["pos4"]})
        res = score_table(cm, df, norm=norm, device=dev)
        o = driver.score(cm, tab, norm=norm)
        out = {}
        for k in o:
            a = np.array(res[k], float); b = np.array(o[k], float)
            with np.errstate(invalid="ignore"):
                out[k] = float(np.nanmax(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))
            out[k + "_nan_eq"] = bool(np.array_equal(np.isnan(a), np.isnan(b)))
        return out

    report["stage_times"] = dev.stage_times()
    print("== stage_times", report["stage_times"])
    dev.close()


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "tiny"
    main(which)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "gpu_check_%s.json" % which), "w") as f:
        json.dump(report, f, indent=1, default=str)
