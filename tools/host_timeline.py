"""Development probe (GPU box): timeline of the device calls of one compute step (SNN_TRACE=1):
when each call was submitted, started and finished on the device thread, and the host laps.
    python tools/host_timeline.py cfg1            # whole config
    python tools/host_timeline.py cfg3 chr1,chr9,chr20   # what one rank of an 8-GPU run holds
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["SNN_TRACE"] = "1"
from stripenn_b200 import compute as cp
from stripenn_b200 import getstripe as gs
from stripenn_b200 import synth
from stripenn_b200.engine import Device

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
chroms, binsize, balanced = synth.config_chroms(cfg)
if len(sys.argv) > 2:
    only = sys.argv[2].split(",")
    chroms = [c for c in chroms if c[0] in only]
cm = synth.make_map(chroms, binsize, balanced=balanced)
norm = "KR" if balanced else "None"
dev = Device(0, timing=True)
bal = cp.resolve_norm(cm, norm, "compute", False)
all_names, all_sizes, names, sizes = cp.compute_chromosomes(cm, ["all"])
qs = [0.95, 0.96, 0.97, 0.98, 0.99]
for it in range(4):
    for cid in list(dev.loaded):
        dev.unload_chrom(cid)
    obj = cp.make_object(cm, bal, dev, all_names, all_sizes, names, sizes, 10, 8, 2.0, 1, 3, 123456789, 0)
    dev.reset_stage_times()
    t = time.perf_counter()
    unf, fil = cp.run_pipeline(obj, cm, qs, names, 0.1, '0', False)
    wall = (time.perf_counter() - t) * 1e3
print("step %.2f ms, %d rows; host laps %s" % (wall, len(unf), {k: round(v, 2) for k, v in obj.last_host_ms.items()}))
print("kernel ms", {k: round(v["ms"], 2) for k, v in dev.stage_times().items() if v["ms"] > 0})
print("%-28s %9s %9s %9s %8s" % ("call", "submit", "start", "end", "dur"))
for n, a, b, c in obj.last_trace:
    print("%-28s %9.2f %9.2f %9.2f %8.2f" % (n, a, b, c, c - b))
print("pipeline total %.2f ms" % obj.last_trace_total_ms)
