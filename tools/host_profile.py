"""Development probe (GPU box): cProfile of one compute step to see where host time goes."""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stripenn_b200 import synth
from stripenn_b200.compute import run_stages
from stripenn_b200.engine import Device

only = sys.argv[1].split(",") if len(sys.argv) > 1 else ["1", "2", "3", "19"]
chroms, binsize, _ = synth.config_chroms("cfg2")
cm = synth.make_map([c for c in chroms if c[0] in only], binsize)
dev = Device(0, timing=True)
for i in range(2):
    for cid in list(dev.loaded):
        dev.unload_chrom(cid)
    t = time.time()
    run_stages(cm, norm="None", device=dev)
    print("warm step", time.time() - t)
for cid in list(dev.loaded):
    dev.unload_chrom(cid)
dev.reset_stage_times()
pr = cProfile.Profile()
t = time.time()
pr.enable()
unf, fil = run_stages(cm, norm="None", device=dev)
pr.disable()
print("profiled step", time.time() - t, len(unf))
print({k: round(v["ms"], 2) for k, v in dev.stage_times().items()})
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
