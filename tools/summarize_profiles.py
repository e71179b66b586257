"""Turn gpurun_out/*.ncu-rep and the launch list into small tracked summaries under profiles/.
Usage: python tools/summarize_profiles.py r01b"""
import csv
import glob
import io
import json
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)

WANT = {
    "gpu__time_duration.sum": "time",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_pct",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "fp64_pipe_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "xu_pipe_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pipe_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "fma_pipe_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "lsu_pipe_pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occupancy_pct",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "smsp__inst_executed.sum": "warp_insts",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pct",
}

summary = {}
for rep in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "prof_%s_*.ncu-rep" % tag))):
    name = os.path.basename(rep)[len("prof_%s_" % tag):-len(".ncu-rep")]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    if len(rows) < 3:
        continue
    hdr, units, vals = rows[0], rows[1], rows[-1]
    d = {"kernel": vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else name}
    for h, u, v in zip(hdr, units, vals):
        if h in WANT:
            try:
                d[WANT[h]] = float(v.replace(",", ""))
            except ValueError:
                d[WANT[h]] = v
            d[WANT[h] + "_unit"] = u
    summary[name] = d

# launch list -> per-kernel totals and shares
launches = os.path.join(ROOT, "gpurun_out", "launches_%s.csv" % tag)
shares = {}
if os.path.exists(launches):
    lines = [l for l in open(launches) if l.startswith('"')]
    rows = list(csv.reader(lines))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = defaultdict(float)
    cnt = defaultdict(int)
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}.get(r[ui], 1.0)
        k = r[ki].split("(")[0]
        tot[k] += v * scale
        cnt[k] += 1
    total = sum(tot.values())
    for k in sorted(tot, key=lambda k: -tot[k]):
        shares[k] = {"launches": cnt[k], "total_us": round(tot[k], 1), "share": round(tot[k] / total, 4)}

json.dump({"tag": tag, "kernels": summary, "launch_list": shares}, open(os.path.join(out_dir, "%s_ncu_summary.json" % tag), "w"), indent=1)

with open(os.path.join(out_dir, "%s_ncu_summary.md" % tag), "w") as f:
    f.write("# ncu summary %s\n\n" % tag)
    f.write("`ncu --set full --clock-control none --import-source on -k regex:^<kernel> -s 2 -c 1 python bench.py "
            "--steps 1 --warmup 0 --no-cpu-baseline --config cfg1` (one launch per kernel; cold-cache, serialised "
            "replays — compare shares and pipe percentages, not absolute times).\n\n")
    f.write("| kernel | time | grid x block | regs | DRAM rd+wr (MB) | DRAM % | SM % | issue % | fp64 % | xu % | alu % | lsu % | L1 % | L2 % | occ % |\n")
    f.write("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|\n")

    def mb(d, k):
        v = d.get(k, 0.0)
        u = d.get(k + "_unit", "byte")
        return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)

    for name, d in summary.items():
        f.write("| %s | %.1f %s | %d x %d | %d | %.1f + %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f |\n" % (
            name, d.get("time", 0), d.get("time_unit", ""), d.get("grid", 0), d.get("block", 0), d.get("regs", 0),
            mb(d, "dram_read"), mb(d, "dram_write"), d.get("dram_pct", 0), d.get("sm_pct", 0), d.get("issue_pct", 0),
            d.get("fp64_pipe_pct", 0), d.get("xu_pipe_pct", 0), d.get("alu_pipe_pct", 0), d.get("lsu_pipe_pct", 0),
            d.get("l1_pct", 0), d.get("l2_pct", 0), d.get("occupancy_pct", 0)))
    if shares:
        f.write("\n## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none` over one bench step)\n\n")
        f.write("| kernel | launches | total us | share |\n|---|---|---|---|\n")
        for k, v in shares.items():
            f.write("| %s | %d | %.1f | %.1f %% |\n" % (k, v["launches"], v["total_us"], 100 * v["share"]))
print(open(os.path.join(out_dir, "%s_ncu_summary.md" % tag)).read())
