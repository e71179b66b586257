#!/usr/bin/env python
"""Join an ncu report's per-SASS-instruction counters with nvdisasm line info and print the
hottest source lines of one kernel (instructions executed, stall samples).

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep stripenn_b200/lib/libstripenn_b200.so score_warp k_score_warp
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def main():
    rep, so, tu, kern = sys.argv[1:5]
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(tmp) if f.startswith(tu + ".")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    # walk the function: line markers then instructions
    lines_of = []   # per instruction: source line
    cur = None
    infun = False
    for ln in dis.splitlines():
        if ln.startswith(".text."):
            infun = kern in ln
            continue
        if not infun:
            continue
        m = re.search(r'//## File ".*?([^/"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s", ln):
            lines_of.append(cur)
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + (["--launch-skip", os.environ["NCU_SKIP"], "--launch-count", "1"] if os.environ.get("NCU_SKIP") else []), capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = next(r for r in rows if "Instructions Executed" in r)
    start = rows.index(hdr) + 1
    ii, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
    agg = {}
    tot = tots = 0
    body = []
    for r in rows[start:]:          # newer ncu versions repeat the table (second view): stop at its header
        if "Instructions Executed" in r:
            break
        if len(r) == len(hdr):
            body.append(r)
    if len(body) != len(lines_of):
        print("warning: %d SASS rows in the report, %d in the disassembly" % (len(body), len(lines_of)))
    for k, r in enumerate(body):
        key = lines_of[k] if k < len(lines_of) else None
        n, s = int(r[ii] or 0), int(r[si] or 0)
        a = agg.setdefault(key, [0, 0, 0])
        a[0] += n
        a[1] += s
        a[2] += 1
        tot += n
        tots += s
    src = {}
    print("total warp instructions %d, samples %d" % (tot, tots))
    for key, (n, s, k) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        text = ""
        if key:
            f = key[0]
            if f not in src:
                p = os.path.join(os.path.dirname(os.path.abspath(so)), "..", "csrc", f)
                src[f] = open(p).read().splitlines() if os.path.exists(p) else []
            if 0 < key[1] <= len(src[f]):
                text = src[f][key[1] - 1].strip()[:90]
        print("%5.1f%% inst %5.1f%% stall %4d sass  %s:%s  %s" % (100 * n / max(tot, 1), 100 * s / max(tots, 1), k,
                                                                 key[0] if key else "?", key[1] if key else "?", text))


if __name__ == "__main__":
    main()
