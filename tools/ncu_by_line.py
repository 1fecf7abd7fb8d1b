#!/usr/bin/env python
"""Aggregates an ncu report's per-SASS-instruction counters by CUDA source line.

usage: ncu_by_line.py <report.ncu-rep> <lib.so> <kernel-substring> [launch-index] [top]
Joins `ncu --page source --print-source=sass --csv` (instruction order) with `nvdisasm -g` line annotations of the
kernel's cubin extracted from the .so.  Prints the top lines by executed warp instructions and by stall samples.
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

rep, so, kern = sys.argv[1], sys.argv[2], sys.argv[3]
launch = int(sys.argv[4]) if len(sys.argv) > 4 else 0
top = int(sys.argv[5]) if len(sys.argv) > 5 else 25

tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
lines_of = None
for f in sorted(os.listdir(tmp)):
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    cur, insts, inside = None, [], False
    for ln in txt.splitlines():
        if ln.lstrip().startswith(".section"):
            if inside and insts:
                break
            inside = (".text." in ln and kern in ln)
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
            insts.append((cur, ln.split("*/", 1)[1].strip().rstrip(";")))
    if insts:
        lines_of = insts
        break
assert lines_of, "kernel not found in any cubin"

out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source=sass"], capture_output=True, text=True).stdout
blocks = out.split('"Kernel Name"')
# newer ncu versions print more than one block per launch: take the launch-th block whose SASS length matches the cubin's
match = []
for b in blocks[1:]:
    rows = list(csv.reader(io.StringIO('"Kernel Name"' + b)))
    hdr = rows[1]
    body = [r for r in rows[2:] if len(r) == len(hdr)]
    if len(body) == len(lines_of) and "Instructions Executed" in hdr:
        match.append((hdr, body))
assert match, "no block of the report has %d instructions" % len(lines_of)
hdr, body = match[min(launch, len(match) - 1)]
ci = {h: i for i, h in enumerate(hdr)}
agg = defaultdict(lambda: [0, 0, 0, 0])
tot = [0, 0, 0, 0]
for (line, sass), r in zip(lines_of, body):
    v = [int(r[ci["Instructions Executed"]]), int(r[ci["Warp Stall Sampling (All Samples)"]]),
         int(float(r[ci["L1 Wavefronts Shared Excessive"]] or 0)), int(float(r[ci["L1 Wavefronts Shared"]] or 0))]
    for k in range(4):
        agg[line][k] += v[k]
        tot[k] += v[k]
_srcs = {}
def srcline(key):
    if not key:
        return "?"
    f, n = key
    if f not in _srcs:
        try:
            _srcs[f] = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "faster_b200", "csrc", f)).read().splitlines()
        except Exception:
            _srcs[f] = []
    L = _srcs[f]
    return "%s:%d %s" % (f.replace("fq_kernels", "k"), n, L[n - 1].strip()[:80] if n <= len(L) else "?")
print("total warp-inst %d  stall samples %d  smem wavefronts %d (excess %d)" % (tot[0], tot[1], tot[3], tot[2]))
for key, name in ((0, "instructions executed"), (1, "stall samples"), (3, "shared-memory wavefronts")):
    print("---- top lines by", name)
    for line, v in sorted(agg.items(), key=lambda kv: -kv[1][key])[:top]:
        print("inst %5.1f%% stall %5.1f%% smem-wavefronts %5.1f%% smem-excess %5.1f%% | %s" %
              (100.0 * v[0] / tot[0], 100.0 * v[1] / max(1, tot[1]), 100.0 * v[3] / max(1, tot[3]), 100.0 * v[2] / max(1, tot[2]),
               srcline(line)))
