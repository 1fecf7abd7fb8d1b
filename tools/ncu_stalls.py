#!/usr/bin/env python
"""Warp-stall reasons of one kernel launch of an ncu report (--set full --import-source on), aggregated by opcode of the
stalled instruction and by CUDA source line (nvdisasm line table of the kernel's cubin, like ncu_by_line.py).
usage: ncu_stalls.py <report.ncu-rep> <lib.so> <kernel-substring> [launch-index]"""
import csv, io, os, re, subprocess, sys, tempfile
from collections import defaultdict
rep, so, kern = sys.argv[1], sys.argv[2], sys.argv[3]
launch = int(sys.argv[4]) if len(sys.argv) > 4 else 0
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
lines_of = None
for f in sorted(os.listdir(tmp)):
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    cur, insts, inside = None, [], False
    for ln in txt.splitlines():
        if ln.lstrip().startswith(".section"):
            if inside and insts: break
            inside = (".text." in ln and kern in ln); continue
        if not inside: continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln): insts.append((cur, ln.split("*/", 1)[1].strip().rstrip(";")))
    if insts: lines_of = insts; break
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source=sass"], capture_output=True, text=True).stdout
blocks = out.split('"Kernel Name"')
match = []
for b in blocks[1:]:
    rows = list(csv.reader(io.StringIO('"Kernel Name"' + b)))
    hdr = rows[1]; body = [r for r in rows[2:] if len(r) == len(hdr)]
    if len(body) == len(lines_of) and "Instructions Executed" in hdr: match.append((hdr, body))
hdr, body = match[min(launch, len(match) - 1)]
ci = {h: i for i, h in enumerate(hdr)}
CTRL = {"BRA","BSSY","BSYNC","WARPSYNC","NOP","YIELD","BREAK","BAR","CALL","RET","EXIT"}
agg = defaultdict(lambda: defaultdict(float)); tot = defaultdict(float)
for (line, sass), r in zip(lines_of, body):
    toks = sass.split(); op = toks[1] if toks[0].startswith("@") else toks[0]; op = op.split(".")[0]
    n = float(r[ci["Instructions Executed"]])
    d = agg[line]
    d["inst"] += n; tot["inst"] += n
    if op in CTRL: d["ctrl"] += n; tot["ctrl"] += n
    for h in ("stall_wait","stall_short_sb","stall_no_inst","stall_branch_resolving","stall_long_sb","stall_selected"):
        v = float(r[ci[h]]); d[h] += v; tot[h] += v
    d["samples"] += float(r[ci["Warp Stall Sampling (All Samples)"]]); tot["samples"] += float(r[ci["Warp Stall Sampling (All Samples)"]])
_s = {}
def src(key):
    if not key: return "?"
    f, n = key
    if f not in _s:
        try: _s[f] = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "faster_b200", "csrc", f)).read().splitlines()
        except Exception: _s[f] = []
    L = _s[f]; return "%s:%d %s" % (f.replace("fq_kernels","k"), n, L[n-1].strip()[:70] if n <= len(L) else "?")
print("totals:", {k: round(v) for k, v in tot.items()})
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
rt = defaultdict(float); byop = defaultdict(lambda: defaultdict(float)); mix = defaultdict(float)
for (line, sass), r in zip(lines_of, body):
    toks = sass.split(); op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
    mix[op] += float(r[ci["Instructions Executed"]])
    for h in reasons:
        v = float(r[ci[h]]); rt[h] += v; byop[h][op] += v
T = sum(rt.values())
print("---- stall reasons (% of samples) and the opcodes that wait")
for h, v in sorted(rt.items(), key=lambda x: -x[1])[:10]:
    print("%-24s %5.1f%%   %s" % (h, 100 * v / T, ", ".join("%s %.1f%%" % (o, 100 * x / T) for o, x in sorted(byop[h].items(), key=lambda x: -x[1])[:8])))
print("---- executed instruction mix:", ", ".join("%s %.1f%%" % (o, 100 * x / tot["inst"]) for o, x in sorted(mix.items(), key=lambda x: -x[1])[:20]))
for key in ("ctrl", "stall_wait", "stall_short_sb", "stall_branch_resolving", "stall_no_inst", "stall_long_sb"):
    print("---- top lines by", key, "(%% of all %s)" % ("instructions" if key == "ctrl" else "samples"))
    den = tot["inst"] if key == "ctrl" else tot["samples"]
    for line, v in sorted(agg.items(), key=lambda kv: -kv[1][key])[:14]:
        print("%5.2f%%  inst %4.1f%% | %s" % (100*v[key]/den, 100*v["inst"]/tot["inst"], src(line)))
