#!/usr/bin/env python
"""Instruction mix of the product kernels from the built library's SASS (cuobjdump -sass): what the hot loop is made of,
and proof of what is NOT there (no tensor-core / TMA instructions: the blocks are <= 45 x 45 fp64 and the data a few KB).
usage: sass_summary.py [lib.so] > profiles/<round>_sass_summary.md      (no GPU needed)"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "faster_b200", "lib",
                                                           "libfaster_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
arch = re.findall(r"arch = (sm_\w+)", txt)
funcs, cur = collections.OrderedDict(), None
for ln in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        funcs[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and cur:
        funcs[cur][m.group(1).split(".")[0]] += 1
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
print("# SASS instruction mix of libfaster_b200.so\n")
print("arch: %s; %d kernels.  Static counts (instructions in the binary, not executed counts).\n" % (sorted(set(arch)), len(funcs)))
want = ["fq_solve_kernel_tILi10ELb1", "fq_solve_kernel_tILi10ELb0", "fq_solve_kernel_tILi15ELb1", "fq_bnb_level_kernelILi10ELb1", "fq_select_multi_kernel",
        "fq_pair_mid_kernel", "fq_dtbase_kernel", "fq_expand_grid_kernel"]
groups = [("fp64 math", ["DFMA", "DMUL", "DADD", "DSETP", "DMNMX", "MUFU"]), ("shared memory", ["LDS", "STS", "LDSM"]),
          ("global / constant", ["LDG", "STG", "LDC", "ULDC", "ATOMG", "ATOM", "RED"]), ("warp-level", ["SHFL", "REDUX", "VOTE", "VOTEU", "WARPSYNC", "BAR"]),
          ("tensor core / TMA / TMEM (none expected)", ["HMMA", "IMMA", "DMMA", "UTCMMA", "UTMALDG", "UTMASTG", "UBLKCP", "TCGEN05", "LDTM", "STTM", "UTCBAR"])]
for key in want:
    for name, c in funcs.items():
        if key in name:
            total = sum(c.values())
            print("## `%s`\n\n%d instructions" % (demangle(name)[:110], total))
            for g, ops in groups:
                parts = ["%s %d" % (o, c[o]) for o in ops if c[o]]
                print("- %s: %s" % (g, ", ".join(parts) if parts else "none"))
            print()
            break
allc = collections.Counter()
for c in funcs.values():
    allc.update(c)
tc = [o for o in allc if re.match(r"(HMMA|IMMA|DMMA|UTC|UTMA|UBLKCP|TCGEN|LDTM|STTM)", o)]
print("Tensor-core / TMA / TMEM opcodes anywhere in the library: %s" % (tc if tc else "none"))
