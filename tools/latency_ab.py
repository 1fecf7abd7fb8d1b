#!/usr/bin/env python
"""Single-replan latency of library variants on one box (run under gpurun): bench.latency_block per library (names under
faster_b200/lib), each in its own process, AB_ROUNDS rounds.  usage: latency_ab.py lib1.so lib2.so ..."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, ROOT)
    import bench
    from faster_b200 import capi
    s = capi.Solver(0)
    bench.latency_block(s, capi)                     # warm-up (clocks, arenas)
    print(json.dumps(bench.latency_block(s, capi)))
    sys.exit(0)
for r in range(int(os.environ.get("AB_ROUNDS", "2"))):
    for lib in sys.argv[1:]:
        env = dict(os.environ, FQ_LIB=os.path.join(ROOT, "faster_b200", "lib", lib))
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=env, capture_output=True, text=True)
        try:
            d = json.loads(p.stdout.strip().splitlines()[-1])
            e = d["early_exit"]
            print("%-26s sweep %.1f us (early exit %.1f)  exact %.1f (%.1f)  chained pair %.1f (%.1f)  N6 all-729 %.1f  same winners %s" %
                  (lib, d["value"], e["value"], d["exact_miqp"], e["exact_miqp"], d["chained_pair_one_corridor"],
                   e["chained_pair_one_corridor"], d["shipped_yaml_N6_P3_all_729_assignments"], e["same_winners"]), flush=True)
        except Exception as ex:
            print(lib, "FAILED", ex, p.stderr[-600:], flush=True)
