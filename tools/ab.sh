#!/bin/bash
# A/B of library variants on ONE box (run under gpurun): tools/ab.sh lib1.so lib2.so ...  (names under faster_b200/lib)
# Each variant: bench.py resident + e2e numbers and the oracle parity check; two rounds to see the noise.
for round in $(seq 1 ${AB_ROUNDS:-2}); do
  for lib in "$@"; do
    FQ_LIB=$PWD/faster_b200/lib/$lib python bench.py --steps 60 --cpu-seconds 0.5 > gpurun_out/ab_tmp.json 2>/dev/null
    python - "$lib" <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/ab_tmp.json").read().strip().splitlines()[-1])
    print(sys.argv[1], "resident %.2f M  e2e %.2f M  launch %.4f ms  parity mismatches %s  max rel cost err %.2e" %
          (d["value"] / 1e6, d["e2e"]["value"] / 1e6, d["roofline"]["kernel_ms"], d.get("parity", {}).get("flag_mismatches"),
           d.get("parity", {}).get("max_rel_cost_err", float("nan"))))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
  done
done
