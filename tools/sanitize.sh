#!/bin/bash
# compute-sanitizer passes over a small parity workload (run under gpurun): memcheck + racecheck + synccheck.
set -u
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -x -q \
      -k "demo_corridor_batch or generic_and_specialised or no_polytopes or exact_miqp or device_side_fill" > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?" >> gpurun_out/sanitizer_$tool.log
done
tail -n 6 gpurun_out/sanitizer_*.log
