#!/bin/bash
# ncu captures for profiles/ (run under gpurun): launch list of the bench command (a short run of the cfg4 chain), full
# set of the two sweep launches (source-level, -lineinfo) and the N=15 kernel of cfg5.  tools/profile_round.sh <tag>
set -u
tag=${1:-r02}
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 1 --inner 4 --quick --single-stream"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$tag.csv \
    $BENCH > gpurun_out/ncu_list_$tag.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:fq_solve_kernel_t -s 6 -c 3 -f -o gpurun_out/prof_$tag \
    $BENCH > gpurun_out/ncu_full_$tag.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:fq_solve_kernel_t -s 1 -c 1 -f -o gpurun_out/prof_cfg5_$tag \
    python bench.py --config cfg5 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_cfg5_$tag.log 2>&1
python tools/ncu_metrics.py gpurun_out/prof_$tag.ncu-rep gpurun_out/kernel_metrics_$tag.json "round 2 run $tag: cfg4 chain, whole (N=10,P=3) and safe (N=10,P=4) sweep launches, 65536 candidates each" > /dev/null 2>gpurun_out/ncu_metrics_$tag.err
python tools/ncu_metrics.py gpurun_out/prof_cfg5_$tag.ncu-rep gpurun_out/kernel_metrics_cfg5_$tag.json "round 2 run $tag: cfg5, N=15, 8 polytopes, 65536 candidates" > /dev/null 2>>gpurun_out/ncu_metrics_$tag.err
ls -la gpurun_out | grep $tag
