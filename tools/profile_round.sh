#!/bin/bash
# ncu captures for profiles/ (run under gpurun): launch list, full set of the two bench launches, FP64 op counts,
# and one level kernel of the exact-MIQP branch-and-bound.
set -u
tag=${1:-r01g}
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_list_$tag.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:fq_solve_kernel_t -s 2 -c 2 -f -o gpurun_out/prof_$tag \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_$tag.log 2>&1
timeout 400 ncu --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum \
    --clock-control none -k regex:fq_solve_kernel_t -s 2 -c 2 --csv --log-file gpurun_out/fp64_$tag.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_fp64_$tag.log 2>&1
timeout 400 ncu --set full --clock-control none -k regex:fq_bnb_level -c 6 -f -o gpurun_out/prof_bnb_$tag \
    python tools/stress_exact.py 2 > gpurun_out/ncu_bnb_$tag.log 2>&1
ls -la gpurun_out | grep $tag
