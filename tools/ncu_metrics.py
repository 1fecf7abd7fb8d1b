#!/usr/bin/env python
"""Extracts the per-launch metrics bench.py quotes (roofline.traffic etc.) from an ncu report into a small JSON.
usage: ncu_metrics.py <report.ncu-rep> <out.json> <label>"""
import csv
import io
import json
import subprocess
import sys

rep, out, label = sys.argv[1], sys.argv[2], sys.argv[3]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units, body = rows[0], rows[1], rows[2:]
ci = {h: i for i, h in enumerate(hdr)}


def val(r, k):
    v = float(r[ci[k]].replace(",", ""))
    u = units[ci[k]].lower()
    u = u.split("/")[0]
    scale = {"kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0,
             "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0}.get(u, 1.0)
    return v * scale


launches = []
for r in body:
    launches.append({
        "kernel": r[ci["Kernel Name"]][:80], "grid": r[ci["Grid Size"]], "block": r[ci["Block Size"]],
        "duration_s": val(r, "gpu__time_duration.sum"),
        "dram_bytes": val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum"),
        "warp_instructions": val(r, "smsp__inst_executed.sum"),
        "registers_per_thread": val(r, "launch__registers_per_thread"),
        "dynamic_smem_bytes": val(r, "launch__shared_mem_per_block_dynamic"),
        "warps_active_pct": val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        "issue_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "fp64_pipe_active_pct": val(r, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
        "smem_bank_conflict_wavefronts": val(r, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
        "smem_wavefronts": val(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
        "lsu_data_pipe_pct_of_peak": val(r, "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
        "sm_cycles_elapsed_avg": val(r, "sm__cycles_elapsed.avg"),
        "sm_cycles_active_avg": val(r, "sm__cycles_active.avg"),
        "sm_cycles_active_max": val(r, "sm__cycles_active.max"),
    })
    # fp64 work actually executed: thread-level DFMA (2 flop), DADD, DMUL per elapsed SM cycle (summed over the SMSPs) x cycles
    try:
        cyc = val(r, "sm__cycles_elapsed.avg")
        per = {k: val(r, "smsp__sass_thread_inst_executed_op_%s_pred_on.sum.per_cycle_elapsed" % k) for k in ("dfma", "dadd", "dmul")}
        launches[-1]["fp64_flop"] = (2.0 * per["dfma"] + per["dadd"] + per["dmul"]) * cyc
        launches[-1]["fp64_thread_inst"] = {k: v * cyc for k, v in per.items()}
        launches[-1]["fp64_peak_flop_per_cycle"] = 2.0 * val(r, "sm__sass_thread_inst_executed_op_dfma_pred_on.sum.peak_sustained")
    except (KeyError, ValueError):
        pass
json.dump({"label": label, "report": rep, "note": "values under ncu replay (cold cache, serialised); per launch",
           "launches": launches}, open(out, "w"), indent=1)
print(json.dumps(launches, indent=1))
