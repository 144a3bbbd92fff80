#!/usr/bin/env python
"""Summarises an .ncu-rep (raw page): duration, DRAM bytes, pipe utilisation, stall reasons.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("==", d.get("Kernel Name", "?")[:90])
    for k in keys:
        if k in d:
            print("   %-70s %s" % (k, d[k]))
    st = [(k, float(v)) for k, v in d.items() if "issue_stalled" in k and k.endswith("per_issue_active.ratio")]
    st.sort(key=lambda x: -x[1])
    print("   stalls (cycles per issue):", ", ".join("%s=%.2f" % (k.split("issue_stalled_")[1].replace("_per_issue_active.ratio", ""), v) for k, v in st[:7]))
