"""Summarise an .ncu-rep (ncu --set full) into the handful of numbers DESIGN.md / profiles/ quote."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "lts__t_sectors_op_write.sum", "lts__t_sectors_op_read.sum", "sm__cycles_elapsed.max",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__pcsamp_warps_issue_stalled_long_scoreboard",
        "local_load", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum"]
for r in rows[2:]:
    for k in keys:
        for i, h in enumerate(hdr):
            if h == k:
                print(f"{k}: {r[i]} {units[i]}")
    stalls = [(h, float(r[i])) for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and r[i] not in ("", "n/a")]
    for h, v in sorted(stalls, key=lambda t: -t[1])[:8]:
        print("  stall", h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), round(v, 3))
    print("----")
