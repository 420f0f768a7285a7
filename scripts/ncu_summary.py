"""Text summary of an `ncu --set full` report, in the format of profiles/*_ncu_full_summary.txt.

    python scripts/ncu_summary.py gpurun_out/<name>.ncu-rep [--mix MIN_EXEC_PER_WARP] > profiles/<name>_ncu_full_summary.txt

`--mix N` appends the SASS instruction mix of the instructions every warp executed at least N times (the steady-state
path), from the source page of the same report.
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__inst_executed_pipe_tmem.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "sm__cycles_elapsed.max", "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active",
]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True, check=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    raw = page(rep, "raw")
    head, units, vals = raw[0], raw[1], raw[2]
    col = {k: i for i, k in enumerate(head)}
    print("Kernel Name:", vals[col["Kernel Name"]])
    for k in KEYS:
        if k in col:
            print(f"{k}: {vals[col[k]]} {units[col[k]]}")
    stalls = []
    for k, i in col.items():
        if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio"):
            try:
                stalls.append((float(vals[i]), k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
    for v, k in sorted(stalls, reverse=True)[:8]:
        print(f"  stall {k} {round(v, 3)}")
    if "--mix" in sys.argv:
        thr = int(sys.argv[sys.argv.index("--mix") + 1])
        src = page(rep, "source")
        hdr = next(i for i, r in enumerate(src) if r and r[0] == "Address")
        c_exec, c_samp = src[hdr].index("Instructions Executed"), src[hdr].index("Warp Stall Sampling (All Samples)")
        rows = src[hdr + 1:]
        n_warps = max(int(r[c_exec]) for r in rows[:4])
        mix, samp = collections.Counter(), collections.Counter()
        for r in rows:
            if int(r[c_exec]) < n_warps * thr:
                continue
            tok = r[1].split()
            op = (tok[1] if tok[0].startswith("@") else tok[0]).split(".")[0]
            mix[op] += int(r[c_exec]) / n_warps
            samp[op] += int(r[c_samp])
        tot = sum(samp.values()) or 1
        print(f"instruction mix of the path every warp runs >= {thr} times ({n_warps} warps; executions per warp, share of stall samples):")
        for op, c in mix.most_common(16):
            print(f"  {op:10s} {c:9.1f}  {100 * samp[op] / tot:5.1f} %")
    print("----")


if __name__ == "__main__":
    main()
