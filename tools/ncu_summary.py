"""Turn ncu reports (--set full) into the JSON summary kept under profiles/.
Usage: python tools/ncu_summary.py out.json label1=report1.ncu-rep [label2=report2.ncu-rep ...]"""
import csv, io, json, subprocess, sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__inst_executed.avg.per_cycle_active", "smsp__inst_executed.sum",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
           "launch__block_size", "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
out = []
for arg in sys.argv[2:]:
    label, rep = arg.split("=", 1)
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics", ",".join(METRICS)],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = {"capture": label}
        for h, u, v in zip(hdr, units, r):
            if h == "Kernel Name":
                d[h] = v
            elif h in METRICS:
                d[h] = (v + " " + u).strip()
        out.append(d)
json.dump(out, open(sys.argv[1], "w"), indent=1)
print(f"{len(out)} kernels -> {sys.argv[1]}")
