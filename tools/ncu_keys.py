"""Print the handful of ncu raw metrics the K1/K2 notes quote: python tools/ncu_keys.py report.ncu-rep"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "smsp__cycles_active.avg", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sectors_op_write.sum",
        "lts__t_sectors_op_read.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "sm__inst_executed_pipe_lsu.sum", "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units = rows[0], rows[1]
for r in rows[2:]:
    print("==", r[h.index("Kernel Name")][:90])
    for i, k in enumerate(h):
        if k in KEYS or ("issue_stalled" in k and k.endswith("per_warp_active.pct")):
            try:
                v = float(r[i])
            except ValueError:
                continue
            if "issue_stalled" in k and v < 3:
                continue
            print(f"  {k:90s} {units[i]:10s} {r[i]}")
