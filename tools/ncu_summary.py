"""Summarise an `ncu --page raw --csv` dump: one block per kernel launch with the metrics that matter
for the roofline (duration, DRAM bytes, tensor-pipe activity, occupancy, top stall reasons).
usage: ncu -i rep.ncu-rep --page raw --csv | python tools/ncu_summary.py"""
import csv
import sys

rows = list(csv.reader(sys.stdin))
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__occupancy_limit_warps",
        "sm__inst_executed.sum", "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]
for r in rows[1:]:
    if len(r) < len(hdr) or not r[0].isdigit():
        continue
    print("==", r[idx["Kernel Name"]][:60], "grid", r[idx["Grid Size"]], "block", r[idx["Block Size"]])
    for k in KEYS:
        if k in idx:
            print(f"   {k} = {r[idx[k]]}")
    stalls = [(h, float(r[i])) for h, i in idx.items() if h.startswith("smsp__average_warps_issue_stalled_")
              and h.endswith("_per_issue_active.ratio") and r[i] not in ("", "n/a")]
    stalls.sort(key=lambda kv: -kv[1])
    print("   stalled warps per issue:", ", ".join(
        f"{h.split('issue_stalled_')[1].split('_per_issue')[0]}={v:.2f}" for h, v in stalls[:7]))
