"""Print the metrics we track from an `ncu --page raw --csv` dump (one column per launch)."""
import csv
import sys

WANT = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_fma.avg.pct", "sm__inst_executed_pipe_alu.avg.pct",
        "sm__pipe_fma_cycles_active.avg.pct", "sm__pipe_alu_cycles_active.avg.pct", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct", "sm__throughput.avg.pct", "l1tex__throughput.avg.pct",
        "lts__throughput.avg.pct", "smsp__warps_eligible.avg.per_cycle_active", "smsp__average_warp", "smsp__average_warps_issue_stalled",
        "smsp__warp_issue_stalled", "sm__inst_executed_pipe_xu", "sm__inst_executed_pipe_lsu", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared"]


def main(path):
    rows = list(csv.reader(open(path)))
    h, units = rows[0], rows[1]
    for ci, name in enumerate(h):
        if any(w in name for w in WANT):
            vals = [r[ci] for r in rows[2:]]
            print(f"{name} [{units[ci]}]: {', '.join(vals)}")


if __name__ == "__main__":
    main(sys.argv[1])
