#!/usr/bin/env python
"""Condense `ncu --page raw --csv` exports into the handful of metrics the roofline discussion needs.

    python tools/ncu_summary.py gpurun_out/*_raw.csv > profiles/<tag>_ncu_summary.txt
"""
import csv
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor insts"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % (active)"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor hmma subpipe %"),
    ("sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor hmma op %"),
    ("sm__inst_executed.sum", "warp insts"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem"),
    ("smsp__cycles_active.avg", "SMSP active cycles"),
    ("sm__cycles_elapsed.max", "SM cycles elapsed"),
    ("smsp__inst_executed.sum", "warp insts (smsp)"),
    ("l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum", "L1 global load bytes"),
    ("l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum", "L1 global store bytes"),
    ("smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "stall long_scoreboard"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_sb /issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier /issue"),
    ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "stall membar /issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_sb /issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait /issue"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math throttle /issue"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg throttle /issue"),
    ("smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio", "stall sleeping /issue"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no_inst /issue"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected /issue"),
    ("smsp__average_warps_issue_stalled_selected_per_issue_active.ratio", "selected /issue"),
    ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "stall dispatch /issue"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall branch /issue"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio throttle /issue"),
    ("smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio", "stall tex throttle /issue"),
    ("smsp__average_warps_issue_stalled_drain_per_issue_active.ratio", "stall drain /issue"),
    ("smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio", "stall imc_miss /issue"),
    ("smsp__average_warps_issue_stalled_gmma_per_issue_active.ratio", "stall gmma /issue"),
]


def main(paths):
    for path in paths:
        with open(path) as f:
            lines = [l for l in f if not l.startswith("==")]
        rows = list(csv.reader(lines))
        if len(rows) < 3:
            print("%s: empty" % path)
            continue
        head, units = rows[0], rows[1]
        idx = {n: i for i, n in enumerate(head)}
        print("==== %s" % path)
        for r in rows[2:]:
            print("-- %s  (id %s)" % (r[idx["Kernel Name"]][:90], r[idx["ID"]]))
            for k, label in KEYS:
                if k in idx:
                    print("   %-28s %18s %s" % (label, r[idx[k]], units[idx[k]]))
            tens = [n for n in head if "utchmma" in n and n.endswith(".avg.pct_of_peak_sustained_elapsed")]
            for n in tens:
                v = r[idx[n]]
                if v not in ("0", "", "n/a"):
                    print("   %-60s %14s %s" % (n, v, units[idx[n]]))


if __name__ == "__main__":
    main(sys.argv[1:])
