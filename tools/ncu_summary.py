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


def traffic(launch_csv, shapes_json, out_json, src):
    """Per-shape DRAM bytes / tensor-pipe % of one eager step: `ncu --metrics ...` launch list of
    `SAMK_BENCH_EAGER=1 SAMK_LAUNCH_LOG=shapes.json bench.py --profile-only` zipped with the logged launch shapes."""
    import collections
    import gzip
    import json
    op = gzip.open if launch_csv.endswith(".gz") else open
    with op(launch_csv, "rt") as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    per = collections.OrderedDict()          # launch id -> {metric: value}
    for r in rows:
        d = per.setdefault(r["ID"], {"name": r["Kernel Name"]})
        try:
            d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            pass
        d[r["Metric Name"] + "/unit"] = r["Metric Unit"]
    launches = list(per.values())
    marks = [i for i, d in enumerate(launches) if "bce_loss_kernel" in d["name"]]
    a, b = marks[-2], marks[-1]              # the last complete step: [loss(k-1) .. loss(k)) = bwd(k-1) + fwd(k)
    step = launches[a:b]
    shapes = json.load(open(shapes_json))    # logged in program order fwd(k) + bwd(k): rotate to bwd + fwd
    n_fwd = None
    gem = [s for s in shapes if s[0] == "gemm"]
    att = [s for s in shapes if s[0] != "gemm"]
    gk = [d for d in step if "gemm_tc_kernel" in d["name"]]
    fk = [d for d in step if "attn_fwd" in d["name"]]
    bk = [d for d in step if "attn_bwd2_kernel" in d["name"] or "attn_bwd_tc_kernel" in d["name"]]
    # forward launches come first in the log; in the ncu step window the backward half comes first
    n_attn_f = len([s for s in att if s[0] == "attn_fwd"])
    fwd_gemms = 0
    seen_f = 0
    for s in shapes:                          # GEMMs logged before the last forward attention ... simpler: count by the
        pass                                  # position of the loss: forward GEMMs are those logged before the first bwd one
    # the log has no explicit loss marker: the forward GEMM count equals the number of gemm_tc launches after the
    # loss kernel... use the window split instead: launches after index of first forward-only kernel (l2norm)
    first_fwd = next(i for i, d in enumerate(step) if "l2norm_kernel" in d["name"])
    gk_b = [d for d in step[:first_fwd] if "gemm_tc_kernel" in d["name"]]
    gk_f = [d for d in step[first_fwd:] if "gemm_tc_kernel" in d["name"]]
    assert len(gk_b) + len(gk_f) == len(gem), (len(gk_b), len(gk_f), len(gem))
    gem_f, gem_b = gem[:len(gk_f)], gem[len(gk_f):]
    out = {}

    def val(d, k):
        v = d.get(k, 0.0)
        u = d.get(k + "/unit", "")
        mult = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1.0, "ns": 1e-3, "ms": 1e3}.get(u, 1.0)
        return v * mult

    acc = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for sh, d in list(zip(gem_f, gk_f)) + list(zip(gem_b, gk_b)):
        key = "gemm_%dx%dx%d_%d%d" % (sh[1], sh[2], sh[3], int(sh[4]), int(sh[5]))
        a_ = acc[key]
        a_[0] += 1
        a_[1] += val(d, "dram__bytes_read.sum") + val(d, "dram__bytes_write.sum")
        a_[2] += val(d, "gpu__time_duration.sum")
        a_[3] += d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0)
    for sh, d in list(zip([s for s in att if s[0] == "attn_fwd"], fk)) + list(zip([s for s in att if s[0] == "attn_bwd"], bk)):
        key = "%s_L%d" % (sh[0], sh[1])
        a_ = acc[key]
        a_[0] += 1
        a_[1] += val(d, "dram__bytes_read.sum") + val(d, "dram__bytes_write.sum")
        a_[2] += val(d, "gpu__time_duration.sum")
        a_[3] += d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0)
    for key, (n, by, us, tp) in acc.items():
        out[key] = {"launches": n, "dram_bytes": by / n, "us_under_ncu": us / n, "tensor_pipe_pct": tp / n, "src": src}
    json.dump(out, open(out_json, "w"), indent=1, sort_keys=True)
    print("wrote %s: %d kernels" % (out_json, len(out)))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--traffic":
        traffic(*sys.argv[2:6])
    else:
        main(sys.argv[1:])
