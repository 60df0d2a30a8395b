"""Summarise `ncu --page raw --csv` / `--page source --csv` exports into the text files kept under profiles/."""
import csv, sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_active.avg.per_cycle_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "lts__t_sector_hit_rate.pct"]


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    for r in data:
        print("kernel:", r[ix["Kernel Name"]][:150] if "Kernel Name" in ix else "?")
        for k in KEYS:
            if k in ix:
                print("  %-82s %s %s" % (k, r[ix[k]], units[ix[k]]))
        for h in hdr:
            if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
                v = float(r[ix[h]])
                if v > 0.05:
                    print("  %-82s %.3f" % (h.replace("smsp__average_warps_issue_stalled_", "stall/issue: ").replace("_per_issue_active.ratio", ""), v))


def source(path, top=30):
    rows = list(csv.reader(open(path)))
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    tot = sum(int(r[ix["# Samples"]]) for r in data)
    print("warp-stall samples: %d; top instructions" % tot)
    for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:top]:
        st = {k.replace("stall_", ""): int(r[ix[k]]) for k in hdr if k.startswith("stall_") and "Not Issued" not in k and r[ix[k]] not in ("", "0")}
        st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:4])
        print("  %6s %10s  %-58s %s" % (r[ix["# Samples"]], r[ix["Instructions Executed"]], r[ix["Source"]].strip()[:58], st))


if __name__ == "__main__":
    raw(sys.argv[1])
    if len(sys.argv) > 2:
        source(sys.argv[2])
