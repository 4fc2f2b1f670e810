"""Text summary (launch configuration, instruction counts, DRAM bytes, stall reasons) of .ncu-rep files: python tools/ncu_summary.py a.ncu-rep ..."""
import csv, subprocess, sys, io
def summary(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
    keys = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
            "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "gpc__cycles_elapsed.avg.per_second",
            "smsp__average_warp_latency_per_inst_issued.ratio"]
    lines = []
    for k in keys:
        if k in d: lines.append("%-75s %s %s" % (k, d[k], u.get(k, "")))
    st = [(float(v), h) for h, v in d.items() if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and v not in ("", "n/a")]
    lines.append("warp cycles per issued instruction, by stall reason:")
    for v, h in sorted(st, reverse=True)[:8]:
        lines.append("   %.2f %s" % (v, h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
    return "\n".join(lines), d
if __name__ == "__main__":
    for rep in sys.argv[1:]:
        t, _ = summary(rep)
        print("==", rep); print(t)
