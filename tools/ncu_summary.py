"""Summarise an .ncu-rep (raw page) into the handful of metrics the roofline discussion needs."""
import csv, subprocess, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
        print("kernel:", d.get("Kernel Name"), " grid", d.get("launch__grid_size"), "block", d.get("launch__block_size"))
        for k in KEYS:
            if k in d: print("  %-75s %s %s" % (k, d[k], u[k]))
        st = {h: float(v) for h, v in d.items() if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and v not in ("", "n/a")}
        print("  top stall reasons (warps per issue-active cycle):")
        for k, v in sorted(st.items(), key=lambda x: -x[1])[:6]:
            print("     %.3f %s" % (v, k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
if __name__ == "__main__":
    for p in sys.argv[1:]: main(p)
