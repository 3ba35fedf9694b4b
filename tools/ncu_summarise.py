"""Turn an ncu report into the two small text files kept under profiles/:
   <out>_metrics.csv  (metric,unit,value for the headline metrics)  and  <out>_details.txt (the details page, trimmed).
Usage:  python tools/ncu_summarise.py gpurun_out/prof_icp_tc.ncu-rep profiles/r1_ncu_icp_tc
"""
import csv
import io
import subprocess
import sys

KEEP = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max",
        "sm__inst_executed.sum", "sm__inst_executed_pipe_tensor_subpipe_hmma.sum", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.sum", "smsp__cycles_active.avg", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__t_bytes.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    with open(out + "_metrics.csv", "w") as f:
        for i, h in enumerate(hdr):
            if h in KEEP or ("tensor" in h and "pct" in h and vals[i] not in ("0", "0.000000", "")):
                f.write(f"{h},{units[i]},{vals[i]}\n")
        for i, h in enumerate(hdr):
            if h == "Kernel Name":
                f.write(f"kernel,,{vals[i]}\n")
    det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
    lines = [ln.rstrip() for ln in det.splitlines() if not ln.startswith("    OPT") or True]
    with open(out + "_details.txt", "w") as f:
        f.write("\n".join(lines[:400]) + "\n")
    print(open(out + "_metrics.csv").read())


if __name__ == "__main__":
    main()
