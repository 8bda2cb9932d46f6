"""Condense an ncu report into the metric table kept under profiles/ (runs where ncu is installed, no GPU needed):

    python scripts/ncu_summary.py gpurun_out/prof_head_0.ncu-rep > profiles/ncu_head_0.txt
"""
import csv
import io
import subprocess
import sys

KEEP = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__time_duration.sum", "l1tex__throughput.avg.pct_of_peak_sustained_active", "launch__block_size",
        "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "lts__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct"]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = next(r for r in rows if "Kernel Name" in r)
    data = [r for r in rows[rows.index(hdr) + 1:] if len(r) == len(hdr) and r[0].strip().isdigit()]
    for r in data:
        rec = dict(zip(hdr, r))
        print("%-70s %s" % ("Kernel Name", rec["Kernel Name"]))
        for k in KEEP:
            if k in rec and rec[k] != "":
                print("%-70s %s" % (k, rec[k]))
        print()


if __name__ == "__main__":
    main()
