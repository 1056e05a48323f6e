#!/usr/bin/env python
"""Selected metrics of every launch in an .ncu-rep (`ncu -i rep --page raw --csv`), one row per launch -> markdown / csv."""
import csv
import io
import subprocess
import sys

WANT = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__cluster_size", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    cols = []
    for w in WANT:
        hit = [i for i, h in enumerate(hdr) if h == w]
        if hit:
            cols.append(hit[0])
    w = csv.writer(sys.stdout)
    w.writerow([hdr[i] + (" [" + units[i] + "]" if units[i] else "") for i in cols])
    for r in data:
        w.writerow([r[i][:70] if hdr[i] == "Kernel Name" else r[i] for i in cols])


if __name__ == "__main__":
    main()
