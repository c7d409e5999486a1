#!/usr/bin/env python
"""Summarise ncu captures into small tracked text files under profiles/: a .ncu-rep, or the `ncu -i x.ncu-rep --page raw --csv`
export of one (x.raw.csv — what comes back from the GPU box when the reports themselves are too large to carry).
  python profiles/summarize.py gpurun_out/prof_r1_c2.ncu-rep gpurun_out/prof_r2_c5.raw.csv [...]  -> profiles/<name>.summary.txt"""
import csv
import io
import os
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio"]


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    for rep in sys.argv[1:]:
        raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        if len(rows) < 3:
            print("no data in", rep); continue
        hdr, units = rows[0], rows[1]
        name = os.path.basename(rep).replace(".raw.csv", "").replace(".ncu-rep", "")
        out = [f"# {name}: ncu --set full --clock-control none (per-launch, cold-cache, serialised)"]
        for r in rows[2:]:
            out.append(f"\nkernel: {r[hdr.index('Kernel Name')]}   grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}")
            for w in WANT:
                if w in hdr:
                    out.append(f"  {w:80s} {r[hdr.index(w)]} {units[hdr.index(w)]}")
        open(os.path.join(here, name + ".summary.txt"), "w").write("\n".join(out) + "\n")
        print("wrote", name + ".summary.txt")


if __name__ == "__main__":
    main()
