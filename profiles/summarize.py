#!/usr/bin/env python
"""Condenses an `ncu --set full` report into the handful of numbers DESIGN.md / bench.py quote.
  python profiles/summarize.py gpurun_out/X.ncu-rep > profiles/X.summary.txt
Reads the report with `ncu -i X --page raw --csv` (no GPU needed)."""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid CTAs"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/CTA"), ("launch__occupancy_limit_registers", "CTA/SM limit (regs)"),
    ("launch__occupancy_limit_shared_mem", "CTA/SM limit (smem)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__inst_executed.avg.per_cycle_elapsed", "IPC (of 4)"),
    ("sm__instruction_throughput.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe active %"),
    ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "FMA-heavy pipe active %"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe active %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "stall long_scoreboard"),
    ("smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio", "stall math_pipe_throttle"),
    ("smsp__average_warp_latency_issue_stalled_mio_throttle.ratio", "stall mio_throttle"),
    ("smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio", "stall short_scoreboard"),
    ("smsp__average_warp_latency_issue_stalled_barrier.ratio", "stall barrier"),
    ("smsp__average_warp_latency_issue_stalled_lg_throttle.ratio", "stall lg_throttle"),
    ("smsp__average_warp_latency_issue_stalled_not_selected.ratio", "stall not_selected"),
    ("smsp__average_warp_latency_issue_stalled_wait.ratio", "stall wait"),
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    print("# %s  (ncu --set full --clock-control none; per launch; serialised, cold cache)" % rep)
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print("\n== %s  grid %s block %s" % (d["Kernel Name"], d.get("Grid Size", ""), d.get("Block Size", "")))
        for k, label in KEYS:
            if k in d and d[k] != "":
                print("  %-34s %14s %s" % (label, d[k], u[k]))
        try:
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
            tr = (float(d["dram__bytes_read.sum"]) * scale[u["dram__bytes_read.sum"]] +
                  float(d["dram__bytes_write.sum"]) * scale[u["dram__bytes_write.sum"]])
            print("  %-34s %14.3f %s" % ("DRAM traffic (read+write)", tr / 1e9, "Gbyte"))
        except (KeyError, ValueError):
            pass


if __name__ == "__main__":
    main()
