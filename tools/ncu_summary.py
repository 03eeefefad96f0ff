#!/usr/bin/env python
"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the handful of metrics
that matter for the roofline discussion.  usage: ncu_summary.py file.ncu-rep [more.ncu-rep ...]"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__cycles_elapsed.max", "sm cycles"),
    ("smsp__cycles_active.avg", "smsp active cycles"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor instr"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", "mem throughput %"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput %"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput2 %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1 throughput %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem"),
    ("smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "stall long_sb %"),
    ("smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "stall long_sb"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
]


def main():
    for path in sys.argv[1:]:
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        print(f"== {path}")
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            print(f"-- {name[:110]}")
            for key, label in WANT:
                if key in hdr:
                    i = hdr.index(key)
                    print(f"   {label:26s} {r[i]:>18s} {units[i]}")


if __name__ == "__main__":
    main()
