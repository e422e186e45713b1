#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box with `ncu -i`) into a small text file for profiles/."""
import csv
import re
import subprocess
import sys

PAT = re.compile(r"^(gpu__time_duration.sum|dram__bytes_read.sum|dram__bytes_write.sum|gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed|"
                 r"sm__throughput.avg.pct_of_peak_sustained_elapsed|sm__warps_active.avg.pct_of_peak_sustained_active|"
                 r"sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active|sm__inst_executed_pipe_tensor.*pct.*|"
                 r"launch__registers_per_thread|launch__grid_size|launch__block_size|launch__shared_mem_per_block_dynamic|"
                 r"l1tex__t_sector_hit_rate.pct|lts__t_sector_hit_rate.pct|smsp__inst_executed.sum|smsp__issue_active.avg.pct_of_peak_sustained_active|"
                 r"smsp__thread_inst_executed_per_inst_executed.ratio|smsp__average_warps_issue_stalled_(long_scoreboard|short_scoreboard|wait|barrier|"
                 r"math_pipe_throttle|mio_throttle|lg_throttle|not_selected|branch_resolving)_per_issue_active.ratio)$")


def main(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# summary of {rep} (ncu --set full --clock-control none); one block per captured launch\n")
        for r in rows[2:]:
            f.write(f"\n== {r[hdr.index('Kernel Name')]}  grid={r[hdr.index('Grid Size')] if 'Grid Size' in hdr else ''}\n")
            for h, u, v in zip(hdr, units, r):
                if PAT.match(h):
                    f.write(f"{h:90s} {v} {u}\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
