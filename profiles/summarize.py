#!/usr/bin/env python
"""Turns ncu output brought back in gpurun_out/ into the small summaries committed under profiles/.

  python profiles/summarize.py launches gpurun_out/launches.csv            > profiles/rNN_launches_X.txt
  python profiles/summarize.py full gpurun_out/raycast_full.ncu-rep [idx]  > profiles/rNN_X_ncu_full.txt

`launches` reads the CSV of `ncu --metrics gpu__time_duration.sum --clock-control none --csv`;
`full` reads a `--set full` report through `ncu -i ... --page raw --csv` and prints the metrics the
roofline discussion in DESIGN.md uses (one block per captured launch, or only launch `idx`)."""
import collections
import csv
import io
import subprocess
import sys

WANT = """gpu__time_duration.sum launch__grid_size launch__block_size launch__registers_per_thread
launch__occupancy_limit_registers launch__occupancy_limit_shared_mem launch__occupancy_limit_warps launch__waves_per_multiprocessor
sm__warps_active.avg.pct_of_peak_sustained_active smsp__warps_active.avg.per_cycle_active smsp__warps_eligible.avg.per_cycle_active
smsp__issue_active.avg.pct_of_peak_sustained_active smsp__inst_executed.sum smsp__thread_inst_executed_per_inst_executed.ratio
sm__throughput.avg.pct_of_peak_sustained_elapsed l1tex__throughput.avg.pct_of_peak_sustained_elapsed
lts__throughput.avg.pct_of_peak_sustained_elapsed dram__throughput.avg.pct_of_peak_sustained_elapsed
dram__bytes_read.sum dram__bytes_write.sum l1tex__t_sector_hit_rate.pct lts__t_sector_hit_rate.pct
l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum
l1tex__t_requests_pipe_lsu_mem_global_op_st.sum l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum
l1tex__data_pipe_lsu_wavefronts.sum l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
l1tex__t_requests_pipe_tex_mem_texture.sum l1tex__t_sectors_pipe_tex_mem_texture.sum
sm__inst_executed_pipe_lsu.sum sm__inst_executed_pipe_alu.sum sm__inst_executed_pipe_fma.sum sm__inst_executed_pipe_xu.sum
sm__inst_executed_pipe_tex.sum
smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_wait_per_issue_active.ratio smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio
smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio
smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio
smsp__average_warps_issue_stalled_membar_per_issue_active.ratio sm__cycles_active.avg sm__cycles_elapsed.max""".split()


def launches(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1e-3)
        agg.setdefault(r[ki], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print("# per-kernel launch times (us), ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised")
    print("%-100s %5s %12s %10s %7s" % ("kernel", "n", "total_us", "avg_us", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print("%-100s %5d %12.1f %10.1f %6.1f%%" % (k[:100], len(v), sum(v), sum(v) / len(v), 100 * sum(v) / tot))


def full(path, only=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for n, r in enumerate(rows[2:]):
        if only is not None and n != only:
            continue
        print("== launch %d: %s" % (n, r[hdr.index("Kernel Name")]))
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print("  %-86s %14s %s" % (w, r[i], units[i]))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else None)
