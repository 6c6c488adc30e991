"""Key metrics per captured launch out of an `ncu -i x.ncu-rep --page raw --csv` export (a readable digest for profiles/).

    python tools/ncu_summary.py gpurun_out/r02_stencils.raw.csv > profiles/r02_full_capture_stencils.txt
"""
import csv
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
        "launch__shared_mem_per_block_static", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio",
        "smsp__average_warp_latency_issue_stalled_barrier.ratio",
        "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio",
        "smsp__average_warp_latency_issue_stalled_wait.ratio",
        "smsp__average_warp_latency_issue_stalled_lg_throttle.ratio",
        "smsp__average_warp_latency_issue_stalled_mio_throttle.ratio",
        "l1tex__t_sector_hit_rate.pct", "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
rows = list(csv.reader(ln for ln in open(sys.argv[1]) if not ln.startswith("==")))
head, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(head)}
for h, i in list(col.items()):  # section-prefixed duplicates ("FBSP.TriageCompute.dram__...") answer to the bare name too
    col.setdefault(h.split(".", 2)[-1] if h[:1].isupper() and h.count(".") >= 2 else h, i)
for r in rows[2:]:
    print("=== %s  grid %s block %s" % (r[col["Kernel Name"]][:110], r[col["Grid Size"]], r[col["Block Size"]]))
    for k in KEYS:
        if k in col:
            print("  %-78s %16s %s" % (k, r[col[k]], units[col[k]]))
