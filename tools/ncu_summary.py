"""Summarise EVERY kernel of an `ncu --set full` report as one block per launch: duration, DRAM bytes, DRAM / SM
throughput, tensor-pipe activity, L2 hit rate, shared-memory bank conflicts, launch shape.

  python tools/ncu_summary.py report.ncu-rep [header text ...] > profiles/xyz.txt
"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size",
        "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg.per_second",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_lsu.sum"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for line in sys.argv[2:]:
    print("# " + line)
name_i = hdr.index("Kernel Name")
for vals in rows[2:]:
    if len(vals) != len(hdr):
        continue
    print("== " + vals[name_i][:160])
    for h, u, v in zip(hdr, units, vals):
        if h in KEYS:
            print("   %-70s %-14s %s" % (h, u, v))
