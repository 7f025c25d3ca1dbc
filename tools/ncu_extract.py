"""Extract the judged metrics from an .ncu-rep (ncu --set full) into a small text table.
usage: python tools/ncu_extract.py gpurun_out/prof.ncu-rep > profiles/xxx.txt"""
import csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__cycles_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active",
        "sm__inst_executed_pipe_tensor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__occupancy_limit", "smsp__cycles_active.avg", "sm__pipe_fma_cycles_active"]
cols = [i for i, h in enumerate(hdr) if any(k in h for k in keys)]
name_i = hdr.index("Kernel Name")
for r in data:
    print("== " + r[name_i][:110])
    for i in cols:
        print("   %-95s %s %s" % (hdr[i][:95], r[i], units[i]))
