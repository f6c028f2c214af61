"""Developer tool: one CSV row per kernel of one or more .ncu-rep files (the columns of profiles/rNN_metrics.csv).

    python tools/ncu_metrics_csv.py "label" report.ncu-rep [first_index count] ...  >> profiles/rNN_metrics.csv
"""
import csv, subprocess, sys
COLS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
label, rep = sys.argv[1], sys.argv[2]
first = int(sys.argv[3]) if len(sys.argv) > 3 else 0
count = int(sys.argv[4]) if len(sys.argv) > 4 else 1000
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
w = csv.writer(sys.stdout)
if "--header" in sys.argv:
    w.writerow(["capture", "kernel"] + COLS)
for r in rows[2 + first:2 + first + count]:
    vals = []
    for c in COLS:
        i = hdr.index(c)
        v = r[i].replace(",", "")
        u = units[i]
        try:                                     # normalise to us / MB like the round-1 file
            f = float(v)
            if u == "ms": f *= 1e3
            if u == "s": f *= 1e6
            if u == "Gbyte": f *= 1e3
            if u == "Kbyte": f *= 1e-3
            if u == "byte": f *= 1e-6
            v = "%.6f" % f
        except ValueError:
            pass
        vals.append(v)
    w.writerow([label, r[hdr.index("Kernel Name")][:90]] + vals)
