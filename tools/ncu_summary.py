"""Developer tool: key counters of the first kernel in an .ncu-rep (issue rate, stalls, shared-memory wavefronts, traffic)."""
import csv, subprocess, sys
rep = sys.argv[1]
idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2 + idx]
want = ['Kernel Name', 'gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'dram__bytes_read.sum',
        'dram__bytes_write.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct', 'launch__registers_per_thread',
        'smsp__sass_inst_executed_op_shared_ld.sum', 'smsp__sass_inst_executed_op_shared_st.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed']
for i, h in enumerate(hdr):
    if h in want:
        print("%-70s %-10s %s" % (h, units[i], vals[i][:90]))
st = []
for i, h in enumerate(hdr):
    if 'issue_stalled' in h and h.endswith('per_issue_active.ratio') and 'not_issued' not in h:
        try:
            st.append((float(vals[i]), h.split('issue_stalled_')[1].split('_per_issue')[0]))
        except ValueError:
            pass
print("stalls per issue:", ", ".join("%s %.2f" % (n, v) for v, n in sorted(st, reverse=True)[:7]))
