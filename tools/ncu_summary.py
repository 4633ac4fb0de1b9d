"""Key metrics of every kernel in an .ncu-rep (raw page)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__grid_size', 'launch__block_size',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__cycles_active.avg', 'smsp__warps_eligible.avg.per_cycle_active']
idx = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    print('-----', r[idx['Kernel Name']][:90])
    for w in want:
        if w in idx:
            print(f"  {w:70s} {r[idx[w]]:>18s} {units[idx[w]]}")
