"""Print selected metrics from `ncu -i X.ncu-rep --page raw --csv` output (stdin or file)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
want = sys.argv[2:] or ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
    'launch__occupancy_limit_registers', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'dram__cycles_active.avg.pct_of_peak_sustained_elapsed']
idx = [(w, hdr.index(w)) for w in want if w in hdr]
for r in rows[2:]:
    print('----')
    for w, i in idx:
        print(f"  {w:72s} {r[i]}  [{rows[1][i]}]")
