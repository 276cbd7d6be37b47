"""Print the headline counters of an `ncu --page raw --csv` dump (one column per profiled launch)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct', 'sm__inst_executed_pipe_alu.avg.pct', 'sm__inst_executed_pipe_xu.avg.pct', 'sm__inst_executed_pipe_lsu.avg.pct',
        'l1tex__t_sector_hit_rate.pct', 'smsp__inst_executed.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__occupancy_limit', 'sm__maximum_warps_per_active_cycle_pct', 'smsp__warps_eligible.avg.per_cycle_active', 'l1tex__data_pipe_lsu_wavefronts.sum ',
        'local_load', 'local_store', 'smsp__inst_executed_op_local', 'launch__grid_size', 'launch__block_size']
extra = sys.argv[2:]
for i, h in enumerate(hdr):
    if any(h == k or h.startswith(k) for k in keys) or any(e in h for e in extra):
        if h.endswith('.pct_of_peak_sustained_elapsed') and 'throughput' not in h: continue
        print("%-75s %-10s %s" % (h, units[i], [r[i] for r in rows[2:]]))
