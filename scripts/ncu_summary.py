"""Print the key metrics of an `ncu --page raw --csv` dump (one block per captured kernel)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
h, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_blocks', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__warps_active.avg.per_cycle_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__cycles_active.avg', 'sm__cycles_elapsed.avg',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_lsu.sum', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.sum', 'lts__t_sectors_op_red.sum', 'lts__t_sectors_op_atom.sum']
for r in rows[2:]:
    print('-' * 100)
    for w in want:
        if w in h:
            print("%-72s %s %s" % (w, r[h.index(w)], units[h.index(w)]))
    st = []
    for i, name in enumerate(h):
        if 'issue_stalled' in name and name.endswith('_per_warp_active.pct'):
            try:
                st.append((float(r[i].replace(',', '')), name))
            except ValueError:
                pass
    for v, name in sorted(st, reverse=True)[:9]:
        print("   stall %-70s %6.1f" % (name.split('issue_stalled_')[1].replace('_per_warp_active.pct', ''), v))
