"""Print the key metrics of every kernel in an .ncu-rep (ncu --set full capture)."""
import csv, subprocess, sys
KEYS = [
 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
 'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
 'smsp__inst_executed.sum', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
 'smsp__issue_active.avg.pct_of_peak_sustained_active',
 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
 'smsp__inst_executed_op_shared_ld.sum', 'smsp__inst_executed_op_shared_st.sum',
 'smsp__inst_executed_op_global_ld.sum', 'smsp__inst_executed_op_global_st.sum',
 'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum',
]
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print('==', d['Kernel Name'][:70], d['Block Size'], d['Grid Size'])
    for k in KEYS:
        if k in d:
            print('   %-70s %-10s %s' % (k, units[hdr.index(k)], d[k]))
    st = [(float(d[k]), k) for k in hdr
          if k.startswith('smsp__average_warps_issue_stalled') and k.endswith('per_issue_active.ratio') and d[k]]
    for v, k in sorted(st, reverse=True)[:6]:
        print('   stall %-50s %.2f' % (k[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')], v))
