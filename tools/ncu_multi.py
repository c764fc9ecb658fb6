"""Summarise every kernel of an `ncu --page raw --csv` export.  usage: ncu_multi.py raw.csv"""
import csv
import sys

W = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
     'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
     'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
     'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
     'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
     'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
     'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
     'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct',
     'lts__t_sector_hit_rate.pct', 'smsp__inst_executed.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
     'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum']
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    print('-' * 80)
    print(r[idx['Kernel Name']][:100])
    for w in W:
        if w in idx:
            print(f"  {w:70s} {r[idx[w]]:>16s} {rows[1][idx[w]]}")
    st = []
    for h in hdr:
        if h.startswith('smsp__pcsamp_warps_issue_stalled_') and 'not_issued' not in h:
            st.append((float(r[idx[h]].replace(',', '') or 0), h.replace('smsp__pcsamp_warps_issue_stalled_', '')))
    st.sort(reverse=True)
    tot = sum(v for v, _ in st) or 1
    print('  stalls:', ', '.join(f"{h} {100 * v / tot:.0f}%" for v, h in st[:6]))
