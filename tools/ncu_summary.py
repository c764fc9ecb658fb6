"""Print the metrics that matter from an `ncu --page raw --csv` export.  usage: ncu_summary.py raw.csv"""
import csv
import sys

WANT = ['Kernel Name', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__waves_per_multiprocessor', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum']
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
stall = [h for h in hdr if 'warp_issue_stalled' in h and h.endswith('per_warp_active.pct')]
for r in rows[2:]:
    print('-' * 100)
    for w in WANT:
        if w in idx:
            print(f"{w:72s} {r[idx[w]]:>18s} {rows[1][idx[w]]}")
    st = sorted(((float(r[idx[h]].replace(',', '') or 0), h) for h in stall), reverse=True)[:6]
    for v, h in st:
        print(f"   stall {h.replace('smsp__warp_issue_stalled_', '').replace('_per_warp_active.pct', ''):40s} {v:8.1f}")
