#!/bin/bash
# key counters + per-region instruction counts of an ncu report: tools/prof_summary.sh REPORT LIB
ncu -i $1 --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
hdr=rows[0]
r=rows[2]
out=[]
for i,h in enumerate(hdr):
    if ('smsp__average_warp' in h and 'issue_stalled' in h and 'not_issued' not in h):
        out.append((float(r[i]),h))
print('stalls/issue:', ', '.join('%s %.2f'%(h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''),v) for v,h in sorted(out,reverse=True)[:8]))
for h in ('gpu__time_duration.sum','smsp__inst_executed.sum','sm__icc_request_hit_rate.pct','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','smsp__thread_inst_executed_per_inst_executed.ratio','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','dram__throughput.avg.pct_of_peak_sustained_elapsed','dram__bytes_read.sum','dram__bytes_write.sum'):
    if h in hdr: print('%-70s %s'%(h, r[hdr.index(h)]))
"
