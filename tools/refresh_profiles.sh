#!/bin/bash
# turns gpurun_out/final_r2 (written by tools/final_r2.sh on the GPU box) into the tracked summaries under profiles/
set -e
R=gpurun_out/final_r2
M="gpu__time_duration.sum,launch__registers_per_thread,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,smsp__thread_inst_executed_per_inst_executed.ratio,sm__icc_request_hit_rate.pct,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"
python tools/make_traffic.py $R/prof.ncu-rep b
ncu -i $R/prof.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
want='$M'.split(',')
rows=list(csv.reader(sys.stdin)); hdr=rows[0]
idx=[hdr.index('Kernel Name')]+[hdr.index(m) for m in want]
w=csv.writer(sys.stdout)
for k,r in enumerate(rows): w.writerow([r[i][:52] if (k>1 and i==idx[0]) else r[i] for i in idx])
" > profiles/r2_ncu_summary.csv
bash tools/prof_summary.sh $R/prof.ncu-rep > profiles/r2_final_ncu_k_fast_step.txt
python tools/ncu_lines.py $R/prof.ncu-rep k_fast_step --sym k_fast_stepILb0ELi1ELb0 --top 45 > profiles/r2_final_ncu_lines.txt 2>&1 || true
python - <<'PY'
import csv,collections,re
rows=list(csv.reader(l for l in open('gpurun_out/final_r2/launches.csv') if l.startswith('"')))
hdr=rows[0]; kn=hdr.index('Kernel Name'); mv=hdr.index('Metric Value'); mu=hdr.index('Metric Unit')
agg=collections.OrderedDict()
for r in rows[1:]:
    v=float(r[mv].replace(',','')); u=r[mu]
    ms = v/1e6 if u in('ns','nsecond') else v/1e3 if u in ('us','usecond') else v
    name=re.sub(r'\(.*','',r[kn])[:60]
    a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=ms
tot=sum(a[1] for a in agg.values())
with open('profiles/r2_launch_shares.csv','w') as f:
    f.write('# ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 python bench.py --steps 12 --warmup 6 (config B, cold-cache serialised launches: compare shares)\n')
    f.write('kernel,launches,total_ms,share_pct,avg_ms\n')
    for n,(c,t) in sorted(agg.items(), key=lambda kv:-kv[1][1]):
        f.write('"%s",%d,%.4f,%.2f,%.5f\n'%(n,c,t,100*t/tot,t/c))
PY
cp $R/launches.csv profiles/r2_final_launches.csv; cp $R/launches.csv profiles/r2_launches.csv
for f in bench_n1 bench_ref bench_c_n1 bench_e_n1 bench_n1_halo2; do [ -s $R/$f.json ] && cp $R/$f.json profiles/r2_final_$f.json; done
[ -f $R/pytest_gpu.log ] && tail -3 $R/pytest_gpu.log > profiles/r2_final_pytest_gpu.log
[ -f $R/pytest_gpu_rest.log ] && tail -3 $R/pytest_gpu_rest.log > profiles/r2_final_pytest_gpu_rest.log
cp $R/smoke.log profiles/r2_final_smoke.log
# SASS excerpt of the hot kernel
{ echo "# cuobjdump -sass starfish_b200/libstarfish_gpu.so, k_fast_step<false, 1> (sm_100a), $(git rev-parse --short HEAD)+"; cuobjdump -sass starfish_b200/libstarfish_gpu.so | awk '/Function : _Z11k_fast_stepILb0ELi1ELb0/{p=1} p&&/Function : /&&!/k_fast_stepILb0ELi1ELb0/{p=0} p' > /tmp/kfs.sass; n=$(grep -cE '^\s+/\*[0-9a-f]{4}\*/' /tmp/kfs.sass); echo "# instruction count: $n; mnemonic histogram:"; grep -E '^\s+/\*[0-9a-f]{4}\*/' /tmp/kfs.sass | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+(@!?U?P[0-9T]+\s+)?([A-Z0-9_.]+).*/\2/' | sort | uniq -c | sort -rn | head -40; echo "# row walk (LDS.128 operands, DFMA chain, conditional flush):"; grep -n "LDS.128" /tmp/kfs.sass | head -3; L=$(grep -n "LDS.128" /tmp/kfs.sass | head -1 | cut -d: -f1); sed -n "$((L-5)),$((L+70))p" /tmp/kfs.sass; } > profiles/r2_k_fast_step_sass.txt
cat profiles/r2_launch_shares.csv; cat profiles/r2_final_ncu_k_fast_step.txt; head -12 profiles/r2_final_ncu_lines.txt; cat profiles/traffic.json
