# ncu --set full of the thread-per-segment walk (one launch), key metrics + stall reasons exported on the box
B="python bench.py --steps 1 --warmup 1 --passes 1 --influence 0.5 --skip-count --skip-sw --skip-cpu --skip-gz --skip-frontend --skip-e2e"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:zscore_tseg_walk -s 2 -c 1 -o gpurun_out/prof_walk_r2j -f $B > gpurun_out/ncu_walkprof.log 2>&1
python profiles/ncu_key.py gpurun_out/prof_walk_r2j.ncu-rep > gpurun_out/ncu_key_metrics_walk_r2j.txt 2>&1
ncu -i gpurun_out/prof_walk_r2j.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
r=list(csv.reader(sys.stdin)); h,u,row=r[0],r[1],r[2]
for a,b,c in zip(h,u,row):
    if ('issue_stalled' in a and a.endswith('_per_warp_active.pct')) or 'pipe_fp64' in a or 'inst_executed_pipe_fp64' in a or a.startswith('smsp__average_warp') or 'achieved_occupancy' in a or 'warps_eligible' in a or a.startswith('launch__') or 'inst_executed_pipe' in a and 'pct' in a:
        print(a,b,c)
" >> gpurun_out/ncu_key_metrics_walk_r2j.txt
cat gpurun_out/ncu_key_metrics_walk_r2j.txt
