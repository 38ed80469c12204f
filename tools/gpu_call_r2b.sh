#!/bin/bash
# Round 2, call B: DRAM bytes / stalls of the fused kernel
mkdir -p gpurun_out
for v in fused_full_l3 fused_r32_l3; do
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu.sum,smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct,smsp__warp_issue_stalled_barrier_per_warp_active.pct,smsp__warp_issue_stalled_membar_per_warp_active.pct,smsp__warp_issue_stalled_sleeping_per_warp_active.pct,smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct,smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct,smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct \
     --clock-control none -k regex:'k_shpf_fused' --launch-skip 2 -c 2 --csv --log-file gpurun_out/dram_$v.csv \
     python tools/kexp.py --only $v --steps 3 --warmup 0 --check-steps 0 > gpurun_out/dram_$v.log 2>&1
done
grep -h "k_shpf" gpurun_out/dram_*.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | tail -60
