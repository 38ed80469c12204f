#!/bin/bash
# Round 2, call A: new parity cases + fused-kernel experiment sweep + DRAM bytes of the variants
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,memory.total,power.limit --format=csv > gpurun_out/gpu_info.txt 2>&1
( timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multi.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log )
tail -15 gpurun_out/pytest_gpu.log
timeout 600 python tools/kexp.py --steps 20 --warmup 5 > gpurun_out/kexp.log 2>&1
cat gpurun_out/kexp.log | cut -c1-200
for v in base fused_full_l3 fused_r32_l3; do
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none \
     -k regex:'k_zline|k_yline_update|k_shpf_fused' --launch-skip 4 -c 4 --csv --log-file gpurun_out/dram_$v.csv \
     python tools/kexp.py --only $v --steps 3 --warmup 0 --check-steps 0 > gpurun_out/dram_$v.log 2>&1
done
grep -h "k_" gpurun_out/dram_*.csv | cut -c1-400 | tail -40
