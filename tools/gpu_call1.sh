#!/bin/bash
# GPU call: validate + L2-residency experiments (development helper)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu_info.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log )
tail -3 gpurun_out/pytest_gpu.log
timeout 900 python tools/kexp.py --steps 10 --warmup 3 > gpurun_out/kexp.log 2>&1
cat gpurun_out/kexp.log | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    j = json.loads(l); print(j.get('name'), j.get('ms_per_step'), j.get('gcell_s'), j.get('bit_identical'), j.get('error',''))"
for v in chunk18 chunk18_hs chunk18_hs_zin chunk18_hs_g2 chunk18_win; do
  timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct \
     --clock-control none --cache-control none --csv --log-file gpurun_out/ncu_$v.csv \
     python tools/kexp.py --only $v --steps 1 --warmup 0 --check-steps 0 > gpurun_out/ncu_$v.log 2>&1
  python tools/ncu_sum.py gpurun_out/ncu_$v.csv
done
