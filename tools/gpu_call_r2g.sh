#!/bin/bash
# fused kernel: discard of consumed scratch lines -- parity, same-box A/B (option off/on, previous build), DRAM bytes
mkdir -p gpurun_out; rm -f gpurun_out/g_*.log
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3 > gpurun_out/g_tests.log
for i in 1 2 3; do
  timeout 120 python tools/ab_lib.py --config headline --steps 40 --opt fused_discard=0 2>&1 | grep lib >> gpurun_out/g_ab.log
  timeout 120 python tools/ab_lib.py --config headline --steps 40 2>&1 | grep lib >> gpurun_out/g_ab.log
done
IES_B200_LIB=$PWD/ab/old.so timeout 120 python tools/ab_lib.py --config headline --steps 40 2>&1 | grep lib >> gpurun_out/g_ab.log
IES_B200_LIB=$PWD/ab/old.so timeout 120 python tools/ab_lib.py --config mie --steps 20 2>&1 | grep lib >> gpurun_out/g_ab.log
timeout 120 python tools/ab_lib.py --config mie --steps 20 2>&1 | grep lib >> gpurun_out/g_ab.log
timeout 120 python tools/ab_lib.py --config all256 --steps 40 2>&1 | grep lib >> gpurun_out/g_ab.log
timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_shpf_fused -c 2 --csv --log-file gpurun_out/g_ncu.csv python tools/ab_lib.py --config headline --steps 2 --warmup 2 > /dev/null 2>&1
cat gpurun_out/g_tests.log gpurun_out/g_ab.log; grep -v "^==" gpurun_out/g_ncu.csv | cut -d, -f5,13- | tail -8
