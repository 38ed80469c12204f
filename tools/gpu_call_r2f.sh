#!/bin/bash
# A/B of the monotonic-counter fused launch against the previous build (ab/old.so) + parity
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3 > gpurun_out/f_tests.log
for i in 1 2 3; do
  IES_B200_LIB=$PWD/ab/old.so timeout 120 python tools/ab_lib.py --config headline --steps 40 >> gpurun_out/f_ab.log 2>&1
  timeout 120 python tools/ab_lib.py --config headline --steps 40 >> gpurun_out/f_ab.log 2>&1
done
IES_B200_LIB=$PWD/ab/old.so timeout 120 python tools/ab_lib.py --config all256 --steps 40 >> gpurun_out/f_ab.log 2>&1
timeout 120 python tools/ab_lib.py --config all256 --steps 40 >> gpurun_out/f_ab.log 2>&1
cat gpurun_out/f_tests.log gpurun_out/f_ab.log
