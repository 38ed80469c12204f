#!/bin/bash
# 512-point lines: fused kernel with the split exchange (+ pair discard), k_zline with the split exchange
mkdir -p gpurun_out; rm -f gpurun_out/i_*.log
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3 > gpurun_out/i_tests.log
IES_B200_LIB=$PWD/ab/zsplit.so timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k 512 2>&1 | tail -3 >> gpurun_out/i_tests.log
ab() { timeout 120 python tools/ab_lib.py "$@" 2>&1 | grep lib | sed "s/^/$1 $2 /" >> gpurun_out/i_ab.log; }
for i in 1 2; do
  ab --config mie --steps 20
  ab --config mie --steps 20 --opt fused=1 --opt fused_discard=0
  ab --config mie --steps 20 --opt fused=1
  IES_B200_LIB=$PWD/ab/zsplit.so ab --config mie --steps 20
done
ab --config mie --steps 20 --opt fused=1 --opt fused_zb=1
ab --config mie --steps 20 --opt fused=1 --opt fused_lead=3
ab --config mie --steps 20 --opt fused=1 --opt fused_lead=10
ab --config x512 --steps 20
ab --config x512 --steps 20 --opt fused=1
cat gpurun_out/i_tests.log gpurun_out/i_ab.log
