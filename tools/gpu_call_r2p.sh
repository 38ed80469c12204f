#!/bin/bash
# A/B: y role looks at its plane's z counter before the FFT phase (skips the poll round trip), then the GPU suite
mkdir -p gpurun_out; rm -f gpurun_out/p_*
ab() { timeout 40 python tools/ab_lib.py "$@" 2>&1 | grep '"lib"' >> gpurun_out/p_ab.log; }
IES_B200_LIB=$PWD/ab/base.so ab --config headline --steps 40
ab --config headline --steps 40
IES_B200_LIB=$PWD/ab/base.so ab --config headline --steps 40
ab --config headline --steps 40
cat gpurun_out/p_ab.log
( timeout 70 python -m pytest tests -m gpu -q -x > gpurun_out/p_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/p_pytest_gpu.log )
tail -3 gpurun_out/p_pytest_gpu.log
