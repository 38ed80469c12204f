#!/bin/bash
# A/B of the previous build (ab/base.so) against: z role loading its first tile before the table staging barrier,
# conflict-free stash columns, 4 z tiles per CTA (option); then the full GPU suite on the new build
mkdir -p gpurun_out; rm -f gpurun_out/m_*
ab() { timeout 100 python tools/ab_lib.py "$@" 2>&1 | grep '"lib"' >> gpurun_out/m_ab.log; }
IES_B200_LIB=$PWD/ab/base.so ab --config headline --steps 40
ab --config headline --steps 40
IES_B200_LIB=$PWD/ab/base.so ab --config headline --steps 40
ab --config headline --steps 40
ab --config headline --steps 40 --opt fused_zb=4
ab --config headline --steps 40 --opt fused_zb=4 --opt fused_lead=8
cat gpurun_out/m_ab.log
( timeout 200 python -m pytest tests -m gpu -q -x > gpurun_out/m_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/m_pytest_gpu.log )
tail -3 gpurun_out/m_pytest_gpu.log
