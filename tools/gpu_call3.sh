#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log )
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | head -40
timeout 600 python tools/kexp.py --only alt --steps 20 --warmup 3 > gpurun_out/kexp.log 2>&1
grep -E '^\{|max abs' gpurun_out/kexp.log | cut -c1-600
