#!/bin/bash
# Round 2, call C: full gpu test suite + bench lines (headline, mie) on one GPU
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log )
tail -8 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 50 --warmup 10 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -3 gpurun_out/bench_n1.err; cut -c1-1500 gpurun_out/bench_n1.json
IES_B200_FUSED=0 timeout 900 python bench.py --steps 50 --warmup 10 --no-cpu > gpurun_out/bench_n1_twokernel.json 2> gpurun_out/bench_n1_twokernel.err
cut -c1-300 gpurun_out/bench_n1_twokernel.json
timeout 900 python bench.py --config mie --steps 30 --warmup 5 > gpurun_out/bench_mie_n1.json 2> gpurun_out/bench_mie_n1.err
tail -3 gpurun_out/bench_mie_n1.err; cut -c1-1500 gpurun_out/bench_mie_n1.json
IES_B200_FUSED=0 timeout 900 python bench.py --config mie --steps 30 --warmup 5 --no-cpu --no-parity > gpurun_out/bench_mie_n1_twokernel.json 2> gpurun_out/bench_mie_n1_twokernel.err
cut -c1-300 gpurun_out/bench_mie_n1_twokernel.json
