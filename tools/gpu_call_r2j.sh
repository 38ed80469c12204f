#!/bin/bash
# Final round-2 measurement call: full GPU test suite, driver-style bench lines (headline, mie), one ncu --set full
# capture of the fused kernel (scratch discard on) and of the 512-point z pass (split exchange), launch list
mkdir -p gpurun_out; rm -f gpurun_out/j_*
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,memory.total,power.limit --format=csv > gpurun_out/j_gpu_info.txt 2>&1
( timeout 330 python -m pytest tests -m gpu -q > gpurun_out/j_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j_pytest_gpu.log )
tail -3 gpurun_out/j_pytest_gpu.log
timeout 240 python bench.py --steps 20 --warmup 5 > gpurun_out/j_bench_n1.json 2> gpurun_out/j_bench_n1.err
cut -c1-300 gpurun_out/j_bench_n1.json
timeout 150 ncu --set full --clock-control none --import-source on -k regex:k_shpf_fused --launch-skip 4 -c 1 -f -o gpurun_out/j_prof_fused \
    python tools/ab_lib.py --config headline --steps 2 --warmup 2 > gpurun_out/j_ncu_fused.log 2>&1
timeout 200 python bench.py --config mie --steps 20 --warmup 5 --no-cpu > gpurun_out/j_bench_mie_n1.json 2> gpurun_out/j_bench_mie_n1.err
cut -c1-300 gpurun_out/j_bench_mie_n1.json
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/j_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-parity > gpurun_out/j_bench_under_ncu.log 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:k_zline --launch-skip 2 -c 1 -f -o gpurun_out/j_prof_zline_mie \
    python tools/ab_lib.py --config mie --steps 2 --warmup 1 > gpurun_out/j_ncu_zmie.log 2>&1
timeout 100 python tools/ab_lib.py --config mie --steps 20 2>&1 | grep lib > gpurun_out/j_ab.log
timeout 100 python tools/ab_lib.py --config all256 --steps 40 2>&1 | grep lib >> gpurun_out/j_ab.log
cat gpurun_out/j_ab.log; ls -la gpurun_out | grep " j_"
