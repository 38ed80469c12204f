#!/bin/bash
# GPU call: parity tests, alt-vs-two-kernel timing, bench line, ncu launch list + full capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu_info.txt 2>&1
( timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log )
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python tools/kexp.py --only alt --steps 20 --warmup 3 > gpurun_out/kexp.log 2>&1
grep '^{' gpurun_out/kexp.log
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
cat gpurun_out/bench_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_shpf_half -c 2 -f -o gpurun_out/prof_half \
    python tools/kexp.py --only alt --steps 1 --warmup 0 --check-steps 0 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
