#!/bin/bash
# Round measurement call: parity tests, bench line, ncu launch list, ncu full captures
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,memory.total,power.limit --format=csv > gpurun_out/gpu_info.txt 2>&1
( timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log )
tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
cut -c1-400 gpurun_out/bench_n1.json
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
cut -c1-300 gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_yline_update --launch-skip 2 -c 1 -f -o gpurun_out/prof_yline \
    python tools/kexp.py --only base --steps 2 --warmup 0 --check-steps 0 > gpurun_out/ncu_full_y.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_zline --launch-skip 2 -c 1 -f -o gpurun_out/prof_zline \
    python tools/kexp.py --only base --steps 2 --warmup 0 --check-steps 0 > gpurun_out/ncu_full_z.log 2>&1
ls -la gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fdtd_vec -c 1 -f -o gpurun_out/prof_fdtd \
    python tools/bench_methods.py --only 0 --steps 1 --warmup 0 > gpurun_out/ncu_full_f.log 2>&1
