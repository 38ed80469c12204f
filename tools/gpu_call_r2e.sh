#!/bin/bash
# Round 2, call E: the numbers and ncu evidence kept under profiles/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,memory.total,power.limit --format=csv > gpurun_out/gpu_info.txt 2>&1
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
cut -c1-300 gpurun_out/bench_n1.json
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
cut -c1-400 gpurun_out/bench_ref.json
timeout 900 python bench.py --config mie --steps 40 --warmup 10 > gpurun_out/bench_mie_n1.json 2> gpurun_out/bench_mie_n1.err
cut -c1-300 gpurun_out/bench_mie_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_shpf_fused --launch-skip 2 -c 1 -f -o gpurun_out/prof_fused \
    python tools/kexp.py --only fused --steps 2 --warmup 0 --check-steps 0 > gpurun_out/ncu_full_fused.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_yline_update --launch-skip 2 -c 1 -f -o gpurun_out/prof_yline \
    python tools/kexp.py --only base --steps 2 --warmup 0 --check-steps 0 > gpurun_out/ncu_full_y.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_zline --launch-skip 2 -c 1 -f -o gpurun_out/prof_zline \
    python tools/kexp.py --only base --steps 2 --warmup 0 --check-steps 0 > gpurun_out/ncu_full_z.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_yline_update --launch-skip 2 -c 1 -f -o gpurun_out/prof_mie_y \
    python tools/kexp.py --config mie --only base --steps 2 --warmup 0 --check-steps 0 > gpurun_out/ncu_full_mie_y.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_zline --launch-skip 2 -c 1 -f -o gpurun_out/prof_mie_z \
    python tools/kexp.py --config mie --only base --steps 2 --warmup 0 --check-steps 0 > gpurun_out/ncu_full_mie_z.log 2>&1
# other methods / dtypes: one full capture of the dominant kernel each (indices into tools/bench_methods.py CASES)
timeout 600 ncu --set full --clock-control none -k regex:k_fdtd_vec -c 1 -f -o gpurun_out/prof_fdtd python tools/bench_methods.py --only 0 --steps 1 --warmup 0 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none -k regex:k_yline_update --launch-skip 1 -c 1 -f -o gpurun_out/prof_f32_y python tools/bench_methods.py --only 5 --steps 1 --warmup 1 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none -k regex:k_shpf_fused --launch-skip 1 -c 1 -f -o gpurun_out/prof_f32_fused python tools/bench_methods.py --only 5 --steps 1 --warmup 1 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none -k regex:k_yline_update --launch-skip 1 -c 1 -f -o gpurun_out/prof_c128_y python tools/bench_methods.py --only 7 --steps 1 --warmup 1 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none -k regex:k_yline_update --launch-skip 1 -c 1 -f -o gpurun_out/prof_allpml_y python tools/bench_methods.py --only 10 --steps 1 --warmup 1 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none -k regex:k_yline_update --launch-skip 1 -c 1 -f -o gpurun_out/prof_pstd_y python tools/bench_methods.py --only 14 --steps 1 --warmup 1 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none -k regex:k_xline --launch-skip 1 -c 1 -f -o gpurun_out/prof_pstd_x python tools/bench_methods.py --only 14 --steps 1 --warmup 1 > /dev/null 2>&1
timeout 900 python tools/bench_methods.py > gpurun_out/methods.log 2>&1
cut -c1-200 gpurun_out/methods.log | grep -v "FFT kernel"
ls -la gpurun_out/*.ncu-rep
