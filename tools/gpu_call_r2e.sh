#!/bin/bash
# Round 2, call E: the numbers and ncu evidence kept under profiles/.  gpurun merges at most 64 MiB
# of gpurun_out/ back, so every full-set report is summarised ON the box (tools/ncu_summary.py) and
# only the two reports with imported source (headline fused kernel, Mie y-line kernel) travel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,memory.total,power.limit --format=csv > gpurun_out/gpu_info.txt 2>&1
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
cut -c1-300 gpurun_out/bench_n1.json
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
cut -c1-300 gpurun_out/bench_ref.json
timeout 900 python bench.py --config mie --steps 40 --warmup 10 > gpurun_out/bench_mie_n1.json 2> gpurun_out/bench_mie_n1.err
cut -c1-300 gpurun_out/bench_mie_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
ARGS=""
cap() {   # cap <name> <suffix> <kernel regex> <skip> <source on|off> -- <command...>
  local name=$1 suffix=$2 rx=$3 skip=$4 src=$5; shift 6
  timeout 900 ncu --set full --clock-control none --import-source $src -k regex:$rx --launch-skip $skip -c 1 -f -o /tmp/$name "$@" > gpurun_out/ncu_$name.log 2>&1
  ARGS="$ARGS /tmp/$name.ncu-rep:$suffix"
}
cap prof_fused "" k_shpf_fused 2 on -- python tools/kexp.py --only fused --steps 2 --warmup 0 --check-steps 0
cap prof_yline "" k_yline_update 2 off -- python tools/kexp.py --only base --steps 2 --warmup 0 --check-steps 0
cap prof_zline "" k_zline 2 off -- python tools/kexp.py --only base --steps 2 --warmup 0 --check-steps 0
cap prof_mie_y _mie k_yline_update 2 on -- python tools/kexp.py --config mie --only base --steps 2 --warmup 0 --check-steps 0
cap prof_mie_z _mie k_zline 2 off -- python tools/kexp.py --config mie --only base --steps 2 --warmup 0 --check-steps 0
cap prof_fdtd _f64 k_fdtd_vec 0 off -- python tools/bench_methods.py --only 0 --steps 1 --warmup 0
cap prof_f32_y _f32 k_yline_update 1 off -- python tools/bench_methods.py --only 5 --steps 1 --warmup 1
cap prof_c128_y _c128 k_yline_update 1 off -- python tools/bench_methods.py --only 7 --steps 1 --warmup 1
cap prof_allpml_y _allpml256 k_yline_update 1 off -- python tools/bench_methods.py --only 10 --steps 1 --warmup 1
cap prof_pstd_y _pstd k_yline_update 1 off -- python tools/bench_methods.py --only 14 --steps 1 --warmup 1
cap prof_pstd_x _pstd k_xline 1 off -- python tools/bench_methods.py --only 14 --steps 1 --warmup 1
python tools/ncu_summary.py $ARGS > gpurun_out/ncu_summary.json 2> gpurun_out/ncu_summary.err
cp /tmp/prof_fused.ncu-rep /tmp/prof_mie_y.ncu-rep gpurun_out/
timeout 900 python tools/bench_methods.py > gpurun_out/methods.log 2>&1
cut -c1-130 gpurun_out/methods.log | grep -v "FFT kernel"
ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out; head -c 600 gpurun_out/ncu_summary.json
