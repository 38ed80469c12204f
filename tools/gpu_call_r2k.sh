#!/bin/bash
# Final build: the default (20 + 100 step) bench line and the methods table
mkdir -p gpurun_out; rm -f gpurun_out/k_*
timeout 200 python bench.py > gpurun_out/k_bench_n1_default.json 2> gpurun_out/k_bench_n1_default.err
cut -c1-260 gpurun_out/k_bench_n1_default.json
timeout 260 python tools/bench_methods.py > gpurun_out/k_methods.log 2>&1; cp gpurun_out/methods.json gpurun_out/k_methods.json
grep -c method gpurun_out/k_methods.json
