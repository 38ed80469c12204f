#!/bin/bash
# discard with keep range (all-face CPML), lead sweep with the discard on, launch list of the mie slab
mkdir -p gpurun_out; rm -f gpurun_out/h_*.log gpurun_out/h_*.csv
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3 > gpurun_out/h_tests.log
ab() { timeout 120 python tools/ab_lib.py "$@" 2>&1 | grep lib | sed "s/^/$1 $2 /" >> gpurun_out/h_ab.log; }
for i in 1 2; do
  ab --config all256 --steps 40 --opt fused_discard=0
  ab --config all256 --steps 40
done
ab --config headline --steps 40
ab --config headline --steps 40 --opt fused_lead=4
ab --config headline --steps 40 --opt fused_lead=8
ab --config headline --steps 40 --opt fused_lead=12
ab --config headline --steps 40 --opt fused_zb=1
ab --config headline --steps 40
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/h_mie_launches.csv python tools/ab_lib.py --config mie --steps 2 --warmup 1 > /dev/null 2>&1
cat gpurun_out/h_tests.log gpurun_out/h_ab.log
grep -v "^==" gpurun_out/h_mie_launches.csv | python -c "
import csv,sys
r=list(csv.reader(sys.stdin)); h=r[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); gi=h.index('Grid Size')
for x in r[-16:]: print(x[ki][:60], x[gi], x[vi])
"
