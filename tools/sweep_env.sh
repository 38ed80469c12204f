#!/bin/bash
run() { echo -n "$1: "; env $1 timeout 300 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print(round(j['value']), round(j['ms_per_step'],3), round(j['roofline']['avg_launch_ms'],3), round(j['roofline']['zline_avg_launch_ms'],3))"; }
run "IES_B200_TABLES_SMEM=1"
run "IES_B200_TABLES_SMEM=0"
