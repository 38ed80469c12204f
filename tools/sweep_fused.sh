#!/bin/bash
# sweep fused-kernel chunking parameters (development helper)
for cx in ${CXS:-4 8 16}; do for la in ${LAS:-1 2 3}; do
  echo -n "cx=$cx la=$la: "
  IES_B200_FUSED_CX=$cx IES_B200_FUSED_LA=$la timeout 300 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print(round(j['value']), round(j['ms_per_step'],3))"
done; done
