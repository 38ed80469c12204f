#!/bin/bash
# scaling bench lines only: tools/gpu_call_scale.sh N "headline mie"
N=${1:-2}; CFGS=${2:-headline}
mkdir -p gpurun_out
for cfg in $CFGS; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
      bench.py --gpus $N --steps 60 --warmup 10 --config $cfg > gpurun_out/bench_${cfg}_n$N.json 2> gpurun_out/bench_${cfg}_n$N.err
  python - <<PY
import json
try:
    j=json.loads([l for l in open('gpurun_out/bench_${cfg}_n$N.json') if l.startswith('{')][0])
    print('${cfg}', 'N=$N', 'value', round(j['value']), 'ms/step', round(j['ms_per_step'],3), 'parity', j['parity_check']['rel_l2_vs_oracle'], 'step_frac', round(j['roofline']['step_frac'],3), 'e2e', round(j['e2e']['value']), j['clocks'])
except Exception as e: print('no json', e)
PY
done
