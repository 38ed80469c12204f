#!/bin/bash
# Round 2, call D (N GPUs): multi-process IPC halo parity + weak/strong scaling bench lines
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
( timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/pytest_multi_n$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi_n$N.log )
tail -12 gpurun_out/pytest_multi_n$N.log
for cfg in headline strong mie; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
      bench.py --gpus $N --steps 40 --warmup 10 --config $cfg > gpurun_out/bench_${cfg}_n$N.json 2> gpurun_out/bench_${cfg}_n$N.err
  echo "rc=$?"; tail -2 gpurun_out/bench_${cfg}_n$N.err | cut -c1-300; cut -c1-260 gpurun_out/bench_${cfg}_n$N.json
  python - <<PY
import json
try:
    j=json.load(open('gpurun_out/bench_${cfg}_n$N.json'))
    print('${cfg}', 'N=$N', 'value', round(j['value']), 'ms/step', round(j['ms_per_step'],3), 'parity', j['parity_check'], 'step_frac', round(j['roofline']['step_frac'],3))
except Exception as e: print('no json', e)
PY
done
