#!/bin/bash
# Final build on 2 GPUs: driver-style launch of the headline and the Mie workload (N-rank parity_check inside)
mkdir -p gpurun_out; rm -f gpurun_out/l_*
for cfg in headline mie; do
  timeout 110 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
      bench.py --gpus 2 --steps 20 --warmup 5 --config $cfg > gpurun_out/l_bench_${cfg}_n2.json 2> gpurun_out/l_bench_${cfg}_n2.err
  echo "rc=$?"; cut -c1-200 gpurun_out/l_bench_${cfg}_n2.json
done
