#!/bin/bash
# Multi-GPU bench of the contract command on N GPUs of one box: tools/gpu_scale.sh <tag> <N> [N ...]
tag=$1; shift
port=29530
for n in "$@"; do
  port=$((port+1))
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/${tag}_bench_${n}gpu.json 2> gpurun_out/${tag}_bench_${n}gpu.err
  python -c "
import json
d=json.loads(open('gpurun_out/${tag}_bench_${n}gpu.json').read().strip().split('\n')[-1])
ex=d.get('extra') or {}
print('$n GPUs', 'value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], '| full_gather', (ex.get('full_gather') or {}).get('evals_per_s'), (ex.get('full_gather') or {}).get('nvlink_ingress_GBps'), '| weak', (ex.get('weak_scaling') or {}).get('evals_per_s'))" || tail -5 gpurun_out/${tag}_bench_${n}gpu.err
done
