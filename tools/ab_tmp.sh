run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 --steps 20 --warmup 5 $3 2>gpurun_out/ab8_$2.err > gpurun_out/ab8_$2.json; python -c "
import json,sys
d=json.loads(open('gpurun_out/ab8_$2.json').read().strip().split('\n')[-1]); print('$2', d['value'], d['ms_per_step'], d['gather']['nvlink_ingress_GBps'], (d.get('extra') or {}).get('costs_only_gather',{}).get('evals_per_s'), (d.get('extra') or {}).get('weak_scaling',{}).get('evals_per_s'), d['e2e']['value'])"; tail -2 gpurun_out/ab8_$2.err; }
run 29511 fused "--gather fused"
run 29512 nccl "--gather nccl --no-extra"
