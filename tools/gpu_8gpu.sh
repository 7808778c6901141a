#!/bin/bash
# 8-GPU box: bench at N = 8 / 4 for the all-reduce protocols (GRX_COMM_ONESHOT = 2: apply_kernel pulls everything; 1: phase-0 kernel + phase 1 pulled), launched like the driver does
TAG=${1:-r3h}
mkdir -p gpurun_out
run() { # n mode port
  GRX_COMM_ONESHOT=$2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $3 bench.py --gpus $1 --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_$1gpu_mode$2.json 2> gpurun_out/${TAG}_bench_$1gpu_mode$2.err
  python - <<EOF
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_$1gpu_mode$2.json").read().strip().splitlines()[-1])
    print("N=$1 mode $2:", round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],2), 'coll', round(d['config']['collection_ms'],2), 'learn', round(d['config']['learn_ms'],2), 'comm_error', d['comm_error'], 'identical', d['replicas_identical'])
except Exception as e:
    print("N=$1 mode $2: FAILED", e); print(open("gpurun_out/${TAG}_bench_$1gpu_mode$2.err").read()[-1500:])
EOF
}
run 8 2 29531
run 4 2 29532
run 4 1 29533
