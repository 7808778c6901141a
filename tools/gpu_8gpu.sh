#!/bin/bash
# 8-GPU box: the 2-GPU parity test, then bench at N = 8 / 4 (config 5) / 2, launched like the driver does
TAG=${1:-r2n}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu.py -q -s > gpurun_out/${TAG}_multigpu_pytest.log 2>&1; grep -E "2-GPU update|passed|failed" gpurun_out/${TAG}_multigpu_pytest.log
run() { # n config port
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $3 bench.py --gpus $1 --steps 10 --warmup 3 --config $2 > gpurun_out/${TAG}_bench_$1gpu_c$2.json 2> gpurun_out/${TAG}_bench_$1gpu_c$2.err
  python - <<EOF
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_$1gpu_c$2.json").read().strip().splitlines()[-1])
    print("N=$1 config $2:", round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],2), 'coll', round(d['config']['collection_ms'],2), 'learn', round(d['config']['learn_ms'],2), 'comm_error', d['comm_error'], 'identical', d['replicas_identical'])
except Exception as e:
    print("N=$1 config $2: FAILED", e); print(open("gpurun_out/${TAG}_bench_$1gpu_c$2.err").read()[-1500:])
EOF
}
run 8 2 29521
run 4 5 29522
run 2 2 29523
run 4 2 29524
