#!/bin/bash
# 2-GPU round: the multi-GPU parity test + bench at N = 2 (launched like the driver does), for the all-reduce protocols 1 and 2 (GRX_COMM_ONESHOT)
TAG=${1:-r3f}
mkdir -p gpurun_out
for mode in 2 1; do
  GRX_COMM_ONESHOT=$mode timeout 600 python -m pytest tests/test_multigpu.py -q -s > gpurun_out/${TAG}_multigpu_pytest_mode$mode.log 2>&1; grep -E "2-GPU update|passed|failed" gpurun_out/${TAG}_multigpu_pytest_mode$mode.log
  GRX_COMM_ONESHOT=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$mode bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_2gpu_mode$mode.json 2> gpurun_out/${TAG}_bench_2gpu_mode$mode.err
  python - <<EOF
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_2gpu_mode$mode.json").read().strip().splitlines()[-1])
    print("mode $mode N=2:", round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],2), 'coll', round(d['config']['collection_ms'],2), 'learn', round(d['config']['learn_ms'],2), 'comm_error', d['comm_error'], 'identical', d['replicas_identical'])
except Exception as e:
    print("mode $mode FAILED", e); print(open("gpurun_out/${TAG}_bench_2gpu_mode$mode.err").read()[-1500:])
EOF
done
