#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_ppo_gpu.py -q -x 2>&1 | tail -15 | tee gpurun_out/r2q_pytest.log
echo "pipe on"; timeout 200 python tools/prof_update.py 4096 64 3 2>&1 | grep "graph replay\|us per minibatch" | tee gpurun_out/r2q_time_update.log
echo "pipe off"; GRX_LAYER_PIPE=0 timeout 200 python tools/prof_update.py 4096 64 3 2>&1 | grep "graph replay" | tee -a gpurun_out/r2q_time_update.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r2q_bench_c2.json 2> gpurun_out/r2q_bench_c2.err; python - <<EOF
import json
d=json.loads(open("gpurun_out/r2q_bench_c2.json").read().strip().splitlines()[-1])
print("config 2:", round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],2), 'coll', round(d['config']['collection_ms'],2), 'learn', round(d['config']['learn_ms'],2))
EOF
