#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ppo_gpu.py tests/test_gemm_gpu.py tests/test_fullbody_gpu.py -q 2>&1 | tail -4 | tee gpurun_out/r2x_pytest.log
GRX_PROF_DIMS=105,234,32 timeout 300 python tools/prof_update.py 4096 64 3 2>&1 | grep "graph replay\|us per minibatch" | cut -c1-120 | tee gpurun_out/r2x_time_update_fullbody.log
timeout 300 python bench.py --steps 5 --warmup 3 --config 6 > gpurun_out/r2x_bench_c6.json 2> gpurun_out/r2x_bench_c6.err; python - <<EOF
import json
d=json.loads(open("gpurun_out/r2x_bench_c6.json").read().strip().splitlines()[-1])
print("config 6:", round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],2), 'coll', round(d['config']['collection_ms'],2), 'learn', round(d['config']['learn_ms'],2), 'env_us', round(d['roofline_env']['us_per_launch'],1))
EOF
