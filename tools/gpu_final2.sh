#!/bin/bash
TAG=${1:-r3j}
mkdir -p gpurun_out
for c in 3 5 6; do
  timeout 200 python bench.py --steps 10 --warmup 3 --config $c > gpurun_out/${TAG}_bench_c$c.json 2> gpurun_out/${TAG}_bench_c$c.err
  python - <<EOF
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_c$c.json").read().strip().splitlines()[-1])
    print("config $c:", round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],2), 'coll', round(d['config']['collection_ms'],2), 'learn', round(d['config']['learn_ms'],2), 'env_us', round(d['roofline_env']['us_per_launch'],1), d['clocks']['reasons'])
except Exception as e:
    print("config $c FAILED", e)
EOF
done
