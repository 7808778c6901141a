#!/bin/bash
TAG=${1:-r3g}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log; tail -3 gpurun_out/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_c2.json 2> gpurun_out/${TAG}_bench_c2.err
python - <<EOF
import json
d=json.loads(open("gpurun_out/${TAG}_bench_c2.json").read().strip().splitlines()[-1])
print("config 2:", round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],2), 'coll', round(d['config']['collection_ms'],2), 'learn', round(d['config']['learn_ms'],2), 'frac', round(d['roofline_ppo']['frac'],4), d['clocks'])
EOF
