#!/bin/bash
# Round-2 measurement call (1 GPU): parity tests, bench of the three BASELINE configs, ncu launch list of the bench command (kernel shares),
# ncu --set full captures of the env kernel and of one minibatch's kernels, text summaries + traffic.json for profiles/.
TAG=${1:-r2g}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
for c in 2 3 5; do
  timeout 600 python bench.py --steps 5 --warmup 3 --config $c > gpurun_out/${TAG}_bench_c$c.json 2> gpurun_out/${TAG}_bench_c$c.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/${TAG}_bench_c$c.json').read().strip().splitlines()[-1])
print('config $c value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],2), 'coll', round(d['config']['collection_ms'],2), 'learn', round(d['config']['learn_ms'],2), 'cpu', round(d['cpu_baseline']['value']), 'launches', d['gpu_launches'])"
done
timeout 300 python tools/time_env.py 4096 200 | tee gpurun_out/${TAG}_time_env.log
timeout 300 python tools/prof_update.py 4096 64 3 | tee gpurun_out/${TAG}_time_update.log
# launch list of the bench command: skip the first learn(1) + one warm-up iteration (one iteration = 64 x 7 + 6 + 200 x 10 + ... launches)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 5200 -c 2500 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/${TAG}_ncu_bench.log 2>&1
python tools/summarize_launches.py gpurun_out/${TAG}_launches.csv gpurun_out/${TAG}_launches_summary.csv | head -18
timeout 600 ncu --set full --clock-control none --import-source on -k regex:env_step_kernel -s 30 -c 1 -f -o gpurun_out/${TAG}_env python tools/time_env.py 4096 40 > gpurun_out/${TAG}_ncu_env.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gemm_tf32|ppo_heads|apply_kernel|gather' -s 195 -c 22 -f -o gpurun_out/${TAG}_upd python tools/prof_update.py 4096 64 2 > gpurun_out/${TAG}_ncu_upd.log 2>&1
ls -la gpurun_out | tail -8
