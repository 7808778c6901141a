#!/bin/bash
# Final round-2 measurement pass on one B200: GPU test suite, bench lines of every config, tool timings, ncu launch list of the bench command,
# ncu --set full of the env kernel and of the update's kernels, summaries + profiles/traffic.json inputs.
TAG=${1:-r2w}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/${TAG}_smoke.log
for c in 2 3 5 6; do
  timeout 600 python bench.py --steps 20 --warmup 5 --config $c > gpurun_out/${TAG}_bench_c$c.json 2> gpurun_out/${TAG}_bench_c$c.err
  python - <<EOF
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_c$c.json").read().strip().splitlines()[-1])
    print("config $c:", round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],2), 'coll', round(d['config']['collection_ms'],2), 'learn', round(d['config']['learn_ms'],2), 'env_us', round(d['roofline_env']['us_per_launch'],1), 'frac', round(d['roofline_ppo']['frac'],4), 'cpu', round(d['cpu_baseline']['value']), d['clocks'])
except Exception as e:
    print("config $c FAILED", e)
EOF
done
timeout 300 python bench.py --impl reference --steps 8 --warmup 1 > gpurun_out/${TAG}_bench_reference_c2.json 2>/dev/null; tail -c 400 gpurun_out/${TAG}_bench_reference_c2.json
timeout 300 python tools/time_env.py 4096 200 | tee gpurun_out/${TAG}_time_env.log
GRX_ENV_GENERIC=1 timeout 300 python tools/time_env.py 4096 60 | tee -a gpurun_out/${TAG}_time_env.log
timeout 300 python tools/prof_update.py 4096 64 3 | tail -3 | tee gpurun_out/${TAG}_time_update.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 4300 -c 2100 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/${TAG}_ncu_bench.log 2>&1
python tools/summarize_launches.py gpurun_out/${TAG}_launches.csv gpurun_out/${TAG}_launches_summary.csv | head -16
timeout 600 ncu --set full --clock-control none --import-source on -k regex:env_step_kernel -s 30 -c 1 -f -o gpurun_out/${TAG}_env python tools/time_env.py 4096 40 > gpurun_out/${TAG}_ncu_env.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gemm_tf32|ppo_heads|apply_kernel|gather' -s 195 -c 20 -f -o gpurun_out/${TAG}_upd python tools/prof_update.py 4096 64 3 > gpurun_out/${TAG}_ncu_upd.log 2>&1
python tools/ncu_report.py gpurun_out/${TAG}_env.ncu-rep > gpurun_out/${TAG}_env_ncu_summary.txt
python tools/ncu_report.py gpurun_out/${TAG}_upd.ncu-rep > gpurun_out/${TAG}_update_kernels_ncu_summary.txt
python tools/ncu_traffic.py gpurun_out/${TAG}_env.ncu-rep gpurun_out/${TAG}_upd.ncu-rep gpurun_out/${TAG}_traffic.json | head -20
ls -la gpurun_out | grep ${TAG} | tail -20
