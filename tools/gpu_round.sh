#!/bin/bash
# One gpurun call: parity tests, bench, ncu launch list, ncu full captures.  Usage: tools/gpu_round.sh <tag>
TAG=${1:-r1}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 3000 gpurun_out/${TAG}_bench.json
timeout 300 python tools/time_env.py 4096 200 > gpurun_out/${TAG}_time_env.log 2>&1; cat gpurun_out/${TAG}_time_env.log
timeout 300 python tools/prof_update.py 4096 64 25 > gpurun_out/${TAG}_time_update.log 2>&1; cat gpurun_out/${TAG}_time_update.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 12000 -c 6200 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:env_step_kernel -s 30 -c 2 -f -o gpurun_out/${TAG}_env python tools/time_env.py 4096 40 > gpurun_out/${TAG}_ncu_env.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gemm|fused|heads' -s 600 -c 30 -f -o gpurun_out/${TAG}_upd python tools/prof_update.py 4096 64 3 > gpurun_out/${TAG}_ncu_upd.log 2>&1
ls -la gpurun_out
